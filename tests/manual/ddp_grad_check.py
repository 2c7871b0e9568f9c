#!/usr/bin/env python
"""2+ GPUs under torchrun: the bucketed, backward-overlapped gradient all-reduce of LightHeadTrainer equals the SUM
of the ranks' own gradients on their own image shards, with and without overlap, for both backbones.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tests/manual/ddp_grad_check.py
Prints DDP_OK on rank 0.  (Weight-gradient kernels accumulate pixel splits with fp32 atomics, so two runs of the same
step differ in the last bits: the comparison is relative to each bucket's magnitude.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import xdet_b200  # noqa: F401,E402
from xdet_b200 import light_head_rfcn_train as lt  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for backbone in ("resnet50", "xception"):
        params = lt.make_params(train_image_size=160, batch_size=2, rpn_pre_nms_top_n=600, rpn_post_nms_top_n=200,
                                rpn_min_size=16.0 / 160, rpn_anchors_per_image=64, roi_one_image=32,
                                ohem_roi_one_image=16, backbone=backbone, resnet_layers=(1, 1, 1, 1))
        tr = lt.LightHeadTrainer(params, seed=0)  # same seed: identical weights on every rank
        batch = lt.synthetic_batch(params, 2, seed=3 + 1000 * rank)  # own shard
        tr.local_only = True
        out = tr.step(*batch, apply_update=False)
        inj = {k: out[k] for k in ("rpn_idx", "rois_all", "roi_idx", "ohem_idx")}
        own = tr.grads.clone()
        gathered = [torch.empty_like(own) for _ in range(world)]
        dist.all_gather(gathered, own)
        want = torch.stack(gathered).double().sum(0)
        tr.local_only = False
        for overlap in (True, False):
            tr.overlap_allreduce = overlap
            tr.step(*batch, apply_update=False, inject=inj)
            torch.cuda.synchronize()
            got = tr.grads.double()
            for name, a, b in tr.buckets:
                if b <= a:
                    continue
                err = float((got[a:b] - want[a:b]).abs().max() / (want[a:b].abs().max() + 1e-30))
                good = err < 1e-4
                ok = ok and good
                if rank == 0:
                    print("%-9s overlap=%-5s bucket %-7s %9d floats  max rel err %.2e %s" % (
                        backbone, overlap, name, b - a, err, "" if good else "FAIL"))
        # and the update itself: after one optimizer step every rank holds the same weights
        tr.overlap_allreduce = True
        tr.step(*batch, inject=inj)
        torch.cuda.synchronize()
        sig = torch.stack([v.double().sum() for _, v in sorted(tr.store.state_dict().items())])
        sigs = [torch.empty_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        same = all(torch.allclose(s, sigs[0], rtol=1e-9, atol=0) for s in sigs)
        ok = ok and same
        if rank == 0:
            print("%-9s weights after one step identical across ranks: %s" % (backbone, same))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DDP_OK" if int(flag) == 1 else "DDP_FAIL")
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
