#!/usr/bin/env python
"""2+ GPUs under torchrun: the bucketed, backward-overlapped gradient all-reduce of LightHeadTrainer equals the SUM
of the ranks' own gradients on their own image shards, with and without overlap, for both backbones.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tests/manual/ddp_grad_check.py
Prints DDP_OK on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import xdet_b200  # noqa: F401,E402
from xdet_b200 import light_head_rfcn_train as lt  # noqa: E402


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def same_on_all_ranks(t, world):
    """Bit-equality of a tensor across ranks (all_gather of the raw bits)."""
    bits = t.contiguous().view(torch.int32)
    got = [torch.empty_like(bits) for _ in range(world)]
    dist.all_gather(got, bits)
    return all(torch.equal(g, got[0]) for g in got)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True

    def report(name, good, detail=""):
        nonlocal ok
        ok = ok and good
        if rank == 0:
            print("%-60s %s %s" % (name, "ok  " if good else "FAIL", detail))

    for backbone in ("resnet50", "xception"):
        params = lt.make_params(train_image_size=160, batch_size=2, rpn_pre_nms_top_n=600, rpn_post_nms_top_n=200,
                                rpn_min_size=16.0 / 160, rpn_anchors_per_image=64, roi_one_image=32,
                                ohem_roi_one_image=16, backbone=backbone, resnet_layers=(1, 1, 1, 1))
        tr = lt.LightHeadTrainer(params, seed=0)  # same seed: identical weights on every rank
        batch = lt.synthetic_batch(params, 2, seed=3 + 1000 * rank)  # own shard
        n = tr.grads.numel()
        # (1) the exchange alone, on a known pattern: every element of the flat buffer is summed over the ranks exactly
        # once, whatever order the buckets are sent in (integers below 2^24: exact in fp32)
        for overlap in (True, False):
            tr.overlap_allreduce = overlap
            pat = (torch.arange(n, device="cuda") % 4093).float()
            tr.grads.copy_(pat * (rank + 1))
            tr._pending = set(b[0] for b in tr.buckets)
            for name, _, _ in reversed(tr.buckets[1:]):  # backward order; the first bucket is left to the finish
                tr._stage_done(name)
            w = tr._allreduce_finish()
            torch.cuda.synchronize()
            want = pat * (world * (world + 1) // 2)
            report("%s exchange pattern, overlap=%s" % (backbone, overlap), w == world and torch.equal(tr.grads, want))
        # (2) the real step: this rank's own gradient first, then the data-parallel step
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
            # every rank holds the SAME reduced gradient, bit for bit: nothing was added to a bucket after it was sent
            report("%s step, overlap=%s: reduced gradient identical on all ranks" % (backbone, overlap),
                   same_on_all_ranks(tr.grads, world))
            # ... and it is the sum of the ranks' own gradients (two runs of the step differ by the rounding noise that
            # batch statistics amplify -- column sums and weight-gradient splits use fp32 atomics -- hence a cosine)
            for name, a, b in tr.buckets:
                if b > a:
                    c = cosine(tr.grads[a:b], want[a:b])
                    report("%s step, overlap=%s: bucket %-7s (%9d floats) vs sum of own" % (backbone, overlap, name, b - a),
                           c > 0.8, "cos %.4f" % c)
        # (3) the update: after an optimizer step every rank holds the same trainable variables, bit for bit
        tr.overlap_allreduce = True
        tr.step(*batch, inject=inj)
        torch.cuda.synchronize()
        same = all(same_on_all_ranks(v, world) for k, v in sorted(tr.store.state_dict().items())
                   if "moving_" not in k)  # (batch-norm moving statistics stay local to each rank by design)
        report("%s weights identical on all ranks after one step" % backbone, same)
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
