"""CPU, world_size 2, gloo: the host-side multi-GPU logic of the training step -- batch sharding and the single
gradient all-reduce (sum, then 1/world inside the optimizer) -- equals the mean of independent per-shard gradients."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import xdet_b200  # noqa: F401
    from xdet_b200.light_head_rfcn_train import allreduce_gradients, shard_batch
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_batch(64, world, rank)
    g = torch.Generator().manual_seed(0)
    per_image = torch.randn((64, 1000), generator=g)           # every rank can compute every image's "gradient"
    flat = per_image[lo:hi].sum(0)                               # ... but only owns its shard
    w = allreduce_gradients(flat)
    q.put((rank, lo, hi, w, flat.clone(), per_image.sum(0)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 32), (32, 64)]
    for rank, lo, hi, w, flat, total in res:
        assert w == 2
        assert torch.allclose(flat, total, atol=1e-4)          # the sum over ranks == the full-batch gradient
    assert torch.equal(res[0][4], res[1][4])                    # identical on every rank -> identical updates


def test_single_process_is_a_no_op():
    sys.path.insert(0, ROOT)
    import xdet_b200  # noqa: F401
    from xdet_b200.light_head_rfcn_train import allreduce_gradients, learning_rate, make_params, shard_batch
    t = torch.ones(8)
    assert allreduce_gradients(t) == 1 and torch.equal(t, torch.ones(8))
    assert shard_batch(64, 8, 3) == (24, 32)
    p = make_params()
    assert learning_rate(p, 0) == 1e-3 and abs(learning_rate(p, 70000) - 8e-4) < 1e-12 and learning_rate(p, 90000) == 1e-4
