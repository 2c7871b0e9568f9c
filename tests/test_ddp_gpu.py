"""GPU (>= 2 devices): the data-parallel exchange of the training step on real NCCL -- bucketed, backward-overlapped
all-reduce == sum of the ranks' own gradients, identical weights after the update (tests/manual/ddp_grad_check.py under
torchrun).  The host-side logic has a world-2 gloo test on the CPU (tests/test_distributed_cpu.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bucketed_allreduce_equals_sum_of_rank_gradients():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "manual", "ddp_grad_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0 and "DDP_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
