#!/usr/bin/env python
"""Per-shape timing of the training step's batch-norm kernels (statistics, finalize, normalise+ReLU, backward) at the
activation shapes of ResNet-50 on 8 x 480^2: microseconds per call and the HBM rate of the bytes each one must move."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import xdet_b200  # noqa: E402,F401
from xdet_b200 import ops  # noqa: E402
from xdet_b200.ops import train as T  # noqa: E402

SHAPES = [(460800, 64), (115200, 64), (115200, 256), (115200, 128), (28800, 128), (28800, 512), (28800, 256),
          (7200, 256), (7200, 1024), (7200, 512), (7200, 2048)]


def timed(fn, flush, reps=20):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3


def main():
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    print("%-16s %22s %22s %22s %22s" % ("rows x C", "col_stats us (GB/s)", "affine_relu", "bn_bwd (reduce+apply)", "finalize"))
    for rows, C in SHAPES:
        x = torch.randn(rows, C, device=dev).to(torch.bfloat16)
        dy = torch.randn(rows, C, device=dev).to(torch.bfloat16)
        gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        st = T.bn_train(x, gamma, beta, 1e-5)
        nb = rows * C * 2
        out = []
        for fl in (None, flush):
            t1 = timed(lambda: T.col_stats(x, True), fl)
            t2 = timed(lambda: ops.affine_relu(x, st.scale, st.shift, True), fl)
            t3 = timed(lambda: T.bn_relu_bwd(dy, x, st, True), fl)
            t4 = timed(lambda: T.bn_train(x, gamma, beta, 1e-5), fl) - t1
            out.append("%7.1f (%5.0f) %7.1f (%5.0f) %7.1f (%5.0f) %6.1f" % (t1, nb / t1 / 1e3, t2, 2 * nb / t2 / 1e3,
                                                                        t3, 5 * nb / t3 / 1e3, t4))
        print("%-16s warm: %s | cold: %s" % ("%d x %d" % (rows, C), out[0], out[1]))


if __name__ == "__main__":
    main()
