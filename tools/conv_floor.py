#!/usr/bin/env python
"""Fixed cost of one conv_gemm_kernel launch: tiny and small problems replayed back to back from a CUDA graph,
with and without programmatic dependent launch.  GPU only."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import xdet_b200  # noqa: E402,F401
from xdet_b200 import _native, ops  # noqa: E402


def time_graph(fn, reps=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps * 1e3)
    return min(ts)


def main():
    gen = torch.Generator(device="cuda").manual_seed(0)
    cases = []
    for (N, H, W, Cin, Cout, k, res, out2) in [
            (8, 120, 120, 64, 256, 1, True, True), (8, 120, 120, 64, 256, 1, False, False),
            (8, 60, 60, 128, 512, 1, True, True), (8, 30, 30, 256, 1024, 1, True, True),
            (8, 30, 30, 512, 2048, 1, True, True), (8, 30, 30, 1024, 256, 1, False, False),
            (8, 60, 60, 512, 128, 1, False, False), (8, 120, 120, 256, 64, 1, False, False),
            (8, 30, 30, 2048, 512, 1, False, False), (8, 30, 30, 256, 256, 3, False, False),
            (8, 60, 60, 128, 128, 3, False, False), (8, 120, 120, 64, 64, 3, False, False)]:
        for bn in (64, 128, 256):
            if bn <= Cout:
                cases.append((N, H, W, Cin, Cout, k, res, out2, bn))
    for pdl in (1,):
        _native.lib().xdet_set_conv_pdl(pdl)
        for (N, H, W, Cin, Cout, k, res, out2, bn) in cases:
            x = torch.randn((N, H, W, Cin), generator=gen, device="cuda").to(torch.bfloat16)
            w = ops.pack_conv_weight(torch.randn((Cout, Cin, k, k), generator=gen, device="cuda") / (Cin * k * k) ** 0.5)
            scale = torch.rand(Cout, device="cuda") + 0.5
            bias = torch.randn(Cout, device="cuda")
            kws = dict(scale=scale, bias=bias, relu=not res, block_n=bn)
            if res:
                kws["residual"] = torch.randn((N, H, W, Cout), generator=gen, device="cuda").to(torch.bfloat16)
            if out2:
                kws["out2"] = torch.empty((N, H, W, Cout), dtype=torch.bfloat16, device="cuda")
                kws["scale2"], kws["bias2"] = scale, bias
            kws["out"] = ops.conv2d_nhwc(x, w, Cout, k, k, **kws)
            for eg in (1, 2):
                kws["epi_groups"] = eg
                t = time_graph(lambda: ops.conv2d_nhwc(x, w, Cout, k, k, **kws))
                print("pdl=%d  %dx%dx%dx%d -> %d k%d %s%s bn=%d eg=%d : %7.2f us" %
                      (pdl, N, H, W, Cin, Cout, k, "R" if res else "", "2" if out2 else "", bn, eg, t), flush=True)
    _native.lib().xdet_set_conv_pdl(1)


if __name__ == "__main__":
    main()
