#!/usr/bin/env python
"""Run ONE convolution layer shape a few times (for ncu captures and quick timing).
   python tools/conv_one.py N H W Cin Cout KH KW [--res] [--out2] [--stride S] [--dil D] [--block_n B] [--layout L]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import xdet_b200  # noqa: E402,F401
from xdet_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dims", type=int, nargs=7)
    ap.add_argument("--res", action="store_true")
    ap.add_argument("--out2", action="store_true")
    ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--dil", type=int, default=1)
    ap.add_argument("--block_n", type=int, default=0)
    ap.add_argument("--layout", default="nhwc_bf16")
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    N, H, W, Cin, Cout, KH, KW = a.dims
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((N, H, W, Cin), generator=g, device="cuda").to(torch.bfloat16)
    w = ops.pack_conv_weight(torch.randn((Cout, Cin, KH, KW), generator=g, device="cuda") / (Cin * KH * KW) ** 0.5)
    scale = torch.rand(Cout, device="cuda") + 0.5
    bias = torch.randn(Cout, device="cuda")
    Ho, Wo = -(-H // a.stride), -(-W // a.stride)
    kws = dict(scale=scale, bias=bias, relu=not a.res, dilation=(a.dil, a.dil), strides=(a.stride, a.stride),
               out_layout=a.layout, block_n=a.block_n)
    if a.stride > 1:
        kws["padding"] = ((KH - 1) // 2, (KW - 1) // 2, Ho, Wo)
    if a.res:
        kws["residual"] = torch.randn((N, Ho, Wo, Cout), generator=g, device="cuda").to(torch.bfloat16)
    if a.out2:
        kws["out2"] = torch.empty((N, Ho, Wo, Cout), dtype=torch.bfloat16, device="cuda")
        kws["scale2"], kws["bias2"] = scale, bias
    out = ops.conv2d_nhwc(x, w, Cout, KH, KW, **kws)
    kws["out"] = out
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ops.conv2d_nhwc(x, w, Cout, KH, KW, **kws)
    ts = []
    for _ in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv2d_nhwc(x, w, Cout, KH, KW, **kws)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    flops = 2.0 * N * Ho * Wo * Cout * Cin * KH * KW
    print("median %.1f us  min %.1f us  %.1f TFLOP/s (L2 flushed before every launch)" % (ts[len(ts) // 2], ts[0], flops / ts[len(ts) // 2] / 1e6))


if __name__ == "__main__":
    main()
