#!/usr/bin/env python
"""Run ONE depthwise 3x3 layer shape (ncu captures / quick timing):  python tools/dw_one.py N H W C [--dil D]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import xdet_b200  # noqa: E402,F401
from xdet_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("dims", type=int, nargs=4)
ap.add_argument("--dil", type=int, default=1)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
N, H, W, C = a.dims
x = torch.randn((N, H, W, C), device="cuda").to(torch.bfloat16)
w = torch.randn((9, C), device="cuda")
from xdet_b200 import _native  # noqa: E402
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ap2 = None
for rows in (0, 1):
  _native.lib().xdet_set_depthwise_rows(rows)
  for _ in range(3):
    ops.depthwise3x3(x, w, dilation=a.dil, relu_in=True)
  ts = []
  for _ in range(a.reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.depthwise3x3(x, w, dilation=a.dil, relu_in=True)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
  ts.sort()
  nbytes = 2 * x.numel() * 2
  print("%dx%dx%dx%d dil %d rolling_rows=%d: median %.1f us  %.0f GB/s" % (N, H, W, C, a.dil, rows, ts[len(ts) // 2], nbytes / ts[len(ts) // 2] / 1e3))
