#!/usr/bin/env python
"""PsRoIAlign at the model's shape (8 x 490 x 30 x 30, 1000 RoIs per image, 7x7 max): timing of every variant."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import xdet_b200  # noqa: E402,F401
from tests import workloads  # noqa: E402
from xdet_b200 import ops  # noqa: E402

N, C, R = 8, 490, 1000
x = torch.from_numpy(workloads.make_map(N, C, 30, 30, seed=4)).cuda()
rois = torch.from_numpy(workloads.make_rois(N, R, seed=5)).cuda()
import sys as _sys
if "--relu" in _sys.argv:  # the model's thin feature map is post-ReLU: ~half of it exact zeros
    x = torch.relu(x)
    print("post-ReLU map")
for variant in ("planes", "select", "gather"):
    for _ in range(3):
        ops.ps_roi_align(x, rois, 7, 7, "max", variant=variant)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.ps_roi_align(x, rois, 7, 7, "max", variant=variant)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print("%-7s median %.1f us  min %.1f us" % (variant, ts[5], ts[0]))
