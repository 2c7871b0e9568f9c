#!/usr/bin/env python
"""A few eager steps of the benchmarked inference workload (ResNet-50 light-head, 8 x 480^2, detections included) for
profiling under ncu: one untimed warm-up step, then --steps steps between cudaProfilerStart/Stop."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import xdet_b200  # noqa: E402,F401
from xdet_b200 import light_head_rfcn_eval as lh  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="f16x2")
    ap.add_argument("--backbone", default="resnet50")
    ap.add_argument("--size", type=int, default=480)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    params = lh.make_params(train_image_size=args.size, backbone=args.backbone, rpn_min_size=16.0 / args.size,
                            precision=args.precision)
    model = lh.LightHeadRFCN(params, seed=0)
    rng = np.random.default_rng(1)
    x = torch.from_numpy((rng.random((args.batch, 3, args.size, args.size), dtype=np.float32) * 2 - 1)).cuda()
    torch.cuda.profiler.stop()
    model(x, detections=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(args.steps):
        model(x, detections=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
