#!/usr/bin/env python
"""A few eager training steps (ResNet-50 light-head, 8 x 480^2, the reference's flags) for profiling under ncu: one
untimed warm-up step, then --steps steps between cudaProfilerStart/Stop."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import xdet_b200  # noqa: E402,F401
from xdet_b200 import light_head_rfcn_train as lt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=480)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--backbone", default="resnet50")
    args = ap.parse_args()
    params = lt.make_params(train_image_size=args.size, batch_size=args.batch, backbone=args.backbone)
    trainer = lt.LightHeadTrainer(params, seed=0)
    tb = lt.synthetic_batch(params, args.batch, seed=3)
    torch.cuda.profiler.stop()
    for _ in range(2):
        trainer.step(*tb)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(args.steps):
        trainer.step(*tb)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
