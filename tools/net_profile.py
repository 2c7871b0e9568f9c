#!/usr/bin/env python
"""Kernel-level time breakdown of one Light-Head R-CNN forward pass (CUPTI through torch.profiler: real,
pipelined kernel durations, not the serialised cold-cache ones of an ncu launch list)."""
import argparse
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import xdet_b200  # noqa: E402,F401
from xdet_b200 import light_head_rfcn_eval as lh  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=480)
    ap.add_argument("--backbone", default="resnet50")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--list", action="store_true", help="also print every launch of the last step in order")
    ap.add_argument("--train", action="store_true", help="profile the training step instead of inference")
    ap.add_argument("--precision", default="bf16")
    args = ap.parse_args()
    if args.train:
        from xdet_b200 import light_head_rfcn_train as lt
        tparams = lt.make_params(train_image_size=args.size, batch_size=args.batch, backbone=args.backbone)
        trainer = lt.LightHeadTrainer(tparams, seed=0)
        tb = lt.synthetic_batch(tparams, args.batch, seed=3)
        model = lambda _images, detections=False: trainer.step(*tb)  # noqa: E731
        images = None
    else:
        params = lh.make_params(train_image_size=args.size, backbone=args.backbone, precision=args.precision,
                                rpn_min_size=16.0 / args.size)
        model = lh.LightHeadRFCN(params, seed=0)
        g = torch.Generator(device="cuda").manual_seed(1)
        images = torch.rand((args.batch, 3, args.size, args.size), generator=g, device="cuda") * 2 - 1
    for _ in range(3):
        model(images, detections=True)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            model(images, detections=True)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    agg = collections.OrderedDict()
    for e in evs:
        a = agg.setdefault(e.name[:70], [0, 0.0])
        a[0] += 1
        a[1] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
    tot = sum(v[1] for v in agg.values())
    print("kernel time per step: %.1f us over %d launches" % (tot / args.steps, len(evs) // args.steps))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %5d  %9.1f us/step  %5.1f%%" % (k, v[0] // args.steps, v[1] / args.steps, 100 * v[1] / tot))
    if args.list:
        n = len(evs) // args.steps
        evs.sort(key=lambda e: e.time_range.start)
        for e in evs[-n:]:
            print("%-72s %9.1f" % (e.name[:70], e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total))


if __name__ == "__main__":
    main()
