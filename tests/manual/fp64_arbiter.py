#!/usr/bin/env python
"""Whose rounding is a delta?  Convolution stages of the Light-Head R-CNN forward in float64 on the CPU (the
oracle's graph, exact for this purpose) against (a) the fp32 CPU oracle and (b) the GPU path in each precision.
GPU only; developer tool.
    python tests/manual/fp64_arbiter.py [--backbone resnet50] [--size 480] [--batch 2] [--chunks 1,2,12]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import xdet_b200  # noqa: F401,E402
from oracle import net as onet  # noqa: E402
from xdet_b200 import light_head_rfcn_eval as lh  # noqa: E402
from xdet_b200.net import variables as _variables  # noqa: E402

_variables.RANDOMIZE_BN = True  # non-trivial batch-norm variables: a folding error cannot hide
from xdet_b200.ops import conv as conv_ops  # noqa: E402

KEYS = ("rpn_feat_map", "backbone_feat", "large_sep_feature", "rpn_cls", "rpn_box")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


def gpu_stages(params, imgs, sd):
    model = lh.LightHeadRFCN(params, seed=0, state_dict=sd)
    out = model(torch.from_numpy(imgs).cuda())
    torch.cuda.synchronize()
    rpn = out["rpn_out"].float().cpu().numpy()
    return {"rpn_feat_map": out["rpn_feat_map"].float().permute(0, 3, 1, 2).cpu().numpy(),
            "backbone_feat": out["backbone_feat"].float().permute(0, 3, 1, 2).cpu().numpy(),
            "large_sep_feature": out["large_sep_feature"].cpu().numpy(), "rpn_cls": rpn[..., :44], "rpn_box": rpn[..., 44:]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backbone", default="resnet50")
    ap.add_argument("--size", type=int, default=480)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--chunks", default="1,2,12")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    base = lh.make_params(train_image_size=args.size, backbone=args.backbone, rpn_min_size=16.0 / args.size)
    imgs = (np.random.default_rng(1).random((args.batch, 3, args.size, args.size), dtype=np.float32) * 2 - 1)
    sd = lh.LightHeadRFCN(dict(base, precision="bf16"), seed=0)
    sd(torch.from_numpy(imgs[:1]).cuda())  # creates every variable
    sd = {k: v.clone() for k, v in sd.store.state_dict().items()}
    exact = onet.conv_stages(imgs, sd, base, torch.float64)
    rows = {"fp32 CPU oracle": onet.conv_stages(imgs, sd, base, torch.float32)}
    for ck in [int(c) for c in args.chunks.split(",")]:
        conv_ops.F16X2_CHUNK_KB = ck
        rows["GPU f16x2 chunk_kb=%d" % ck] = gpu_stages(dict(base, precision="f16x2"), imgs, sd)
    conv_ops.F16X2_CHUNK_KB = 0
    rows["GPU fp32x3"] = gpu_stages(dict(base, precision="fp32x3"), imgs, sd)
    rows["GPU bf16"] = gpu_stages(dict(base, precision="bf16"), imgs, sd)
    print("max |x - float64| / max |float64| per stage (%s, %d x %d^2)" % (args.backbone, args.batch, args.size))
    print("%-24s" % "" + "".join("%20s" % k for k in KEYS))
    for name, st in rows.items():
        print("%-24s" % name + "".join("%20.3e" % rel(st[k], exact[k]) for k in KEYS))
    print(json.dumps({n: {k: rel(st[k], exact[k]) for k in KEYS} for n, st in rows.items()}))


if __name__ == "__main__":
    main()
