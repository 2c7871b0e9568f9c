#!/usr/bin/env python
"""Per-stage error of the Light-Head R-CNN forward against the fp32 CPU oracle, for the bf16 throughput path and
the fp32x3 parity mode (same seeded variables, 2 x 160x160 images).  GPU only.
    python tools/parity_report.py [--backbone resnet50|xception]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import xdet_b200  # noqa: F401,E402
from oracle import net as onet  # noqa: E402
from oracle import proposals as op  # noqa: E402
from xdet_b200 import light_head_rfcn_eval as lh  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def run(backbone, precision, seed):
    params = lh.make_params(train_image_size=160, rpn_pre_nms_top_n=600, rpn_post_nms_top_n=100,
                            rpn_min_size=16.0 / 160, backbone=backbone, precision=precision)
    model = lh.LightHeadRFCN(params, seed=seed)
    g = torch.Generator(device="cuda").manual_seed(seed + 100)
    images = torch.rand((2, 3, 160, 160), generator=g, device="cuda") * 2 - 1
    keys = torch.rand((2, 100), generator=g, device="cuda")
    out = model(images, shuffle_keys=keys)
    torch.cuda.synchronize()
    fm = out["rpn_feat_map"].shape[1]
    anchors = op.layer_anchors((160, 160), (fm, fm), [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], [0.1], [1., 2., .5], 16)
    ref = onet.model(images.cpu().numpy(), model.store.state_dict(), params, anchors, shuffle_keys=keys.cpu().numpy())
    rpn = out["rpn_out"].cpu().numpy()
    rows = [("rpn_feat_map (rel)", rel(out["rpn_feat_map"].float().permute(0, 3, 1, 2).cpu().numpy(), ref["rpn_feat_map"])),
            ("backbone_feat (rel)", rel(out["backbone_feat"].float().permute(0, 3, 1, 2).cpu().numpy(), ref["backbone_feat"])),
            ("large_sep_feature (rel)", rel(out["large_sep_feature"].cpu().numpy(), ref["large_sep_feature"])),
            ("rpn_cls (rel)", rel(rpn[..., :44], ref["rpn_cls"])), ("rpn_box (rel)", rel(rpn[..., 44:], ref["rpn_box"])),
            ("rpn_object_score (abs)", float(np.abs(out["rpn_object_score"].cpu().numpy() - ref["rpn_object_score"]).max())),
            ("proposals_bboxes (abs)", float(np.abs(out["proposals_bboxes"].cpu().numpy() - ref["proposals_bboxes"]).max())),
            ("cls_score (abs)", float(np.abs(out["cls_score"].cpu().numpy().reshape(-1, 21) - ref["cls_score"]).max())),
            ("bboxes_reg (abs)", float(np.abs(out["bboxes_reg"].cpu().numpy().reshape(-1, 4) - ref["bboxes_reg"]).max())),
            ("head_cls_score (abs)", float(np.abs(out["head_cls_score"].cpu().numpy() - ref["head_cls_score"]).max())),
            ("bboxes_predict (abs)", float(np.abs(out["bboxes_predict"].cpu().numpy() - ref["bboxes_predict"]).max()))]
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backbone", default="resnet50,xception")
    ap.add_argument("--max-k", default="1536", help="comma list of ops.conv.PARITY_MAX_K values for the fp32x3 column(s)")
    args = ap.parse_args()
    from xdet_b200.ops import conv as conv_ops
    for bb, seed in (("resnet50", 3), ("xception", 5)):
        if bb not in args.backbone.split(","):
            continue
        cols = {}
        for prec in ["bf16"] + ["fp32x3/K%s" % k for k in args.max_k.split(",")]:
            try:
                if "/" in prec:
                    conv_ops.PARITY_MAX_K = int(prec.split("K")[1])
                cols[prec] = run(bb, prec.split("/")[0], seed)
            except Exception as e:  # noqa: BLE001
                print(bb, prec, "FAILED:", repr(e))
        print("== %s: max error vs the fp32 CPU oracle (no injected intermediates)" % bb)
        names = [n for n, _ in next(iter(cols.values()))] if cols else []
        for i, n in enumerate(names):
            print("  %-26s" % n + "".join("  %s %.3e" % (p, cols[p][i][1]) for p in cols))


if __name__ == "__main__":
    main()
