#!/usr/bin/env python
"""Error against the fp32 CPU oracle AND throughput of one precision mode of the Light-Head R-CNN forward at a
BASELINE shape (default: config 2, ResNet-50, 480x480).  GPU only; developer tool (tests/ hold the assertions).

    python tests/manual/parity_bench.py --precision fp32x3 --backbone resnet50 --size 480 --parity-batch 2 --batch 8
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import xdet_b200  # noqa: F401,E402
from xdet_b200 import light_head_rfcn_eval as lh  # noqa: E402
from xdet_b200.net import variables as _variables  # noqa: E402

_variables.RANDOMIZE_BN = True  # non-trivial batch-norm variables: a folding error cannot hide


def deltas(out, ref):
    def rel(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))

    def mabs(a, b):
        return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())
    rpn = out["rpn_out"].float().cpu().numpy()
    return {
        "rpn_feat_map_rel": rel(out["rpn_feat_map"].float().permute(0, 3, 1, 2).cpu().numpy(), ref["rpn_feat_map"]),
        "backbone_feat_rel": rel(out["backbone_feat"].float().permute(0, 3, 1, 2).cpu().numpy(), ref["backbone_feat"]),
        "large_sep_feature_rel": rel(out["large_sep_feature"].cpu().numpy(), ref["large_sep_feature"]),
        "rpn_cls_rel": rel(rpn[..., :44], ref["rpn_cls"]), "rpn_box_rel": rel(rpn[..., 44:], ref["rpn_box"]),
        "rpn_object_score_abs": mabs(out["rpn_object_score"].cpu().numpy(), ref["rpn_object_score"]),
        "proposals_bboxes_abs": mabs(out["proposals_bboxes"].cpu().numpy(), ref["proposals_bboxes"]),
        "cls_score_abs": mabs(out["cls_score"].float().cpu().numpy().reshape(-1, 21), ref["cls_score"]),
        "cls_score_max": float(np.abs(ref["cls_score"]).max()),
        "bboxes_reg_abs": mabs(out["bboxes_reg"].float().cpu().numpy().reshape(-1, 4), ref["bboxes_reg"]),
        "head_cls_score_abs": mabs(out["head_cls_score"].cpu().numpy(), ref["head_cls_score"]),
        "bboxes_predict_abs": mabs(out["bboxes_predict"].cpu().numpy(), ref["bboxes_predict"]),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp32x3")
    ap.add_argument("--backbone", default="resnet50")
    ap.add_argument("--size", type=int, default=480)
    ap.add_argument("--parity-batch", type=int, default=2)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-time", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--chunk-kb", type=int, default=0)
    ap.add_argument("--cluster", type=int, default=1, help="f16x2: 1 = independent CTAs, 2 = 2-CTA clusters everywhere, "
                    "0 = let the tuner pick per layer (with --autotune)")
    ap.add_argument("--autotune", action="store_true")
    args = ap.parse_args()
    from xdet_b200.ops import conv as conv_ops
    conv_ops.F16X2_CHUNK_KB = args.chunk_kb
    conv_ops.F16X2_CLUSTER = args.cluster
    conv_ops.AUTOTUNE = args.autotune
    params = lh.make_params(train_image_size=args.size, backbone=args.backbone, rpn_min_size=16.0 / args.size,
                            precision=args.precision)
    model = lh.LightHeadRFCN(params, seed=0)
    rec = {"precision": args.precision, "backbone": args.backbone, "size": args.size, "cluster": args.cluster,
           "autotune": args.autotune}
    rng = np.random.default_rng(1)
    imgs = (rng.random((max(args.batch, args.parity_batch), 3, args.size, args.size), dtype=np.float32) * 2 - 1)
    if not args.no_parity:
        from oracle import net as onet
        from oracle import proposals as op
        torch.set_num_threads(os.cpu_count() or 1)
        x = torch.from_numpy(imgs[:args.parity_batch]).cuda()
        keys = torch.rand((args.parity_batch, params["rpn_post_nms_top_n"]), device="cuda")
        out = model(x, shuffle_keys=keys)
        torch.cuda.synchronize()
        fm = args.size // 16
        anchors = op.layer_anchors((args.size, args.size), (fm, fm), [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], [0.1],
                                   [1., 2., .5], 16)
        t0 = time.perf_counter()
        ref = onet.model(x.cpu().numpy(), model.store.state_dict(), params, anchors, shuffle_keys=keys.cpu().numpy())
        rec["oracle_s"] = time.perf_counter() - t0
        rec["parity_batch"] = args.parity_batch
        rec["deltas"] = deltas(out, ref)
        # how many proposal rows differ at all (selection flips), independent of magnitudes
        pa, pb = out["proposals_bboxes"].cpu().numpy(), np.asarray(ref["proposals_bboxes"], np.float32)
        rec["proposal_rows_differing_gt_1e-4"] = int((np.abs(pa - pb).max(axis=-1) > 1e-4).sum())
        # paired by box instead of by row position (an early NMS flip shifts every later row)
        unp, worst = 0, {k: 0.0 for k in ("cls_score", "bboxes_reg", "head_cls_score", "bboxes_predict")}
        R = pa.shape[1]
        for n in range(pa.shape[0]):
            d = np.abs(pa[n][:, None, :] - pb[n][None, :, :]).max(axis=-1)
            j = d.argmin(axis=1)
            ok = d[np.arange(R), j] < 1e-4
            unp += int((~ok).sum())
            for k in worst:
                a = out[k].float().cpu().numpy().reshape(ref[k].shape)
                worst[k] = max(worst[k], float(np.abs(a[n * R + np.nonzero(ok)[0]] - ref[k][n * R + j[ok]]).max()))
        rec["proposals_without_partner"] = unp
        rec["paired_row_deltas"] = worst
    if not args.no_time:
        x = torch.from_numpy(imgs[:args.batch]).cuda()
        run = lambda: model(x, detections=True)  # noqa: E731
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            run()
        e1.record()
        torch.cuda.synchronize()
        rec["eager_ms_per_step"] = e0.elapsed_time(e1) / args.steps
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                run()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            rec["graph_ms_per_step"] = e0.elapsed_time(e1) / args.steps
            rec["images_per_s"] = args.batch / (rec["graph_ms_per_step"] * 1e-3)
        except Exception as e:  # noqa: BLE001
            rec["graph_error"] = repr(e)
        rec["batch"] = args.batch
    line = json.dumps(rec)
    print(line)
    if args.out:
        with open(args.out, "a") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
