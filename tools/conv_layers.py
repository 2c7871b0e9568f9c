#!/usr/bin/env python
"""Per-layer timing of every convolution / dense launch of one Light-Head R-CNN forward pass.

Records the arguments of each ``ops.conv2d_nhwc`` call during one eager pass, then replays every call
``--reps`` times back to back (CUDA events around the batch) for each N-tile width in ``--block_n``.
Prints one line per layer: shape, algorithmic GFLOP, and per width the time, TFLOP/s and GB/s (algorithmic
bytes = input + weights + output (+ residual, + second output) once)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import xdet_b200  # noqa: E402,F401
from xdet_b200 import light_head_rfcn_eval as lh  # noqa: E402
from xdet_b200 import ops  # noqa: E402
from xdet_b200.ops import conv as conv_mod  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=480)
    ap.add_argument("--backbone", default="resnet50")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--block_n", default="0,64,128,256")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    widths = [int(v) for v in args.block_n.split(",")]

    params = lh.make_params(train_image_size=args.size, backbone=args.backbone)
    model = lh.LightHeadRFCN(params, seed=0)
    g = torch.Generator(device="cuda").manual_seed(1)
    images = torch.rand((args.batch, 3, args.size, args.size), generator=g, device="cuda") * 2 - 1
    model(images)
    torch.cuda.synchronize()

    calls = []
    orig = conv_mod.conv2d_nhwc

    def rec(x, w, cout, kh, kw, **kws):
        out = orig(x, w, cout, kh, kw, **kws)
        calls.append((x, w, cout, kh, kw, dict(kws), out))
        return out

    conv_mod.conv2d_nhwc = rec
    ops.conv2d_nhwc = rec
    import xdet_b200.net.resnet_v2 as rn
    import xdet_b200.net.xception_body as xb
    for mod in (rn, xb):
        if hasattr(mod, "ops"):
            mod.ops.conv2d_nhwc = rec
    model(images)
    torch.cuda.synchronize()
    conv_mod.conv2d_nhwc = orig
    ops.conv2d_nhwc = orig

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    tot = {w: 0.0 for w in widths}
    tot_flops = 0.0
    for (x, w, cout, kh, kw, kws, out) in calls:
        N, H, W, cs = x.shape
        cin = kws.get("cin") or cs
        if kws.get("fold_w") is not None:
            W = kws["fold_w"][0]
        if out is None:
            out = kws["out2"]
        if out.dim() == 4 and kws.get("out_layout", "nhwc_bf16") == "nchw_f32":
            Ho, Wo = out.shape[2], out.shape[3]
        else:
            Ho, Wo = out.shape[1], out.shape[2]
        flops = 2.0 * N * Ho * Wo * cout * cin * kh * kw
        nbytes = x.numel() * 2 + w.numel() * 2 + out.numel() * out.element_size()
        for k in ("residual", "out2"):
            if kws.get(k) is not None:
                nbytes += kws[k].numel() * 2
        row = {"in": [N, H, W, cin], "cout": cout, "k": [kh, kw], "stride": list(kws.get("strides", (1, 1))),
               "dil": list(kws.get("dilation", (1, 1))), "gflop": flops / 1e9, "mbytes": nbytes / 1e6,
               "res": kws.get("residual") is not None, "out2": kws.get("out2") is not None,
               "out": kws.get("out_layout", "nhwc_bf16"), "t_us": {}}
        for bn in widths:
            k2 = dict(kws)
            if not kws.get("skip_out"):
                k2["out"] = out
            k2["block_n"] = bn
            try:
                for _ in range(3):
                    orig(x, w, cout, kh, kw, **k2)
                torch.cuda.synchronize()
                # the reps are replayed from a CUDA graph: no host launch overhead between the kernels
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    for _ in range(args.reps):
                        orig(x, w, cout, kh, kw, **k2)
                graph.replay()
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                graph.replay()
                e1.record()
                torch.cuda.synchronize()
                t = e0.elapsed_time(e1) / args.reps * 1e3
            except (ValueError, RuntimeError):
                t = float("nan")
            row["t_us"][bn] = t
            if t == t:
                tot[bn] += t
        tot_flops += flops
        rows.append(row)
        desc = "%dx%dx%dx%d -> %d k%dx%d s%d d%d %s%s%s" % (N, H, W, cin, cout, kh, kw, row["stride"][0], row["dil"][0],
                                                          "R" if row["res"] else "", "2" if row["out2"] else "",
                                                          "" if row["out"] == "nhwc_bf16" else " " + row["out"])
        cols = "  ".join("%4d:%7.1fus %6.1fTF %6.0fGB/s" % (bn, row["t_us"][bn], flops / row["t_us"][bn] / 1e6,
                                                        nbytes / row["t_us"][bn] / 1e3) for bn in widths)
        print("%-46s %7.2fGF %6.1fMB  %s" % (desc, flops / 1e9, nbytes / 1e6, cols))
    print("TOTAL %.1f GFLOP; " % (tot_flops / 1e9) + "  ".join("bn=%d: %.1f us" % (bn, tot[bn]) for bn in widths))
    best = sum(min(v for v in r["t_us"].values() if v == v) for r in rows)
    print("sum of per-layer best: %.1f us (%.1f TFLOP/s)" % (best, tot_flops / best / 1e6))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(rows, f)


if __name__ == "__main__":
    main()
