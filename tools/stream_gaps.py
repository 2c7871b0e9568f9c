#!/usr/bin/env python
"""Where does the training step wait?  Runs eager steps under the profiler, then walks the LAST step's kernels per CUDA
stream: busy time per stream, and the idle gaps of the busiest (main) stream with the kernels either side of them."""
import argparse
import collections
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import xdet_b200  # noqa: E402,F401
from xdet_b200 import light_head_rfcn_train as lt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backbone", default="resnet50")
    ap.add_argument("--graph", action="store_true", help="profile a CUDA-graph replay instead of eager launches")
    ap.add_argument("--gap-us", type=float, default=4.0)
    args = ap.parse_args()
    params = lt.make_params(train_image_size=480, batch_size=8, backbone=args.backbone)
    trainer = lt.LightHeadTrainer(params, seed=0)
    tb = lt.synthetic_batch(params, 8, seed=3)
    for _ in range(3):
        trainer.step(*tb)
    torch.cuda.synchronize()
    run = lambda: trainer.step(*tb)  # noqa: E731
    if args.graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            trainer.step(*tb)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            trainer.step(*tb)
        run = g.replay
        run()
        torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run()
        torch.cuda.synchronize()
        run()
        torch.cuda.synchronize()
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    # the second run = the second half of the events (same count per run)
    ev = ev[len(ev) // 2:]
    t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
    print("step span %.1f us, %d launches" % (t1 - t0, len(ev)))
    per = collections.defaultdict(list)
    for e in ev:
        per[e["args"].get("stream")].append(e)
    for s, lst in sorted(per.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
        print("stream %s: %4d launches, busy %8.1f us, first at +%.0f us, last ends at +%.0f us" % (
            s, len(lst), sum(e["dur"] for e in lst), lst[0]["ts"] - t0, lst[-1]["ts"] + lst[-1]["dur"] - t0))
    main_s = max(per, key=lambda s: sum(e["dur"] for e in per[s]))
    lst = per[main_s]
    for s_, l_ in per.items():
        agg = collections.OrderedDict()
        for e in l_:
            a = agg.setdefault(e["name"][:60], [0, 0.0])
            a[0] += 1
            a[1] += e["dur"]
        print("stream %s by kernel:" % s_)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
            print("    %-62s %4d %9.1f us" % (k, v[0], v[1]))
    gaps = []
    for a, b in zip(lst, lst[1:]):
        gap = b["ts"] - (a["ts"] + a["dur"])
        if gap > 0:
            gaps.append((gap, a["name"][:48], b["name"][:48], a["ts"] + a["dur"] - t0))
    print("main stream %s: idle %.1f us in %d gaps (%.1f us in gaps <= %.0f us)" % (
        main_s, sum(g[0] for g in gaps), len(gaps), sum(g[0] for g in gaps if g[0] <= args.gap_us), args.gap_us))
    for gap, a, b, at in sorted(gaps, key=lambda g: -g[0])[:25]:
        print("  %7.1f us at +%7.0f  after %-48s before %s" % (gap, at, a, b))


if __name__ == "__main__":
    main()
