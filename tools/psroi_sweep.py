#!/usr/bin/env python
"""BASELINE config 5: PsRoIAlign forward throughput sweep (128..16384 RoIs x {7x7, 15x15} bins,
maps 30x30 / 50x50) as GB/s of algorithmic bytes against the measured HBM peak.  GPU only.
    python tools/psroi_sweep.py [--out gpurun_out/psroi_sweep.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import xdet_b200  # noqa: F401,E402
from tests import workloads  # noqa: E402
from xdet_b200 import ops  # noqa: E402


def time_op(fn, flush, iters=10):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2], ms[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "psroi_sweep.json"))
    ap.add_argument("--variants", default="select,planes,gather")
    ap.add_argument("--quick", action="store_true", help="R in {1024, 4096, 16384} only")
    ap.add_argument("--methods", default="max")
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for (C, g, hw) in [(980, 7, 30), (900, 15, 30), (490, 7, 30), (980, 7, 50), (900, 15, 50)]:
        x = torch.from_numpy(workloads.make_map(1, C, hw, hw, seed=4)).cuda()
        for R in ([1024, 4096, 16384] if args.quick else [128, 256, 512, 1024, 2048, 4096, 8192, 16384]):
            rois = torch.from_numpy(workloads.make_rois(1, R, seed=5)).cuda()
            nbytes = 8 * R * C + 16 * R + 4 * C * hw * hw
            for method in args.methods.split(","):
                for variant in args.variants.split(","):
                    try:
                        med, best = time_op(lambda: ops.ps_roi_align(x, rois, g, g, method, variant=variant), flush)
                    except ValueError as e:
                        print("skip", C, g, hw, R, variant, e)
                        continue
                    row = {"C": C, "bins": g, "map": hw, "R": R, "method": method, "variant": variant, "us_median": med * 1e3,
                           "us_best": best * 1e3, "alg_bytes": nbytes, "GBps": nbytes / med / 1e6,
                           "frac_of_measured_hbm": nbytes / med / 1e6 / peaks["hbm_gbs"]}
                    rows.append(row)
                    print("C=%4d %2dx%-2d map=%2d R=%5d %-4s %-6s %9.1f us  %8.1f GB/s  %.3f of HBM" %
                          (C, g, g, hw, R, method, variant, row["us_median"], row["GBps"], row["frac_of_measured_hbm"]),
                          flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"peak_hbm_gbs": peaks["hbm_gbs"], "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
