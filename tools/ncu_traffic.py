#!/usr/bin/env python
"""Reduce an ncu launch list (CSV from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv --log-file ...`) to per-kernel totals: launches, device time, DRAM bytes read + written.

    python tools/ncu_traffic.py gpurun_out/x/launches.csv [--steps N] [--json out.json]
--steps N divides the totals by the number of profiled steps."""
import argparse
import collections
import csv
import json
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    rows = []
    with open(args.csv, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    agg = collections.OrderedDict()
    units = {}
    for r in rows:
        name = re.sub(r"\(.*", "", r.get("Kernel Name", "?"))
        name = name.replace("xdet::(anonymous namespace)::", "").replace("void ", "")
        m, v = r.get("Metric Name"), r.get("Metric Value", "0").replace(",", "")
        try:
            v = float(v)
        except ValueError:
            continue
        u = r.get("Metric Unit", "")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        a = agg.setdefault(name, {"launches": 0, "time_us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
        if m == "gpu__time_duration.sum":
            a["launches"] += 1
            a["time_us"] += v * scale
        elif m == "dram__bytes_read.sum":
            a["dram_read"] += v * scale
        elif m == "dram__bytes_write.sum":
            a["dram_write"] += v * scale
        units[m] = u
    tot = sum(a["time_us"] for a in agg.values())
    out = {}
    print("%-52s %8s %12s %7s %14s %14s" % ("kernel", "launches", "time us/step", "share", "DRAM rd MB/step", "DRAM wr MB/step"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["time_us"]):
        s = args.steps
        print("%-52s %8.1f %12.1f %6.1f%% %14.1f %14.1f" % (k[:52], a["launches"] / s, a["time_us"] / s, 100 * a["time_us"] / max(tot, 1e-9),
                                                        a["dram_read"] / s / 1e6, a["dram_write"] / s / 1e6))
        out[k] = {"launches_per_step": a["launches"] / s, "time_us_per_step": a["time_us"] / s,
                  "share_of_kernel_time": a["time_us"] / max(tot, 1e-9),
                  "dram_bytes_per_step": (a["dram_read"] + a["dram_write"]) / s}
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
