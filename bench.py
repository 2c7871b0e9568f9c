#!/usr/bin/env python
"""bench.py -- measurement contract of the driver.

    python bench.py --gpus N --steps K --warmup W            (own arm, CUDA)
    python bench.py --impl reference --gpus N --steps K ...  (reference CPU arm)

One "step" = one pass of the hot path over one batch of synthetic input.  Prints ONE JSON line.
The workload is selected with --workload (default: see WORKLOADS / DEFAULT_WORKLOAD).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as e:  # NVML missing: report, do not fail the bench
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------
# Workload: PsRoIAlign forward, top of the BASELINE config-5 sweep.
# ----------------------------------------------------------------------------------------------
class PsroiSweepTop:
    """PsRoIAlign forward, R=16384 RoIs x 7x7 bins over a 980-channel 30x30 thin feature map
    (BASELINE.json configs[4] top point; 1024 channels is rejected by the op)."""
    name = "psroi_align_fwd R=16384 C=980 7x7 map=30x30 max"
    metric, unit, dtype = "psroi_align_fwd_GBps", "GB/s", "f32 geometry + f64 blend"
    N, C, H, W, R, gw, gh = 1, 980, 30, 30, 16384, 7, 7
    cpu_sample_rois = 2048

    def algorithmic_bytes(self, R=None):
        R = self.R if R is None else R
        # features out + index out + rois in + map read once
        return 4 * self.N * R * self.C * 2 + 16 * self.N * R + 4 * self.N * self.C * self.H * self.W

    def host_inputs(self, rank):
        from tests import workloads
        x = workloads.make_map(self.N, self.C, self.H, self.W, seed=4 + 100 * rank)
        rois = workloads.make_rois(self.N, self.R, seed=5 + 100 * rank)
        return x, rois


WORKLOADS = {"psroi_sweep_top": PsroiSweepTop}
DEFAULT_WORKLOAD = "psroi_sweep_top"


def run_reference(args, wl):
    """Reference arm: the reference's own CPU implementation (oracle/_ref when it was compiled in
    the build container, else the C port) on all host cores, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import psroi
    kind = "reference" if psroi.have_ref() else "port"
    impl = "ref" if kind == "reference" else "oracle"
    cores = os.cpu_count() or 1
    x, rois = wl.host_inputs(0)
    Rs = wl.cpu_sample_rois
    rois_s = np.ascontiguousarray(rois[:, :Rs])
    for _ in range(max(1, args.warmup)):
        psroi.psroi_align_fwd(x, rois_s, wl.gw, wl.gh, "max", threads=cores, impl=impl)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        psroi.psroi_align_fwd(x, rois_s, wl.gw, wl.gh, "max", threads=cores, impl=impl)
    dt = (time.perf_counter() - t0) / args.steps
    val = wl.algorithmic_bytes(Rs) / dt / 1e9
    sample = "first %d of %d RoIs per step, all %d host threads" % (Rs, wl.R, cores)
    print(json.dumps({
        "impl": "reference", "metric": wl.metric, "value": val, "unit": wl.unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "config": {"workload": wl.name, "sample": sample},
        "cpu_baseline": {"value": val, "unit": wl.unit, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_own(args, wl):
    import torch
    import torch.distributed as dist

    import xdet_b200  # noqa: F401
    from xdet_b200 import _native, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (own arm) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _native.lib()
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    x_h, rois_h = wl.host_inputs(rank)
    x_pin = torch.from_numpy(x_h).pin_memory()
    rois_pin = torch.from_numpy(rois_h).pin_memory()
    x = x_pin.cuda()
    rois = rois_pin.cuda()
    G = wl.gw * wl.gh
    out_pin = torch.empty((wl.N, wl.R, G, wl.C // G), dtype=torch.float32).pin_memory()
    idx_pin = torch.empty((wl.N, wl.R, G, wl.C // G), dtype=torch.int32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step():
        return ops.ps_roi_align(x, rois, wl.gw, wl.gh, "max")

    # ---- device-resident throughput (value) + per-launch kernel time (roofline) -------------
    for _ in range(max(3, args.warmup)):
        flush.zero_()
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _native.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for a, b in ev:
        flush.zero_()  # evict the previous step's output / the map from L2 (not timed)
        a.record()
        step()
        b.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = _native.launch_count() - launches0
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    step_ms = sum(kernel_ms) / len(kernel_ms)

    # ---- end-to-end through the operator with HOST buffers (e2e) -----------------------------
    def e2e_step():
        xd = x_pin.cuda(non_blocking=True)
        rd = rois_pin.cuda(non_blocking=True)
        p, i = ops.ps_roi_align(xd, rd, wl.gw, wl.gh, "max")
        out_pin.copy_(p, non_blocking=True)
        idx_pin.copy_(i, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
    clocks = sampler.finish()

    # max over ranks
    t = torch.tensor([step_ms, e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, e2e_ms = float(t[0]), float(t[1])

    bytes_step = wl.algorithmic_bytes()
    value = world * bytes_step / (step_ms * 1e-3) / 1e9
    e2e_value = world * bytes_step / (e2e_ms * 1e-3) / 1e9
    achieved = bytes_step / (step_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"],
                "kernel": "psroi_fwd_planes_kernel<max>", "algorithmic_bytes_per_launch": bytes_step,
                "kernel_ms": step_ms}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import psroi
        kind = "reference" if psroi.have_ref() else "port"
        cores = os.cpu_count() or 1
        Rs = wl.cpu_sample_rois
        rs = np.ascontiguousarray(rois_h[:, :Rs])
        psroi.psroi_align_fwd(x_h, rs, wl.gw, wl.gh, "max", threads=cores, impl="ref" if kind == "reference" else "oracle")
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            psroi.psroi_align_fwd(x_h, rs, wl.gw, wl.gh, "max", threads=cores,
                                  impl="ref" if kind == "reference" else "oracle")
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": wl.algorithmic_bytes(Rs) / dt / 1e9, "unit": wl.unit, "cores": cores, "kind": kind,
               "sample": "first %d of %d RoIs, best-effort all host threads, %d reps" % (Rs, wl.R, reps)}

    if rank == 0:
        print(json.dumps({
            "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
            "config": {"workload": wl.name, "l2": "flushed between timed iterations (256 MB memset)",
                       "sharding": "independent RoI batches per rank, no collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": wl.unit, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(x_h.nbytes + rois_h.nbytes),
                    "d2h_bytes_per_step": int(out_pin.numel() * 4 + idx_pin.numel() * 4)},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "wall_s_timed_region": t_wall,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]()
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_own(args, wl)


if __name__ == "__main__":
    main()
