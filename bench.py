#!/usr/bin/env python
"""bench.py -- measurement contract of the driver.

    python bench.py --gpus N --steps K --warmup W            (own arm, CUDA)
    python bench.py --impl reference --gpus N --steps K ...  (reference-semantics CPU arm)

One "step" = one pass of the hot path over one batch of synthetic input.  Prints ONE JSON line.

Default workload (BASELINE.json configs[1]): Light-Head R-CNN, ResNet-50 backbone, inference, batch 8 per
GPU, 480x480 synthetic images; metric = images/s.  `--workload psroi_sweep_top` measures the PsRoIAlign
operator alone (BASELINE configs[4] top point) as GB/s against the HBM roofline.
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
                time.sleep(0.01)
        except Exception as e:  # NVML missing: report, do not fail the bench
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def ncu_traffic(precision, kernel_prefix):
    """DRAM bytes (read + written) per step of the kernels whose name starts with ``kernel_prefix``, from the committed
    ncu launch list of this workload (profiles/ncu_traffic_<precision>_r2.json, made by tools/ncu_step.py +
    tools/ncu_traffic.py: cold-cache, serialised launches).  None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic_%s_r2.json" % precision)
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    hits = [v["dram_bytes_per_step"] for k, v in d.items() if k.split("::")[-1].startswith(kernel_prefix)]
    return float(sum(hits)) if hits else None


def dist_env():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))



def event_pair_overhead(torch, n=200):
    """Median CUDA-event interval around NOTHING, with the records pre-queued behind a sleep kernel: the share of a
    per-launch event interval that is event/dispatch latency rather than kernel execution."""
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda._sleep(20_000_000)
    for a, b in pairs:
        a.record()
        b.record()
    torch.cuda.synchronize()
    d = sorted(a.elapsed_time(b) for a, b in pairs)
    return d[len(d) // 2]


def dist_init(local):
    """NCCL process group for the ranks torchrun started.  NCCL_DEBUG is left as the caller set it (the driver reads
    the rank lines of NCCL_DEBUG=INFO); NCCL's log goes to stderr so that stdout carries exactly ONE JSON line."""
    import torch
    import torch.distributed as dist
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def dist_finish(world):
    """Leave without tearing NCCL down under live CUDA graphs (destroy_process_group can hang there): flush, meet at a
    barrier, and exit the process directly."""
    import torch
    import torch.distributed as dist
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


def own_context():
    """Device, process group and measured peaks for the own arm (one process per GPU)."""
    import torch
    import torch.distributed as dist

    import xdet_b200  # noqa: F401
    from xdet_b200 import _native

    world, rank, local = dist_env()
    assert torch.cuda.is_available(), "bench.py (own arm) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist_init(local)
    _native.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    return {"world": world, "rank": rank, "local": local, "barrier": barrier, "peaks": load_peaks()}


# ==============================================================================================
# Workload 1 (default): Light-Head R-CNN ResNet-50 inference, batch 8/GPU, 480x480
# ==============================================================================================
class LightHeadResnet50:
    name = "Light-Head R-CNN ResNet-50 inference, batch=8 per GPU, 480x480 synthetic"
    metric, unit, dtype = "images_per_sec_480x480", "images/s", "bf16"
    batch, size, backbone = 8, 480, "resnet50"

    def images(self, rank):
        rng = np.random.default_rng(1 + 1000 * rank)  # U(-1,1): img*2 - mean/127.5 (common_preprocessing.py:391-392)
        return (rng.random((self.batch, 3, self.size, self.size), dtype=np.float32) * 2 - 1).astype(np.float32)

    # precision of the headline number: the one that holds north_star's 1e-4 against the reference's fp32 graph
    precision = "f16x2"
    DTYPES = {"f16x2": "fp32-accurate: 2 x fp16 planes per operand on tcgen05 (kind::f16), fp32 accumulate, chunked "
                       "round-to-nearest flushes; fp32 activations",
              "bf16": "bf16 operands and activations, fp32 accumulate"}

    def measure(self, args, precision, ctx, state_dict=None):
        """One precision of the inference path: device-resident img/s (CUDA-graph replay, CUDA events), end to end
        with HOST buffers, live per-launch timing of the convolution kernel.  Returns a dict (+ the model)."""
        import torch
        import torch.distributed as dist

        from xdet_b200 import _native
        from xdet_b200 import light_head_rfcn_eval as lh
        from xdet_b200.ops import conv as conv_ops
        world, rank, local, barrier, peaks = ctx["world"], ctx["rank"], ctx["local"], ctx["barrier"], ctx["peaks"]

        params = lh.make_params(train_image_size=self.size, backbone=self.backbone, rpn_min_size=16.0 / self.size,
                                precision=precision)
        model = lh.LightHeadRFCN(params, seed=0, state_dict=state_dict)
        conv_ops.AUTOTUNE = not args.no_autotune  # one-time tile-shape tuning per layer shape during the eager pass
        imgs_h = torch.from_numpy(self.images(rank)).pin_memory()
        imgs_d = imgs_h.cuda()
        # the step ends with the per-class detections (bboxes_eval): what the reference's eval loop consumes
        ncls, ndet = params["num_classes"] - 1, params["nms_topk"]
        probs_h = torch.empty((self.batch, ncls, ndet), dtype=torch.float32).pin_memory()
        boxes_h = torch.empty((self.batch, ncls, ndet, 4), dtype=torch.float32).pin_memory()
        run = lambda x: model(x, detections=True)  # noqa: E731

        # ---- build: one eager pass creates the variables / packed weights, then the whole forward is
        # captured into a CUDA graph (launch-bound otherwise: ~80 kernels per step) ------------------
        static_in = torch.empty_like(imgs_d)
        static_in.copy_(imgs_d)
        out = run(static_in)
        torch.cuda.synchronize()
        graph = None
        if not args.no_graph:
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(2):
                        out = run(static_in)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = run(static_in)
                torch.cuda.synchronize()
            except Exception as e:  # report and fall back to eager launches
                sys.stderr.write("CUDA graph capture failed (%s: %s); timing eager launches\n" % (type(e).__name__, e))
                graph = None
                out = run(static_in)
                torch.cuda.synchronize()

        def step():
            if graph is not None:
                graph.replay()
                return out
            return run(static_in)

        l0 = _native.launch_count()
        run(static_in)
        torch.cuda.synchronize()
        launches_per_step = _native.launch_count() - l0

        # ---- device-resident throughput (value) ---------------------------------------------------
        for _ in range(max(3, args.warmup)):
            step()
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        step_ms = e0.elapsed_time(e1) / args.steps

        # ---- end to end with HOST buffers: H2D of the images, D2H of the detections, every step -----------
        # Double-buffered: the pinned images of step i+1 travel host->device on a copy stream while step i computes,
        # and the detections of step i travel device->host while step i+1 computes (what a serving loop does).  Every
        # step's copies are inside the timed region; the host waits for the detections of step i-1 before it
        # enqueues step i+1 (bounded queue), and for everything at the end.
        stage_in = [torch.empty_like(imgs_d) for _ in range(2)]
        det_s_d = [torch.empty((self.batch, ncls, ndet), dtype=torch.float32, device="cuda") for _ in range(2)]
        det_b_d = [torch.empty((self.batch, ncls, ndet, 4), dtype=torch.float32, device="cuda") for _ in range(2)]
        probs_hh = [probs_h, torch.empty_like(probs_h).pin_memory()]
        boxes_hh = [boxes_h, torch.empty_like(boxes_h).pin_memory()]
        cur = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        ev_in = [torch.cuda.Event() for _ in range(2)]     # images of slot b are on the device
        ev_free = [torch.cuda.Event() for _ in range(2)]   # slot b's staging buffer has been consumed
        ev_det = [torch.cuda.Event() for _ in range(2)]    # detections of slot b are in their device buffer
        ev_done = [torch.cuda.Event() for _ in range(2)]   # detections of slot b are on the host

        def e2e_run(n):
            with torch.cuda.stream(s_in):
                stage_in[0].copy_(imgs_h, non_blocking=True)
                ev_in[0].record(s_in)
            for i in range(n):
                b = i & 1
                if i + 1 < n:  # next step's images: host -> device, behind the compute of this step
                    with torch.cuda.stream(s_in):
                        if i >= 1:
                            s_in.wait_event(ev_free[b ^ 1])
                        stage_in[b ^ 1].copy_(imgs_h, non_blocking=True)
                        ev_in[b ^ 1].record(s_in)
                cur.wait_event(ev_in[b])
                static_in.copy_(stage_in[b], non_blocking=True)  # device->device, 7 us
                ev_free[b].record(cur)
                o = step()
                if i >= 2:
                    cur.wait_event(ev_done[b])  # the D2H that last read these device buffers has finished
                det_s_d[b].copy_(o["det_scores"], non_blocking=True)
                det_b_d[b].copy_(o["det_bboxes"], non_blocking=True)
                ev_det[b].record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_det[b])
                    probs_hh[b].copy_(det_s_d[b], non_blocking=True)
                    boxes_hh[b].copy_(det_b_d[b], non_blocking=True)
                    ev_done[b].record(s_out)
                if i >= 1:
                    ev_done[b ^ 1].synchronize()  # the host consumes the detections of step i-1 here
            torch.cuda.synchronize()

        e2e_run(3)
        barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        barrier()
        e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
        clocks = sampler.finish()

        # ---- dominant kernel (the convolution): live per-launch CUDA-event timing, eager pass ------
        # (the GPU first spins in a ~20 ms sleep kernel so that the host runs ahead and every launch is already queued
        # when its start event is reached: the intervals are kernel execution, not launch latency)
        conv_ops.PROFILE = []
        for _ in range(3):
            conv_ops.PROFILE.clear()
            torch.cuda._sleep(40_000_000)
            run(static_in)
            torch.cuda.synchronize()
        prof = conv_ops.PROFILE
        conv_ops.PROFILE = None
        ev_over = event_pair_overhead(torch)
        conv_ms_raw = sum(a.elapsed_time(b) for a, b, _, _ in prof)
        conv_ms = sum(max(a.elapsed_time(b) - ev_over, 0.0) for a, b, _, _ in prof)
        conv_flops = sum(f for _, _, f, _ in prof)
        n_conv = len(prof)

        t = torch.tensor([step_ms, e2e_ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, e2e_ms = float(t[0]), float(t[1])
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12
        # the convolutions were event-timed one by one at boost clocks: the burst cuBLAS figure is the honest peak
        peak = peaks["bf16_tflops"]
        products = 3 if precision == "f16x2" else 1
        kname = "conv_gemm_f16x2_kernel" if precision == "f16x2" else "conv_gemm_kernel"
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": ncu_traffic(precision, kname) if (self.size == 480 and self.backbone == "resnet50") else None,
                    "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of all launches of the kernel in one "
                                    "step, from the committed ncu launch list (profiles/ncu_traffic_%s_r2.txt)" % precision,
                    "peak_source": peaks["source"] + " (burst cuBLAS bf16; fp16 runs at the same rate)",
                    "kernel": "%s (tcgen05 implicit GEMM), %d launches per step" % (kname, n_conv),
                    "algorithmic_flops_per_step": conv_flops, "kernel_ms_per_step": conv_ms,
                    "kernel_ms_per_step_raw": conv_ms_raw, "event_pair_overhead_ms": ev_over,
                    "kernel_share_of_step": conv_ms / step_ms,
                    "tensor_products_per_algorithmic_product": products,
                    "tensor_pipe_tflops": achieved * products, "tensor_pipe_frac": achieved * products / peak,
                    "note": "achieved = ALGORITHMIC flops (2*MACs of every conv/dense layer, bias/BN/ReLU excluded) / "
                            "kernel time; per-launch CUDA-event times from an eager pass with the launches pre-queued "
                            "behind a sleep kernel, minus the measured empty event-pair interval per launch (raw sum "
                            "kept beside it)" + ("; the fp32-accurate kernel issues 3 tensor-core products per algorithmic "
                                                 "product, so its ceiling is peak/3" if products == 3 else "")}
        return {"precision": precision, "dtype": self.DTYPES[precision], "value": world * self.batch / (step_ms * 1e-3),
                "ms_per_step": step_ms,
                "e2e": {"value": world * self.batch / (e2e_ms * 1e-3), "unit": self.unit, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(imgs_h.numel() * 4),
                        "d2h_bytes_per_step": int(probs_h.numel() * 4 + boxes_h.numel() * 4)},
                "gpu_launches_per_step": int(launches_per_step), "cuda_graph": graph is not None, "clocks": clocks,
                "roofline": roofline, "_model": model, "_params": params}

    def parity(self, models, cpu_out, img, keys):
        """Deltas of each precision against the CPU oracle's outputs on the image the cpu_baseline leg just ran (same
        variables, same shuffle keys).  Proposals are paired by box (an NMS near-tie flip shifts row positions)."""
        import torch
        res = {}
        pb = np.asarray(cpu_out["proposals_bboxes"], np.float32)[0]
        for name, model in models.items():
            out = model(torch.from_numpy(img).cuda(), shuffle_keys=torch.from_numpy(keys).cuda())
            torch.cuda.synchronize()
            pa = out["proposals_bboxes"].cpu().numpy()[0]
            d = np.abs(pa[:, None, :] - pb[None, :, :]).max(axis=-1)
            j = d.argmin(axis=1)
            ok = d[np.arange(pa.shape[0]), j] < 1e-4
            rec = {"rpn_object_score_max_abs": float(np.abs(out["rpn_object_score"].cpu().numpy() - cpu_out["rpn_object_score"]).max()),
                   "large_sep_feature_max_rel": float(np.abs(out["large_sep_feature"].cpu().numpy() - cpu_out["large_sep_feature"]).max()
                                                      / np.abs(cpu_out["large_sep_feature"]).max()),
                   "proposals_paired": int(ok.sum()), "proposals": int(ok.size)}
            if ok.any():
                for k, kk in (("head_cls_score", "scores_max_abs"), ("bboxes_predict", "boxes_max_abs")):
                    a = out[k].float().cpu().numpy().reshape(cpu_out[k].shape)
                    rec[kk] = float(np.abs(a[np.nonzero(ok)[0]] - cpu_out[k][j[ok]]).max())
            res[name] = rec
        return res

    def run_own(self, args):
        ctx = own_context()
        world, rank = ctx["world"], ctx["rank"]
        main = self.measure(args, self.precision, ctx)
        model, params = main.pop("_model"), main.pop("_params")
        subs = {}
        if not args.headline_only:
            # the same network with bf16 operands (the fast mode: ~1e-2 per stage against the fp32 graph)
            fast = self.measure(args, "bf16", ctx, state_dict=model.store.state_dict())
            fast_model = fast.pop("_model")
            fast.pop("_params")
            subs["bf16_mode"] = fast
            if self.size == 480 and self.backbone == "resnet50":
                subs["psroi"] = PsroiSweepTop().measure(args, ctx)
                subs["train"] = LightHeadResnet50Train().measure(args, ctx)

        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            cpu, cpu_out, img, keys = cpu_network_baseline(self, params, model.store.state_dict(), want_outputs=True)
            models = {self.precision: model}
            if "bf16_mode" in subs:
                models["bf16"] = fast_model
            par = self.parity(models, cpu_out, img, keys)
            main_par = par.pop(self.precision)
            if "bf16_mode" in subs:
                subs["bf16_mode"]["parity_vs_cpu_oracle"] = par["bf16"]
        else:
            main_par = None

        if rank == 0:
            line = {
                "metric": self.metric, "value": main["value"], "unit": self.unit, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": main["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": main["dtype"], "data": "synthetic",
                "config": {"workload": self.name, "global_batch": world * self.batch, "precision": main["precision"],
                           "l2": "working set per step (activations > 1 GB) exceeds the 126 MB L2",
                           "sharding": "images partitioned across ranks, no collective (inference)",
                           "cuda_graph": main["cuda_graph"], "conv_autotune": not args.no_autotune,
                           "proposals": "pre_nms_top_n=5000 post_nms_top_n=1000 nms=0.7 (eval flags)",
                           "postprocess": "bboxes_eval on the GPU inside the step: 20 classes x top-400 x NMS 0.3 -> 200"},
                "clocks": main["clocks"], "e2e": main["e2e"],
                "gpu_launches": int(main["gpu_launches_per_step"] * args.steps),
                "gpu_launches_per_step": main["gpu_launches_per_step"],
                "roofline": main["roofline"], "cpu_baseline": cpu,
                "parity_vs_cpu_oracle": main_par,
            }
            line.update(subs)
            print(json.dumps(line))
        dist_finish(world)

    def run_reference(self, args):
        world, rank, _ = dist_env()
        if rank != 0:
            return
        import torch

        # eval flags of the reference (light_head_rfcn_eval.py:95-115); variables are created on demand by the
        # oracle with the reference's shapes (random: no checkpoint exists offline)
        params = {"model_scope": "xception_lighthead", "num_classes": 21, "rpn_pre_nms_top_n": 5000,
                  "rpn_post_nms_top_n": 1000, "rpn_nms_thres": 0.7, "rpn_min_size": 16.0 / self.size,
                  "backbone": self.backbone}
        sd = {}
        base = cpu_network_baseline(self, params, sd, reps=max(1, args.steps), warmup=1 if args.warmup else 0)
        print(json.dumps({
            "impl": "reference", "metric": self.metric, "value": base["value"], "unit": self.unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / base["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": self.name, "sample": base["sample"]},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": self.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        del torch


def cpu_network_baseline(wl, params, sd, reps=None, warmup=1, want_outputs=False, budget_s=10.0):
    """The reference-semantics CPU path (oracle/net.py, PyTorch CPU fp32 -- TF1 itself is not installable) on a
    bounded sample: ONE image of the batch per repetition, all host threads; ``reps`` repetitions, or (None) as many as
    fit ~``budget_s`` seconds of CPU work (at least 3, at most 40) after one warm-up."""
    import torch

    from oracle import net as onet
    from oracle import proposals as op
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fmap = wl.size // 16
    anchors = op.layer_anchors((wl.size, wl.size), (fmap, fmap), [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], [0.1],
                               [1., 2., .5], 16)
    img = wl.images(0)[:1]
    sd = dict(sd)
    onet.model(img[:, :, :64, :64], sd, dict(params, rpn_pre_nms_top_n=50, rpn_post_nms_top_n=10),
               op.layer_anchors((64, 64), (4, 4), [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], [0.1], [1., 2., .5], 16),
               create_seed=0)  # creates any missing variable (tiny input), untimed
    from oracle import detections as od

    keys = np.random.default_rng(7).random((1, params["rpn_post_nms_top_n"]), dtype=np.float32)
    last = {}

    def one():
        o = onet.model(img, sd, params, anchors, shuffle_keys=keys)
        last["o"] = o
        od.bboxes_eval_select(o["head_cls_score"], o["bboxes_predict"], np.array([0, 0, 1, 1], np.float32),
                              (wl.size, wl.size), params["num_classes"], train_image_size=wl.size)

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    done = 0
    while True:
        one()
        done += 1
        if reps is not None:
            if done >= reps:
                break
        elif done >= 40 or (done >= 3 and time.perf_counter() - t0 >= budget_s):
            break
    reps = done
    dt = (time.perf_counter() - t0) / reps
    base = {"value": 1.0 / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "1 image of the batch per step (whole graph incl. proposals/NMS, PsRoIAlign C oracle, per-class NMS), "
                      "PyTorch-CPU fp32 restatement, %d threads, %d rep(s)" % (cores, reps)}
    return (base, last["o"], img, keys) if want_outputs else base


# ==============================================================================================
# Workload 2: PsRoIAlign forward, top of the BASELINE config-5 sweep
# ==============================================================================================
class PsroiSweepTop:
    """PsRoIAlign forward, R=16384 RoIs x 7x7 bins over a 980-channel 30x30 thin feature map
    (BASELINE.json configs[4] top point; 1024 channels is rejected by the op)."""
    name = "psroi_align_fwd R=16384 C=980 7x7 map=30x30 max"
    metric, unit, dtype = "psroi_align_fwd_GBps", "GB/s", "f32 geometry + f64 blend"
    N, C, H, W, R, gw, gh = 1, 980, 30, 30, 16384, 7, 7
    cpu_sample_rois = 2048

    def algorithmic_bytes(self, R=None):
        R = self.R if R is None else R
        return 4 * self.N * R * self.C * 2 + 16 * self.N * R + 4 * self.N * self.C * self.H * self.W

    def host_inputs(self, rank):
        from tests import workloads
        return (workloads.make_map(self.N, self.C, self.H, self.W, seed=4 + 100 * rank),
                workloads.make_rois(self.N, self.R, seed=5 + 100 * rank))

    def cpu(self, x_h, rois_h, reps):
        from oracle import psroi
        kind = "reference" if psroi.have_ref() else "port"
        impl = "ref" if kind == "reference" else "oracle"
        cores = os.cpu_count() or 1
        rs = np.ascontiguousarray(rois_h[:, :self.cpu_sample_rois])
        psroi.psroi_align_fwd(x_h, rs, self.gw, self.gh, "max", threads=cores, impl=impl)
        t0 = time.perf_counter()
        for _ in range(reps):
            psroi.psroi_align_fwd(x_h, rs, self.gw, self.gh, "max", threads=cores, impl=impl)
        dt = (time.perf_counter() - t0) / reps
        return {"value": self.algorithmic_bytes(self.cpu_sample_rois) / dt / 1e9, "unit": self.unit, "cores": cores,
                "kind": kind, "sample": "first %d of %d RoIs per step, %d host threads, %d reps" %
                (self.cpu_sample_rois, self.R, cores, reps)}, dt

    def run_reference(self, args):
        if dist_env()[1] != 0:
            return
        x, rois = self.host_inputs(0)
        base, dt = self.cpu(x, rois, max(1, args.steps))
        print(json.dumps({
            "impl": "reference", "metric": self.metric, "value": base["value"], "unit": self.unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": self.dtype, "data": "synthetic",
            "config": {"workload": self.name, "sample": base["sample"]}, "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": self.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))

    def measure(self, args, ctx):
        import torch
        import torch.distributed as dist

        from xdet_b200 import _native, ops
        world, rank, local, barrier, peaks = ctx["world"], ctx["rank"], ctx["local"], ctx["barrier"], ctx["peaks"]
        steps = max(5, min(args.steps, 20))

        x_h, rois_h = self.host_inputs(rank)
        x_pin, rois_pin = torch.from_numpy(x_h).pin_memory(), torch.from_numpy(rois_h).pin_memory()
        x, rois = x_pin.cuda(), rois_pin.cuda()
        G = self.gw * self.gh
        out_pin = torch.empty((self.N, self.R, G, self.C // G), dtype=torch.float32).pin_memory()
        idx_pin = torch.empty((self.N, self.R, G, self.C // G), dtype=torch.int32).pin_memory()
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

        def step():
            return ops.ps_roi_align(x, rois, self.gw, self.gh, "max")

        for _ in range(max(3, args.warmup)):
            flush.zero_()
            step()
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        launches0 = _native.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for a, b in ev:
            flush.zero_()  # evict the previous step's output / the map from L2 (not timed)
            a.record()
            step()
            b.record()
        barrier()
        launches = _native.launch_count() - launches0
        step_ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)

        def e2e_step():
            xd = x_pin.cuda(non_blocking=True)
            rd = rois_pin.cuda(non_blocking=True)
            p, i = ops.ps_roi_align(xd, rd, self.gw, self.gh, "max")
            out_pin.copy_(p, non_blocking=True)
            idx_pin.copy_(i, non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        barrier()
        e2e_ms = (time.perf_counter() - t0) / steps * 1e3
        clocks = sampler.finish()
        t = torch.tensor([step_ms, e2e_ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, e2e_ms = float(t[0]), float(t[1])
        nbytes = self.algorithmic_bytes()
        achieved = nbytes / (step_ms * 1e-3) / 1e9
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            cpu, _ = self.cpu(x_h, rois_h, 3)
        del flush
        return {"metric": self.metric, "workload": self.name, "value": world * achieved, "unit": self.unit,
                "steps": steps, "ms_per_step": step_ms, "dtype": self.dtype,
                "l2": "flushed between timed iterations (256 MB memset)", "clocks": clocks,
                "e2e": {"value": world * nbytes / (e2e_ms * 1e-3) / 1e9, "unit": self.unit, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(x_h.nbytes + rois_h.nbytes),
                        "d2h_bytes_per_step": int(out_pin.numel() * 4 + idx_pin.numel() * 4)},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"],
                             "kernel": "psroi_prep_geom_kernel + psroi_prep_perm_kernel + psroi_fwd_select_kernel",
                             "algorithmic_bytes_per_launch": nbytes, "kernel_ms": step_ms},
                "cpu_baseline": cpu}

    def run_own(self, args):
        ctx = own_context()
        m = self.measure(args, ctx)
        if ctx["rank"] == 0:
            print(json.dumps({
                "metric": self.metric, "value": m["value"], "unit": self.unit, "n_gpus": ctx["world"],
                "steps": m["steps"], "warmup": max(3, args.warmup), "ms_per_step": m["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": self.dtype, "data": "synthetic",
                "config": {"workload": self.name, "l2": m["l2"],
                           "sharding": "independent RoI batches per rank, no collective"},
                "clocks": m["clocks"], "e2e": m["e2e"], "gpu_launches": m["gpu_launches"], "roofline": m["roofline"],
                "cpu_baseline": m["cpu_baseline"],
            }))
        dist_finish(ctx["world"])


class LightHeadResnet50Train:
    """BASELINE.json configs[3]: ResNet-50 light-head TRAINING step (RPN + head losses, OHEM, backward, momentum SGD,
    gradient all-reduce), 8 images of 480x480 per GPU (batch 64 on 8 GPUs)."""
    name = "Light-Head R-CNN ResNet-50 training step, batch=8 per GPU, 480x480 synthetic"
    metric, unit, dtype = "train_images_per_sec_480x480", "images/s", "bf16"
    batch, size = 8, 480

    backbone = "resnet50"

    def measure(self, args, ctx):
        import torch
        import torch.distributed as dist

        from xdet_b200 import _native
        from xdet_b200 import light_head_rfcn_train as lt
        from xdet_b200.ops import conv as conv_ops
        world, rank, local, barrier, peaks = ctx["world"], ctx["rank"], ctx["local"], ctx["barrier"], ctx["peaks"]

        params = lt.make_params(train_image_size=self.size, batch_size=self.batch, backbone=self.backbone)
        trainer = lt.LightHeadTrainer(params, seed=0)  # same seed on every rank: identical initial weights
        images, gt, gl, keys = lt.synthetic_batch(params, self.batch, seed=3 + 1000 * rank, device="cpu")
        host = [images.pin_memory(), gt.pin_memory(), gl.pin_memory()] + [keys[k].pin_memory() for k in sorted(keys)]
        dev = [t.cuda() for t in host]
        losses_h = torch.empty(3, dtype=torch.float32).pin_memory()

        def step(tensors):
            kd = dict(zip(sorted(keys), tensors[3:]))
            return trainer.step(tensors[0], tensors[1], tensors[2], kd)

        eager_step = step
        eager_step(dev)  # first pass: packs, one-time tile-shape tuning
        torch.cuda.synchronize()
        l0 = _native.launch_count()
        eager_step(dev)
        torch.cuda.synchronize()
        launches_per_step = _native.launch_count() - l0
        # the whole step (forward, backward, all-reduce, optimizer) replays from ONE CUDA graph: ~650 launches per
        # step are host-bound otherwise.  Static input buffers; the learning rate is constant until step 60000.
        graph, static, gout = None, [torch.empty_like(t) for t in dev], None
        for s_, d_ in zip(static, dev):
            s_.copy_(d_)
        if not args.no_graph:
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(2):
                        eager_step(static)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    gout = eager_step(static)
                torch.cuda.synchronize()
            except Exception as e:
                import traceback
                sys.stderr.write("CUDA graph capture failed (%s: %s); timing eager launches\n%s" %
                                 (type(e).__name__, e, traceback.format_exc()))
                graph = None
                torch.cuda.synchronize()

        def step(tensors):
            if tensors is not static:  # (pinned host tensors of the end-to-end leg: H2D into the step's input buffers)
                for s_, d_ in zip(static, tensors):
                    s_.copy_(d_, non_blocking=True)
            if graph is None:
                return eager_step(static)
            graph.replay()
            return gout

        dev = static
        for _ in range(max(3, args.warmup)):
            step(dev)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            out = step(dev)
        e1.record()
        barrier()
        step_ms = e0.elapsed_time(e1) / args.steps

        # ---- end to end with HOST buffers, every step: H2D of the batch (images, boxes, labels, shuffle keys), D2H of
        # the three losses.  Double-buffered like the inference arm: step i+1's batch travels host->device on a copy
        # stream while step i computes; the host reads the losses of step i-1 before it enqueues step i+1.
        stage = [[torch.empty_like(t) for t in static] for _ in range(2)]
        loss_d = [torch.empty(3, dtype=torch.float32, device="cuda") for _ in range(2)]
        loss_hh = [losses_h, torch.empty_like(losses_h).pin_memory()]
        cur = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        ev_loss = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]

        def upload(b):
            for s_, h_ in zip(stage[b], host):
                s_.copy_(h_, non_blocking=True)

        def e2e_run(n):
            with torch.cuda.stream(s_in):
                upload(0)
                ev_in[0].record(s_in)
            for i in range(n):
                b = i & 1
                if i + 1 < n:
                    with torch.cuda.stream(s_in):
                        if i >= 1:
                            s_in.wait_event(ev_free[b ^ 1])
                        upload(b ^ 1)
                        ev_in[b ^ 1].record(s_in)
                cur.wait_event(ev_in[b])
                for s_, d_ in zip(static, stage[b]):
                    s_.copy_(d_, non_blocking=True)  # device -> device into the step's (graph-captured) input buffers
                ev_free[b].record(cur)
                o = step(static)
                if i >= 2:
                    cur.wait_event(ev_done[b])
                torch.stack([o["rpn_cross_entropy_loss"], o["rpn_location_loss"], o["head_loss"]], out=loss_d[b])
                ev_loss[b].record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_loss[b])
                    loss_hh[b].copy_(loss_d[b], non_blocking=True)
                    ev_done[b].record(s_out)
                if i >= 1:
                    ev_done[b ^ 1].synchronize()  # the host has the losses of step i-1
            torch.cuda.synchronize()

        e2e_run(3)
        barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        barrier()
        e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
        clocks = sampler.finish()

        conv_ops.PROFILE = []
        for _ in range(2):
            conv_ops.PROFILE.clear()
            torch.cuda._sleep(100_000_000)  # host runs ahead of the GPU: intervals = kernel execution
            eager_step(static)
            torch.cuda.synchronize()
        prof = conv_ops.PROFILE
        conv_ops.PROFILE = None
        ev_over = event_pair_overhead(torch)
        conv_ms = sum(max(a.elapsed_time(b) - ev_over, 0.0) for a, b, _, _ in prof)
        conv_flops = sum(f for _, _, f, _ in prof)

        t = torch.tensor([step_ms, e2e_ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, e2e_ms = float(t[0]), float(t[1])
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        h2d = int(sum(t_.numel() * t_.element_size() for t_ in host))
        comm = getattr(trainer, "comm_info", lambda: None)()
        return {"metric": self.metric, "workload": self.name, "value": world * self.batch / (step_ms * 1e-3),
                "unit": self.unit, "ms_per_step": step_ms, "dtype": self.dtype, "global_batch": world * self.batch,
                "l2": "working set per step (saved activations ~2 GB) exceeds the 126 MB L2",
                "sharding": "images partitioned across ranks; NCCL all-reduce (sum) of the flat fp32 gradient buffer "
                            "(%.0f MB) per step" % (trainer.grads.numel() * 4 / 1e6),
                "allreduce_bytes_per_step": int(trainer.grads.numel() * 4) if world > 1 else 0, "allreduce": comm,
                "cuda_graph": graph is not None,
                "flags": "rpn 10000->1800 @0.7, 64 RoIs/img @25% fg, OHEM 32, 256 RPN samples/img, momentum 0.9",
                "clocks": clocks,
                "e2e": {"value": world * self.batch / (e2e_ms * 1e-3), "unit": self.unit, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12},
                "gpu_launches_per_step": int(launches_per_step),
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": achieved / peak,
                             "traffic": ncu_traffic("train" if self.backbone == "resnet50" else "xception_train", "conv_"),
                             "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of the convolution kernels "
                                             "(forward, input and weight gradients) in one step, from the committed ncu "
                                             "launch list (profiles/ncu_traffic_train_r2.txt)",
                             "peak_source": peaks["source"] + " (sustained cuBLAS bf16)",
                             "kernel": "conv_gemm_kernel (forward + input gradients) and conv_wgrad_kernel, %d launches per step" % len(prof),
                             "algorithmic_flops_per_step": conv_flops, "kernel_ms_per_step": conv_ms,
                             "note": "per-launch CUDA-event times of an eager step"},
                "losses": {k: float(out[k]) for k in ("rpn_cross_entropy_loss", "rpn_location_loss", "head_loss")}}

    def run_own(self, args):
        ctx = own_context()
        m = self.measure(args, ctx)
        if ctx["rank"] == 0:
            print(json.dumps({
                "metric": self.metric, "value": m["value"], "unit": self.unit, "n_gpus": ctx["world"],
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": m["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": self.dtype, "data": "synthetic",
                "config": {"workload": self.name, "global_batch": m["global_batch"], "l2": m["l2"], "sharding": m["sharding"],
                           "cuda_graph": m["cuda_graph"], "flags": m["flags"]},
                "clocks": m["clocks"], "e2e": m["e2e"],
                "gpu_launches": int(m["gpu_launches_per_step"] * args.steps),
                "gpu_launches_per_step": m["gpu_launches_per_step"], "roofline": m["roofline"], "cpu_baseline": None,
                "allreduce": m["allreduce"], "allreduce_bytes_per_step": m["allreduce_bytes_per_step"],
                "losses": m["losses"],
            }))
        dist_finish(ctx["world"])

    def run_reference(self, args):
        world, rank, _ = dist_env()
        if rank != 0:
            return
        print(json.dumps({"impl": "reference", "metric": self.metric, "unavailable":
                          "the CPU autograd restatement of the training step (oracle/net_train.py) needs the GPU step's "
                          "discrete selections injected; it is a parity checker, not a timed arm"}))


class LightHeadXceptionTrain(LightHeadResnet50Train):
    """The reference's own training configuration (light_head_rfcn_train.py:289: XceptionBody), 8 images of 480x480 per
    GPU."""
    name = "Light-Head R-CNN Xception training step, batch=8 per GPU, 480x480 synthetic"
    backbone = "xception"


class LightHeadXception800(LightHeadResnet50):
    """BASELINE.json configs[2]: the reference's own backbone (XceptionBody), batch 32, 800x800."""
    name = "Light-Head R-CNN Xception inference, batch=32 per GPU, 800x800 synthetic"
    metric = "images_per_sec_800x800"
    batch, size, backbone = 32, 800, "xception"


class LightHeadXception480(LightHeadResnet50):
    name = "Light-Head R-CNN Xception inference, batch=8 per GPU, 480x480 synthetic"
    batch, size, backbone = 8, 480, "xception"


WORKLOADS = {"lighthead_resnet50": LightHeadResnet50, "lighthead_xception_800": LightHeadXception800,
             "lighthead_xception_480": LightHeadXception480, "lighthead_resnet50_train": LightHeadResnet50Train,
             "lighthead_xception_train": LightHeadXceptionTrain,
             "psroi_sweep_top": PsroiSweepTop}
DEFAULT_WORKLOAD = "lighthead_resnet50"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="skip the bf16 / psroi / train sub-records")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA graph replay")
    ap.add_argument("--no-autotune", action="store_true", help="keep the heuristic conv tile shapes (no per-shape tuning)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: Python's prints keep the real stdout, while file descriptor 1 -- where C code
    # writes (NCCL prints its version banner there whatever NCCL_DEBUG_FILE says) -- is pointed at stderr
    sys.stdout.flush()
    real_out = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_out, "w", buffering=1)
    wl = WORKLOADS[args.workload]()
    if args.impl == "reference":
        wl.run_reference(args)
    else:
        wl.run_own(args)


if __name__ == "__main__":
    main()
