#!/usr/bin/env python3
"""Benchmark of the YOLO-LP post-processing hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2] [--impl ours|reference]

A "step" is one pass of the hot path -- confidence filter + sort + greedy NMS + output
rows (``lp_nms_f32``: K1 ``lp::filter_kernel`` + K2 ``lp::nms_kernel``) -- over one batch of
synthetic head tensors ``[B, A, 290]`` that is already resident in HBM.  Workload at N=1 is
BASELINE.json configs[1] (YOLO-LP-s 640x640, batch 32, conf 0.25 / IoU 0.45, max_det 300);
N>1 is weak scaling, every rank owning its own 32-image shard of a 32*N-image batch with no
collective on the data path (N=8 is configs[2]'s 256-image batch).

One JSON line is printed by rank 0 with
  value          images/s over all ranks, device-timed (CUDA events), max over ranks
  e2e            the same metric through the public API ``non_max_suppression`` with HOST
                 buffers: H2D of the head tensor from pinned memory and D2H of the detections
                 are inside the timed region
  roofline       K1's algorithmic bytes / its CUDA-event time vs the measured HBM peak
  cpu_baseline   the torch-CPU port of the reference (oracle/torch_port.py) on this host
``--impl reference`` times only that CPU port (the reference itself is not on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "post-proc images/sec"
UNIT = "images/s"
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback
CPU_CHUNK = 8                      # reference is driven in <= 8-image chunks (its 10 s time_limit)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json config id (1-5) used as the per-GPU workload")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="host-buffer steps (default: min(steps, 20))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the decode-kernel and fused-path side measurements")
    return ap.parse_args()


def workload(cid: int):
    from yolo_lp_b200 import synth
    cfg = dict(synth.CONFIGS[cid])
    name = {1: "cfg1 YOLO-LP-s 640x640 batch 1", 2: "cfg2 YOLO-LP-s 640x640 batch 32",
            3: "cfg3 YOLO-LP-n 640x640 batch 256", 4: "cfg4 eval stress 640x640 batch 64 conf 0.001",
            5: "cfg5 dense-plate 1280x1280 batch 32"}[cid]
    return cfg, name


def config_block(cfg, name, n_gpus, B_local):
    return {"workload": name, "images_per_gpu": B_local, "global_batch": B_local * n_gpus, "anchors": cfg["A"],
            "row_floats": 290, "conf_thres": cfg["conf"], "iou_thres": cfg["iou"], "max_det": cfg["max_det"],
            "sharding": f"images x{n_gpus}, no collective",
            "l2": "input %.1f MB per GPU > 126 MB L2 (no flush needed)" % (B_local * cfg["A"] * 1160 / 1e6)}


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_pass(pred, cfg):
    """One pass of the reference's CPU path over the sample, <= CPU_CHUNK images per call;
    the input is cloned outside the timed region because the port mutates it like nms.py:76."""
    import torch
    from oracle import torch_port
    chunks = [pred[s:s + CPU_CHUNK].clone() for s in range(0, pred.shape[0], CPU_CHUNK)]
    t0 = time.perf_counter()
    n = 0
    for c in chunks:
        out = torch_port.non_max_suppression(c, cfg["conf"], cfg["iou"], max_det=cfg["max_det"])
        n += len(out)
    return time.perf_counter() - t0, n


def cpu_sample(cfg, budget_images):
    from yolo_lp_b200 import synth
    B = min(cfg["B"], budget_images)
    return synth.synth_head(B, cfg["A"], cfg["img"], cfg["n_plates"], cfg["n_pos"], cfg["seed"])


def cpu_sample_size(cid):
    # ~10 ms/img normal, ~0.5 s/img dense, ~0.13 s/img at 1280^2 on 8 cores
    return {1: 1, 2: 32, 3: 32, 4: 4, 5: 8}[cid]


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU reference must run on all host cores."""
    import torch
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def run_reference_arm(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    cfg, name = workload(args.config)
    pred = cpu_sample(cfg, cpu_sample_size(args.config))
    for _ in range(max(1, args.warmup)):
        cpu_pass(pred, cfg)
    times = []
    for _ in range(args.steps):
        dt, _n = cpu_pass(pred, cfg)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = pred.shape[0] / (ms / 1e3)
    sample = f"{pred.shape[0]} images of {name} per step, {CPU_CHUNK}-image calls, torch CPU port of nms.py + torchvision.ops.nms"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_block(cfg, name, args.gpus, cfg["B"]),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=JSON_OUT, flush=True)


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while a timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def _once(self):
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _loop(self):
        while not self._stop.is_set():
            self._once()
            time.sleep(0.004)

    def __enter__(self):
        if self.ok:
            self._stop.clear()
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        if self.ok:
            self._once()
            self._stop.set()
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_near_gpu(physical_index):
    """Pin this process to the CPUs NVML reports as local to the GPU, so that the pinned host buffers
    allocated next land on the GPU's NUMA node (51 vs 55.5 GB/s H2D were measured for buffers on the
    far / near node).  Returns the previous affinity (restore it before the CPU baseline) or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(physical_index)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        prev = os.sched_getaffinity(0)
        cpus &= prev
        if cpus:
            os.sched_setaffinity(0, cpus)
            return prev
    except Exception:
        pass
    return None


def restore_affinity(prev):
    """Undo bind_near_gpu for EVERY thread of the process (OpenMP workers spawned while bound inherited
    the narrow mask), so the CPU baseline really gets all host cores."""
    if not prev:
        return
    try:
        for tid in os.listdir("/proc/self/task"):
            try:
                os.sched_setaffinity(int(tid), prev)
            except OSError:
                pass
    except OSError:
        os.sched_setaffinity(0, prev)


def visible_to_physical(local_index):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_index])
        except Exception:
            return local_index
    return local_index


# --------------------------------------------------------------------------------------------- side rows
def side_measurements(cfg, B, dev, peak, conf, iou, K=50):
    """SURVEY 8 rows next to the headline path, on the same shape: the Detect eval-tail decode kernel
    (raw level tensors -> [B,A,290]) and the fused path (raw level tensors -> detections).  Synthetic
    level tensors generated on the device; CUDA events; informational (not part of `value`)."""
    import torch
    from yolo_lp_b200 import _abi, synth
    from yolo_lp_b200.head import DecodePlan, PostprocessPlan, PostprocessPipeline
    img = cfg["img"]
    levels = synth.synth_levels(B, img, img, dev, seed=cfg["seed"])
    A = cfg["A"]

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(K):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / K

    dec = DecodePlan(levels, (8, 16, 32))
    t_dec = timed(dec.run)
    plans = [PostprocessPlan(levels, (8, 16, 32), cfg["max_det"]) for _ in range(2)]
    # KF alone, on the CTA count the serial entry uses (#SMs - #SMs/6); the stage entry on its own
    # would leave one SM per image free for an overlapping K2, which is what the pipelined leg measures
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    _abi.call("lp_tune", 0, sms - min(B, sms // 6))
    try:
        t_kf = timed(lambda: plans[0].run_filter(conf))
    finally:
        _abi.call("lp_tune", 0, 0)
    t_serial = timed(lambda: plans[0].run(conf, iou))
    pipe = PostprocessPipeline(plans)

    def burst():
        pipe.start()
        for _ in range(K):
            pipe.submit(conf, iou)
        pipe.finish()
    burst()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    burst()
    b.record()
    torch.cuda.synchronize(dev)
    t_pipe = a.elapsed_time(b) / K
    dec_bytes, kf_bytes = B * A * (289 + 290) * 4, B * A * 277 * 4
    # the same fused path on fp16 level tensors (model.half()): lp_detect_postprocess_f16 / _pipelined_f16
    half_levels = [{k: v.half() for k, v in lv.items()} for lv in levels]
    del dec, plans, pipe, levels
    plans = [PostprocessPlan(half_levels, (8, 16, 32), cfg["max_det"]) for _ in range(2)]
    fused_half = None
    if plans[0].half:
        t_serial_h = timed(lambda: plans[0].run(conf, iou))
        pipe = PostprocessPipeline(plans)
        burst()
        torch.cuda.synchronize(dev)
        a.record()
        burst()
        b.record()
        torch.cuda.synchronize(dev)
        fused_half = {"what": "the fused path on fp16 level tensors (lp_detect_postprocess_f16): exact upcast on load",
                      "serial_ms_per_step": t_serial_h, "pipelined_ms_per_step": a.elapsed_time(b) / K,
                      "images_per_s_pipelined": B / (a.elapsed_time(b) / K) * 1e3}
        del pipe
    del plans, half_levels
    torch.cuda.empty_cache()
    return {
        "fused_path_half_levels": fused_half,
        "decode_kernel": {"kernel": "lp::decode_tma_kernel", "ms": t_dec, "algorithmic_bytes": dec_bytes,
                          "achieved_gbs": dec_bytes / t_dec / 1e6, "frac_of_hbm_peak": dec_bytes / t_dec / 1e6 / peak},
        "fused_path": {"what": "raw level tensors -> detections (lp_detect_postprocess_f32), no [B,A,290] tensor",
                       "kf_kernel": "lp::levels_filter_tma_kernel", "kf_ms": t_kf, "kf_algorithmic_bytes": kf_bytes,
                       "kf_achieved_gbs": kf_bytes / t_kf / 1e6, "kf_frac_of_hbm_peak": kf_bytes / t_kf / 1e6 / peak,
                       "serial_ms_per_step": t_serial, "pipelined_ms_per_step": t_pipe,
                       "images_per_s_pipelined": B / t_pipe * 1e3, "steps": K}}


def half_measurements(cfg, B, dev, peak, conf, iou, pred, host_pred, counts32):
    """the same workload with the head tensor stored as fp16 (the reference's
    --half mode) and read natively (exact upcast on load, fp32 arithmetic).  Informational: the graded
    metric is the fp32 line.  Device-resident pipelined throughput, K1's roofline on the halved bytes,
    and the host-buffer end-to-end rate (half the PCIe bytes)."""
    import time
    import torch
    from yolo_lp_b200.nms import NmsPipeline, NmsPlan, non_max_suppression
    A, max_det, K = cfg["A"], cfg["max_det"], 100
    ph = pred.half()
    pipe = NmsPipeline(B, A, max_det, dev)

    def burst():
        pipe.start()
        for _ in range(K):
            pipe.submit(ph, conf, iou)
        pipe.finish()
    burst()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    burst()
    b.record()
    torch.cuda.synchronize(dev)
    t_step = a.elapsed_time(b) / K
    plan = NmsPlan(B, A, max_det, dev)
    for _ in range(5):
        plan.run_filter(ph, conf)
    a.record()
    for _ in range(K):
        plan.run_filter(ph, conf)
    b.record()
    torch.cuda.synchronize(dev)
    t_k1 = a.elapsed_time(b) / K
    host_h = torch.empty(host_pred.shape, dtype=torch.float16, pin_memory=True)
    host_h.copy_(host_pred)
    for _ in range(3):
        non_max_suppression(host_h, conf, iou, max_det=max_det)
    Ke = 10
    t0 = time.perf_counter()
    for _ in range(Ke):
        non_max_suppression(host_h, conf, iou, max_det=max_det)
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / Ke
    k1_bytes = B * A * 580
    del pipe, plan, ph, host_h
    torch.cuda.empty_cache()
    return {"what": "fp16 head tensor [B,A,290] (reference --half mode), upcast exactly on load; results == fp32 path on pred.float()",
            "kernel": "lp::filter_half_kernel", "images_per_s_pipelined": B / t_step * 1e3, "pipelined_ms_per_step": t_step,
            "k1_ms_serial_incl_memset": t_k1, "k1_algorithmic_bytes": k1_bytes, "k1_achieved_gbs": k1_bytes / t_k1 / 1e6,
            "k1_frac_of_hbm_peak": k1_bytes / t_k1 / 1e6 / peak,
            "e2e_images_per_s": B / e2e_ms * 1e3, "e2e_ms_per_step": e2e_ms, "h2d_bytes_per_step": k1_bytes}


# --------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from yolo_lp_b200 import synth
    from yolo_lp_b200.nms import NmsPipeline, NmsPlan, non_max_suppression
    from yolo_lp_b200.host import host_pipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:   # launched bare: re-exec under torchrun, one rank per GPU
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000)] + sys.argv
            sys.exit(subprocess.call(cmd, stdout=JSON_OUT.fileno()))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: yolo_lp_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg, name = workload(args.config)
    B = cfg["B"] if args.config != 3 else 32      # per-GPU shard; config 3 is the 8-GPU aggregate of config 2's shape
    first = rank * B                               # this rank's contiguous image range of the global batch
    prev_affinity = bind_near_gpu(visible_to_physical(local))
    host_pred = synth.synth_head(B, cfg["A"], cfg["img"], cfg["n_plates"], cfg["n_pos"], cfg["seed"],
                                 first_index=first, pin_memory=True)
    pred = host_pred.to(dev)
    plan = NmsPlan(B, cfg["A"], cfg["max_det"], dev)
    conf, iou = cfg["conf"], cfg["iou"]
    K, W = args.steps, max(3, args.warmup)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg (headline): K steps through the two-stream pipeline -- K1 (filter) of
    # step k+1 overlaps K2 (sort/NMS/gather) of step k; CUDA events round every K1 launch
    pipe = NmsPipeline(B, cfg["A"], cfg["max_det"], dev)
    pipe.start()
    for _ in range(W):
        pipe.submit(pred, conf, iou)
    pipe.finish()
    # every TIME_EVERY-th K1 launch is bracketed by timing events.  A timed launch is fenced off from
    # its neighbours (consecutive K1s otherwise overlap each other's drain and ramp-up, and a bracket
    # would then measure queueing, not the kernel), which costs ~20 us of bubble per timed launch --
    # so only a handful of launches of the timed region (one per 32 steps, six at most) are measured this way.
    n_timed = max(1, min(6, K // 32))
    TIME_EVERY = max(1, K // n_timed)
    ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(2)] for k in list(range(TIME_EVERY // 2, K, TIME_EVERY))[:n_timed]}   # never step 0: the pipeline is still filling
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(visible_to_physical(local))
    barrier()
    with clocks:
        t_begin.record()
        pipe.start()
        for k in range(K):
            pipe.submit(pred, conf, iou, timing=ev.get(k))
        pipe.finish()
        t_end.record()
        barrier()
    total_ms = max_over_ranks(t_begin.elapsed_time(t_end))
    ms_per_step = total_ms / K
    filt_ms = [e[0].elapsed_time(e[1]) for e in ev.values()]
    counts = pipe.plans[0].counts.cpu()
    assert all(torch.equal(pl.counts.cpu(), counts) for pl in pipe.plans)
    value = world * B / (ms_per_step / 1e3)

    # ---- latency leg: the same K1, K2 strictly one after the other on one stream (no overlap)
    Kl = min(K, 100)
    for _ in range(W):
        plan.run(pred, conf, iou)
    lv = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(Kl)]
    barrier()
    for k in range(Kl):
        lv[k][0].record()
        plan.run_filter(pred, conf)
        lv[k][1].record()
        plan.run_suppress(pred, iou)
        lv[k][2].record()
    barrier()
    step_ms = [lv[k][0].elapsed_time(lv[k][2]) for k in range(Kl)]
    lat_filter = sum(lv[k][0].elapsed_time(lv[k][1]) for k in range(Kl)) / Kl
    lat_nms = sum(lv[k][1].elapsed_time(lv[k][2]) for k in range(Kl)) / Kl
    assert torch.equal(plan.counts.cpu(), counts), "pipelined and serial paths disagree"
    cand = plan.candidate_counts().cpu()

    # ---- roofline of the dominant kernel (K1): algorithmic bytes = every row read once + 8 B per survivor
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    algo_bytes = B * cfg["A"] * 1160 + int(cand.sum()) * 8
    filt_avg = sum(filt_ms) / len(filt_ms)
    achieved = algo_bytes / (filt_avg * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"cfg{args.config}", {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"kernel": "lp::filter_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": filt_avg,
                "share_of_step": filt_avg / ms_per_step,
                "share_of_serial_step": lat_filter / (lat_filter + lat_nms),   # the figure an ncu launch list (serialised) shows
                "timed_launches": len(filt_ms),
                "note": "K1 timed inside the pipelined region, K2 of the previous step running concurrently; the timed launches "
                        "(one per 32 steps, six at most) are fenced off from the neighbouring K1s, which otherwise overlap each other's drain and "
                        "ramp-up -- hence share_of_step > 1"}

    # ---- end-to-end leg: public API on a pinned HOST tensor; H2D + kernels + D2H per step
    Ke = args.e2e_steps or max(3, min(K, 20))
    pipe = host_pipeline(B, cfg["A"], cfg["max_det"])
    for _ in range(3):
        res = non_max_suppression(host_pred, conf, iou, max_det=cfg["max_det"])
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        res = non_max_suppression(host_pred, conf, iou, max_det=cfg["max_det"])
    torch.cuda.synchronize(dev)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / Ke
    assert [int(r.shape[0]) for r in res] == counts.tolist(), "host-buffer path disagrees with the device path"
    e2e = {"value": world * B / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": Ke,
           "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
           "api": "yolo_lp_b200.non_max_suppression(cpu pinned tensor) -> lp_nms_f32 per 48 MiB chunk"}
    barrier()

    extras = None
    if not args.no_extras and rank == 0:
        extras = side_measurements(cfg, B, dev, peak, conf, iou)
        extras["half_head_tensor"] = half_measurements(cfg, B, dev, peak, conf, iou, pred, host_pred, counts)
    barrier()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_block(cfg, name, world, B),
                "pipeline": "K1 alternates between 2 streams, K2 on a third, 2 workspaces: K1 of step k+1 overlaps K2 of step k "
                            "and the drain of K1 of step k",
                "p50_batch_latency_ms": statistics.median(step_ms), "p95_batch_latency_ms": sorted(step_ms)[int(0.95 * (Kl - 1))],
                "serial_stage_ms": {"filter_avg": lat_filter, "nms_avg": lat_nms},
                "detections_per_image": sum(counts.tolist()) / B, "candidates_per_image": float(cand.sum()) / B,
                "roofline": roofline, "e2e": e2e, "clocks": clocks.summary(), "extras": extras,
                "gpu_launches": K * NmsPlan.KERNELS_PER_CALL}
        if world == 1 and not args.no_cpu_baseline:
            restore_affinity(prev_affinity)
            use_all_host_threads()
            sample = cpu_sample(cfg, cpu_sample_size(args.config))
            cpu_pass(sample, cfg)                      # warm-up
            spent, images, passes = 0.0, 0, 0
            while spent < 10.0 and passes < 1000:     # ~10 s of CPU work on the bounded sample
                dt, n = cpu_pass(sample, cfg)
                spent, images, passes = spent + dt, images + n, passes + 1
            line["cpu_baseline"] = {
                "value": images / spent, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                "host_cpus": os.cpu_count(),
                "sample": f"{passes} passes over {sample.shape[0]} images of {name} ({spent:.1f} s), "
                          f"{CPU_CHUNK}-image calls, torch CPU port of nms.py + torchvision.ops.nms"}
            # second baseline of SURVEY 8-d: the reference's own GPU route (what tools/infer.py --device 0
            # does today) -- the same port on CUDA tensors: ~45 ATen launches per image + torchvision's
            # generic CUDA nms kernel.  Not used for parity (it is not bit-identical to the CPU kernel).
            try:
                gsample = sample.to(dev)
                cpu_pass(gsample, cfg)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                n_img = 0
                for _ in range(3):
                    _dt, n = cpu_pass(gsample, cfg)
                    n_img += n
                torch.cuda.synchronize(dev)
                line["reference_cuda_route"] = {
                    "value": n_img / (time.perf_counter() - t0), "unit": UNIT,
                    "what": "torch port of the reference on CUDA tensors of the same B200 (ATen kernels + "
                            "torchvision CUDA nms), input clone included, 8-image calls"}
            except Exception as exc:  # torchvision CUDA ops missing etc.: report, do not fail the bench
                line["reference_cuda_route"] = {"unavailable": repr(exc)[:200]}
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def claim_stdout():
    """stdout must carry exactly ONE JSON line, but native libraries print there too (NCCL's
    "NCCL version ..." banner under torchrun).  Keep a private duplicate of the real stdout for the
    JSON line and point file descriptor 1 at stderr for everything else."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


if __name__ == "__main__":
    a = parse()
    JSON_OUT = claim_stdout()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)
