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
  value          images/s over all ranks: the K timed steps are ONE CUDA-graph launch of the
                 multi-stream pipeline (K1 of step k+1 overlaps K2 of step k), CUDA events, max over ranks
  e2e            the same metric through the public API ``non_max_suppression`` with HOST
                 buffers: H2D of the head tensor from pinned memory and D2H of the detections
                 are inside the timed region; ``frac_of_h2d_ceiling`` holds it against the bare
                 concurrent pinned cudaMemcpyAsync rate of the same ranks
  e2e_device_inputs  the production flow when the head runs on the same GPU: fused path from
                 device-resident level tensors to HOST detections (D2H inside the timed region)
  roofline       K1's algorithmic bytes / its CUDA-event time (8 fenced launches) vs the measured HBM peak
  latency        p50 / p95 of one serial batch: config 1 (batch 1) and this workload
  extras         cfg3 strong scaling (256/N images per rank), cfg5 (1280x1280) and, at N=1, cfg4,
                 the decode kernel, the fused path and fp16 head tensors
  cpu_baseline   the UNMODIFIED reference (oracle/_ref, staged by oracle/stage_ref.py) on this host's cores
``--impl reference`` times only that CPU reference.
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
NAMES = {1: "cfg1 YOLO-LP-s 640x640 batch 1", 2: "cfg2 YOLO-LP-s 640x640 batch 32",
         3: "cfg3 YOLO-LP-n 640x640 batch 256", 4: "cfg4 eval stress 640x640 batch 64 conf 0.001",
         5: "cfg5 dense-plate 1280x1280 batch 32"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json config id (1-5) used as the per-GPU workload")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="host-buffer steps (default: min(steps, 20))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the side measurements (other configs, decode, fused path, fp16)")
    return ap.parse_args()


def workload(cid: int):
    from yolo_lp_b200 import synth
    return dict(synth.CONFIGS[cid]), NAMES[cid]


def config_block(cfg, name, n_gpus, B_local):
    return {"workload": name, "images_per_gpu": B_local, "global_batch": B_local * n_gpus, "anchors": cfg["A"],
            "row_floats": 290, "conf_thres": cfg["conf"], "iou_thres": cfg["iou"], "max_det": cfg["max_det"],
            "sharding": f"images x{n_gpus}, no collective",
            "l2": "input %.1f MB per GPU > 126 MB L2 (no flush needed)" % (B_local * cfg["A"] * 1160 / 1e6)
                  if B_local * cfg["A"] * 1160 > 126e6 else
                  "input %.1f MB per GPU fits L2: a 160 MB buffer is rewritten between timed launches" % (B_local * cfg["A"] * 1160 / 1e6)}


# --------------------------------------------------------------------------------------------- CPU arm
def reference_callable():
    """The reference's own ``non_max_suppression``: the unmodified copy staged under oracle/_ref
    (kind "reference"), else the torch-CPU port that the CPU suite pins to it (kind "port")."""
    from oracle import stage_ref
    if stage_ref.is_staged():
        return stage_ref.reference().non_max_suppression, "reference", \
            "unmodified yolov6.utils.nms.non_max_suppression (oracle/_ref) + torchvision.ops.nms"
    from oracle import torch_port
    return torch_port.non_max_suppression, "port", "torch CPU port of nms.py + torchvision.ops.nms"


def cpu_pass(fn, pred, cfg):
    """One pass of the reference's CPU path over the sample, <= CPU_CHUNK images per call;
    the input is cloned outside the timed region because the reference mutates it (nms.py:76)."""
    chunks = [pred[s:s + CPU_CHUNK].clone() for s in range(0, pred.shape[0], CPU_CHUNK)]
    t0 = time.perf_counter()
    n = 0
    for c in chunks:
        out = fn(c, cfg["conf"], cfg["iou"], max_det=cfg["max_det"])
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
    fn, kind, what = reference_callable()
    pred = cpu_sample(cfg, cpu_sample_size(args.config))
    for _ in range(max(1, args.warmup)):
        cpu_pass(fn, pred, cfg)
    times = []
    for _ in range(args.steps):
        dt, _n = cpu_pass(fn, pred, cfg)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = pred.shape[0] / (ms / 1e3)
    sample = f"{pred.shape[0]} images of {name} per step, {CPU_CHUNK}-image calls, {what}"
    B_local = cfg["B"] if args.config != 3 else 32
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_block(cfg, name, args.gpus, B_local),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=JSON_OUT, flush=True)


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while a timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.001):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self.period = period
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
            time.sleep(self.period)

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


# --------------------------------------------------------------------------------------------- helpers
class Ranks:
    """Barrier / max-over-ranks plumbing (NCCL only for these two control-plane collectives)."""

    def __init__(self, world, dev):
        self.world, self.dev = world, dev

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max(self, x):
        import torch
        if self.world == 1:
            return x
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


class L2Flush:
    """Workloads smaller than the 126 MB L2 (config 1: 9.7 MB) would be timed out of cache: rewrite a
    160 MB buffer between timed launches."""

    def __init__(self, dev, needed):
        import torch
        self.buf = torch.empty(160 << 20, dtype=torch.uint8, device=dev) if needed else None

    def __call__(self):
        if self.buf is not None:
            self.buf.add_(1)


def graph_leg(ranks, pipe_factory, capture, K, W, B, clocks=None):
    """``K`` pipelined steps as ONE graph launch, timed with CUDA events round it, max over ranks.
    ``capture(pipe, steps)`` returns a GraphedSteps.  Warm-up: >= W steps through the same graph."""
    import torch
    graphed = capture(pipe_factory(), K)
    replays = max(1, -(-W // K))
    for _ in range(replays):
        graphed.launch()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ranks.barrier()
    ctx = clocks if clocks is not None else _Null()
    with ctx:
        # One more untimed replay goes in FRONT of the timed one, back to back on the same stream: while the
        # GPU runs it the host queues [t0][timed graph][t1], so the events bracket exactly the K timed steps
        # and not the host's launch latency (with 8 ranks on 32 vCPUs a cudaGraphLaunch took ~150 us to
        # arrive -- 15 % of a 20-step window -- which is what capped round 1's scaling curve).
        graphed.launch()
        h0 = time.perf_counter()
        t0.record()
        graphed.launch()
        t1.record()
        host_us = (time.perf_counter() - h0) * 1e6
        ranks.barrier()
    total_ms = ranks.max(t0.elapsed_time(t1))
    return {"ms_per_step": total_ms / K, "total_ms": total_ms, "host_submit_us_per_step": host_us / K,
            "warmup_steps_run": (replays + 1) * K, "graphed": graphed}


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def eager_leg(ranks, pipe, submit, K):
    """The same pipeline driven step by step from Python (one C call per step): what a caller
    without a graph gets, and how much of a step the host needs to submit it."""
    import torch
    pipe.start()
    for _ in range(10):
        submit(pipe)
    pipe.finish()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ranks.barrier()
    t0.record()
    h0 = time.perf_counter()
    pipe.start()
    for _ in range(K):
        submit(pipe)
    pipe.finish()
    host_us = (time.perf_counter() - h0) * 1e6
    t1.record()
    ranks.barrier()
    return ranks.max(t0.elapsed_time(t1)) / K, host_us / K


def fenced_k1(pipe, pred, conf, iou, n=8):
    """``n`` K1 launches inside a running pipeline, each bracketed by CUDA events on its stream and
    fenced off from its neighbours (consecutive K1s otherwise overlap each other's drain and
    ramp-up, and a bracket would measure queueing); K2 of the previous step runs beside them."""
    import torch
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(n)]
    pipe.start()
    for _ in range(4):
        pipe.submit(pred, conf, iou)
    for k in range(n):
        pipe.submit(pred, conf, iou, timing=ev[k])
        pipe.submit(pred, conf, iou)
    pipe.finish()
    torch.cuda.synchronize(pipe.device)
    return [e[0].elapsed_time(e[1]) for e in ev]


def serial_latency(plan, pred, conf, iou, n, flush):
    """p50 / p95 of one batch run strictly serially (K1 then K2 on one stream), per-stage averages."""
    import torch
    for _ in range(5):
        plan.run(pred, conf, iou)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
    torch.cuda.synchronize(plan.device)
    for k in range(n):
        flush()
        ev[k][0].record()
        plan.run_filter(pred, conf)
        ev[k][1].record()
        plan.run_suppress(pred, iou)
        ev[k][2].record()
    torch.cuda.synchronize(plan.device)
    step = sorted(e[0].elapsed_time(e[2]) for e in ev)
    return {"p50_ms": statistics.median(step), "p95_ms": step[int(0.95 * (n - 1))],
            "filter_avg_ms": sum(e[0].elapsed_time(e[1]) for e in ev) / n,
            "nms_avg_ms": sum(e[1].elapsed_time(e[2]) for e in ev) / n, "batches": n}


def fused_latency(levels, cfg, n, flush):
    import torch
    from yolo_lp_b200.head import PostprocessPlan
    plan = PostprocessPlan(levels, (8, 16, 32), cfg["max_det"])
    for _ in range(5):
        plan.run(cfg["conf"], cfg["iou"])
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(n)]
    torch.cuda.synchronize(plan.device)
    for k in range(n):
        flush()
        ev[k][0].record()
        plan.run(cfg["conf"], cfg["iou"])
        ev[k][1].record()
    torch.cuda.synchronize(plan.device)
    step = sorted(e[0].elapsed_time(e[1]) for e in ev)
    return {"p50_ms": statistics.median(step), "p95_ms": step[int(0.95 * (n - 1))], "batches": n}


def api_wall_latency(pred, cfg, n=100):
    """Wall-clock p50 / p95 of the public call on a device tensor -- ``non_max_suppression(pred)`` returning
    the list of per-image rows, i.e. including the launch latencies and the one host sync that reads counts."""
    import torch
    from yolo_lp_b200.nms import non_max_suppression
    for _ in range(5):
        non_max_suppression(pred, cfg["conf"], cfg["iou"], max_det=cfg["max_det"])
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(pred.device)
        t0 = time.perf_counter()
        non_max_suppression(pred, cfg["conf"], cfg["iou"], max_det=cfg["max_det"])
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return {"p50_ms": statistics.median(ts), "p95_ms": ts[int(0.95 * (n - 1))], "calls": n}


def latency_block(cfg, B, pred, dev, plan, in_l2):
    """BASELINE.json's "p50 batch latency": config 1 (tools/infer.py's batch of one image, max_det 1000)
    and this workload, head-tensor entry (lp_nms_f32) and fused entry (lp_detect_postprocess_f32)."""
    from yolo_lp_b200 import synth
    from yolo_lp_b200.nms import NmsPlan
    c1 = dict(synth.CONFIGS[1])
    p1 = synth.synth_head(1, c1["A"], c1["img"], c1["n_plates"], c1["n_pos"], c1["seed"]).to(dev)
    flush = L2Flush(dev, True)
    out = {"cfg1_batch1": {"lp_nms_f32": serial_latency(NmsPlan(1, c1["A"], c1["max_det"], dev), p1, c1["conf"], c1["iou"], 100, flush),
                           "lp_detect_postprocess_f32": fused_latency(synth.synth_levels(1, 640, 640, dev, seed=0), c1, 100, flush),
                           "public_api_wall": api_wall_latency(p1, c1),
                           "l2": "160 MB rewritten before every timed batch of the two device-timed rows (the 9.7 MB input would "
                                 "otherwise sit in L2); public_api_wall is host wall clock per call, input warm"}}
    main = serial_latency(plan, pred, cfg["conf"], cfg["iou"], 100, L2Flush(dev, in_l2))
    out["workload"] = {"lp_nms_f32": main, "public_api_wall": api_wall_latency(pred, cfg),
                       "lp_detect_postprocess_f32": fused_latency(synth.synth_levels(B, cfg["img"], cfg["img"], dev, seed=cfg["seed"]),
                                                                  cfg, 50, L2Flush(dev, in_l2))}
    return out, main


def h2d_ceiling(ranks, host_pred, dev_buf, reps):
    """Bare pinned cudaMemcpyAsync rate of this rank's batch while every other rank does the same."""
    import torch
    s = torch.cuda.Stream(dev_buf.device)
    with torch.cuda.stream(s):
        dev_buf.copy_(host_pred, non_blocking=True)
    s.synchronize()
    ranks.barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(s):
        for _ in range(reps):
            dev_buf.copy_(host_pred, non_blocking=True)
    s.synchronize()
    ms = ranks.max((time.perf_counter() - t0) * 1e3) / reps
    return host_pred.numel() * host_pred.element_size() / ms / 1e6   # GB/s per rank (slowest rank)


def device_inputs_e2e(ranks, cfg, B, dev, Ke):
    """Production flow when the head runs on this GPU: raw level tensors are already in HBM, the
    fused path (KF + K2) turns them into detections and those go D2H into pinned memory -- the timed
    region includes that copy, every step, and the host takes every batch's detections in order.  A batch
    takes ~150 us from its submission to its detections being visible on the host (queueing behind the
    previous batches + KF 56 + K2 33 + D2H 26 us), so four are kept in flight (the host waits for batch
    i - 3 before it submits batch i + 1; tools/device_inputs_probe.py: 142 / 79 / 53 us per step with 2 / 3
    / 4 in flight, 52 us without any host wait), each submitted by ONE native call
    (lp_detect_pipelined_to_host_f32)."""
    import torch
    from yolo_lp_b200 import synth
    from yolo_lp_b200.head import PostprocessPlan, PostprocessPipeline
    DEPTH = 4
    levels = synth.synth_levels(B, cfg["img"], cfg["img"], dev, seed=cfg["seed"])
    plans = [PostprocessPlan(levels, (8, 16, 32), cfg["max_det"]) for _ in range(DEPTH)]
    pipe = PostprocessPipeline(plans)
    out_host = [torch.empty(tuple(p.out.shape), dtype=torch.float32, pin_memory=True) for p in plans]
    cnt_host = [torch.empty((B,), dtype=torch.int32, pin_memory=True) for p in plans]
    copied = [torch.cuda.Event() for _ in plans]
    copy_stream = torch.cuda.Stream(dev)     # D2H on its own stream: K2 of the next batch does not queue behind it

    def run(n):
        """n steps; a slot is resubmitted only after the host has taken its previous batch."""
        inflight, dets = [], 0
        pipe.start()
        for _ in range(n):
            if len(inflight) == DEPTH - 1:
                slot = inflight.pop(0)
                copied[slot].synchronize()
                dets += int(cnt_host[slot][0])             # the host reads the batch it waited for
            slot = pipe.n % DEPTH
            inflight.append(pipe.submit_to_host(cfg["conf"], cfg["iou"], out_host[slot], cnt_host[slot], copy_stream, copied[slot]))
        for slot in inflight:
            copied[slot].synchronize()
        pipe.finish()
        return dets

    run(2 * DEPTH)
    torch.cuda.synchronize(dev)
    ranks.barrier()
    t0 = time.perf_counter()
    run(Ke)
    torch.cuda.synchronize(dev)
    ms = ranks.max((time.perf_counter() - t0) * 1e3) / Ke
    d2h = out_host[0].numel() * 4 + cnt_host[0].numel() * 4
    return {"value": ranks.world * B / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "steps": Ke, "h2d_bytes_per_step": 0,
            "d2h_bytes_per_step": d2h, "detections_per_image": float(cnt_host[0].float().mean()), "batches_in_flight": DEPTH,
            "api": "PostprocessPipeline.submit_to_host (lp_detect_pipelined_to_host_f32: KF + K2 + pinned D2H of out[B,max_det,28] "
                   "and counts[B] in one native call) on device-resident level tensors; the host waits for every batch, in "
                   "order, three steps behind the submission"}


# --------------------------------------------------------------------------------------------- side rows
def other_config(ranks, cid, B_local, first, dev, peak, K, tile_from=None):
    """Pipelined device-resident throughput (graph of K steps) and K1's roofline for another
    BASELINE config on every rank: cfg3 strong scaling, cfg5 (1280^2), cfg4 (dense eval)."""
    import torch
    from yolo_lp_b200 import synth
    from yolo_lp_b200.nms import NmsPipeline
    cfg = dict(synth.CONFIGS[cid])
    unique = min(B_local, tile_from or B_local)
    host = synth.synth_head(unique, cfg["A"], cfg["img"], cfg["n_plates"], cfg["n_pos"], cfg["seed"], first_index=first)
    pred = host.to(dev)
    if unique < B_local:      # large single-GPU shards: the first `unique` images repeated (same work per image)
        pred = pred.repeat((B_local + unique - 1) // unique, 1, 1)[:B_local].contiguous()
    conf, iou = cfg["conf"], cfg["iou"]
    leg = graph_leg(ranks, lambda: NmsPipeline(B_local, cfg["A"], cfg["max_det"], dev),
                    lambda p, k: p.capture(pred, conf, iou, k), K, 3, B_local)
    filt = fenced_k1(NmsPipeline(B_local, cfg["A"], cfg["max_det"], dev), pred, conf, iou, 6)
    counts = leg["graphed"].pipe.plans[0].counts
    bytes_ = B_local * cfg["A"] * 1160
    avg = sum(filt) / len(filt)
    res = {"workload": NAMES[cid], "images_per_gpu": B_local, "global_batch": B_local * ranks.world,
           "value": ranks.world * B_local / leg["ms_per_step"] * 1e3, "unit": UNIT, "ms_per_step": leg["ms_per_step"], "steps": K,
           "k1_avg_launch_ms": avg, "k1_frac_of_hbm_peak": bytes_ / avg / 1e6 / peak, "k1_timed_launches": len(filt),
           "detections_per_image": float(counts.float().mean())}
    if unique < B_local:
        res["note"] = f"{unique} seeded images repeated to fill the {B_local}-image shard"
    del leg, pred
    torch.cuda.empty_cache()
    return res


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
    plans[0].opts = _abi.opts(filter_ctas=sms - min(B, sms // 6))
    t_kf = timed(lambda: plans[0].run_filter(conf))
    plans[0].opts = None
    t_serial = timed(lambda: plans[0].run(conf, iou))
    ranks1 = Ranks(1, dev)
    leg = graph_leg(ranks1, lambda: PostprocessPipeline(plans), lambda p, k: p.capture(conf, iou, k), K, 5, B)
    t_pipe = leg["ms_per_step"]
    dec_bytes, kf_bytes = B * A * (289 + 290) * 4, B * A * 277 * 4
    # the same fused path on fp16 level tensors (model.half()): lp_detect_postprocess_f16 / _pipelined_f16
    half_levels = [{k: v.half() for k, v in lv.items()} for lv in levels]
    del dec, plans, leg, levels
    plans = [PostprocessPlan(half_levels, (8, 16, 32), cfg["max_det"]) for _ in range(2)]
    fused_half = None
    if plans[0].half:
        t_serial_h = timed(lambda: plans[0].run(conf, iou))
        leg = graph_leg(ranks1, lambda: PostprocessPipeline(plans), lambda p, k: p.capture(conf, iou, k), K, 5, B)
        fused_half = {"what": "the fused path on fp16 level tensors (lp_detect_postprocess_f16): exact upcast on load",
                      "serial_ms_per_step": t_serial_h, "pipelined_ms_per_step": leg["ms_per_step"],
                      "images_per_s_pipelined": B / leg["ms_per_step"] * 1e3}
        del leg
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
                       "kf_bytes_over_pipelined_step_frac_of_hbm_peak": kf_bytes / t_pipe / 1e6 / peak,
                       "note": "kf_ms is the kernel alone, launches back to back on ONE stream: its ramp and its finisher "
                               "tail (~15 % of the launch, HBM idle) are inside; in the pipelined step consecutive KFs "
                               "overlap them on alternating streams",
                       "images_per_s_pipelined": B / t_pipe * 1e3, "steps": K}}


def half_measurements(cfg, B, dev, peak, conf, iou, pred, host_pred):
    """the same workload with the head tensor stored as fp16 (the reference's
    --half mode) and read natively (exact upcast on load, fp32 arithmetic).  Informational: the graded
    metric is the fp32 line.  Device-resident pipelined throughput, K1's roofline on the halved bytes,
    and the host-buffer end-to-end rate (half the PCIe bytes)."""
    import torch
    from yolo_lp_b200.nms import NmsPipeline, NmsPlan, non_max_suppression
    A, max_det, K = cfg["A"], cfg["max_det"], 100
    ph = pred.half()
    leg = graph_leg(Ranks(1, dev), lambda: NmsPipeline(B, A, max_det, dev), lambda p, k: p.capture(ph, conf, iou, k), K, 5, B)
    t_step = leg["ms_per_step"]
    del leg
    plan = NmsPlan(B, A, max_det, dev)
    for _ in range(5):
        plan.run_filter(ph, conf)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
    del plan, ph, host_h
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
    ranks = Ranks(world, dev)

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
    in_l2 = B * cfg["A"] * 1160 <= 126e6

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"

    # ---- device-resident leg (headline): the K timed steps are one CUDA-graph launch of the multi-stream
    # pipeline -- K1 (filter) of step k+1 overlaps K2 (sort / NMS / gather) of step k -- so the GPU
    # never waits for the host between steps
    clocks = ClockSampler(visible_to_physical(local))
    leg = graph_leg(ranks, lambda: NmsPipeline(B, cfg["A"], cfg["max_det"], dev),
                    lambda p, k: p.capture(pred, conf, iou, k), K, W, B, clocks)
    ms_per_step = leg["ms_per_step"]
    value = world * B / (ms_per_step / 1e3)
    gplans = leg["graphed"].pipe.plans
    counts = gplans[0].counts.cpu()
    assert all(torch.equal(pl.counts.cpu(), counts) for pl in gplans)

    # ---- the same pipeline driven step by step from Python, with the clocks sampled over a longer run
    Kl = max(K, 200)
    clocks_long = ClockSampler(visible_to_physical(local), period=0.002)
    epipe = NmsPipeline(B, cfg["A"], cfg["max_det"], dev)
    with clocks_long:
        eager_ms, eager_host_us = eager_leg(ranks, epipe, lambda p: p.submit(pred, conf, iou), Kl)
    assert all(torch.equal(pl.counts.cpu(), counts) for pl in epipe.plans), "graphed and eager pipelines disagree"

    # ---- roofline of the dominant kernel (K1): 8 fenced launches inside the running pipeline
    filt_ms = fenced_k1(epipe, pred, conf, iou, 8)

    # ---- latency legs (strictly serial batches)
    latency, main_lat = latency_block(cfg, B, pred, dev, plan, in_l2)
    assert torch.equal(plan.counts.cpu(), counts), "pipelined and serial paths disagree"
    cand = plan.candidate_counts().cpu()

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
    lat_filter, lat_nms = main_lat["filter_avg_ms"], main_lat["nms_avg_ms"]
    roofline = {"kernel": "lp::filter_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": filt_avg,
                "launch_ms": [round(t, 5) for t in filt_ms],
                "share_of_step": filt_avg / ms_per_step,
                "share_of_serial_step": lat_filter / (lat_filter + lat_nms),   # the figure an ncu launch list (serialised) shows
                "timed_launches": len(filt_ms),
                "note": "8 K1 launches of a leg that follows the timed region, inside the running pipeline (K2 of the previous "
                        "step beside them), each fenced off from the neighbouring K1s -- which in the timed region overlap each "
                        "other's drain and ramp-up, hence share_of_step can exceed 1"
                        + ("; this workload fits L2, so the launches read from cache" if in_l2 else "")}

    # ---- end-to-end leg: public API on a pinned HOST tensor; H2D + kernels + D2H per step
    Ke = args.e2e_steps or max(3, min(K, 20))
    hpipe = host_pipeline(B, cfg["A"], cfg["max_det"])
    for _ in range(3):
        res = non_max_suppression(host_pred, conf, iou, max_det=cfg["max_det"])
    ranks.barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        res = non_max_suppression(host_pred, conf, iou, max_det=cfg["max_det"])
    torch.cuda.synchronize(dev)
    e2e_ms = ranks.max((time.perf_counter() - t0) * 1e3) / Ke
    assert [int(r.shape[0]) for r in res] == counts.tolist(), "host-buffer path disagrees with the device path"
    ceiling_gbs = h2d_ceiling(ranks, host_pred, torch.empty_like(pred), max(3, Ke // 2))
    e2e_gbs = hpipe.h2d_bytes / e2e_ms / 1e6
    e2e = {"value": world * B / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": Ke,
           "h2d_bytes_per_step": hpipe.h2d_bytes, "d2h_bytes_per_step": hpipe.d2h_bytes,
           "h2d_gbs_per_rank": e2e_gbs, "h2d_ceiling_gbs_per_rank": ceiling_gbs, "frac_of_h2d_ceiling": e2e_gbs / ceiling_gbs,
           "ceiling": "bare pinned cudaMemcpyAsync of the same tensor on every rank at once (slowest rank)",
           "api": "yolo_lp_b200.non_max_suppression(cpu pinned tensor) -> lp_nms_f32 per 48 MiB chunk"}
    ranks.barrier()
    e2e_dev = device_inputs_e2e(ranks, cfg, B, dev, max(Ke, 50))
    ranks.barrier()

    extras = {}
    if not args.no_extras:
        Kx = min(max(K, 20), 100)
        # BASELINE config 3: the 256-image batch image-sharded over the ranks (strong scaling: 256/N per GPU)
        extras["cfg3_strong"] = other_config(ranks, 3, 256 // world, rank * (256 // world), dev, peak, Kx, tile_from=32)
        extras["cfg3_strong"]["scaling"] = "strong: global batch 256 fixed, 256/N images per GPU"
        # BASELINE config 5: 1280x1280, 32 images per GPU (weak)
        extras["cfg5_1280"] = other_config(ranks, 5, 32, rank * 32, dev, peak, min(Kx, 40), tile_from=8)
        if world == 1:
            extras["cfg4_dense_eval"] = other_config(ranks, 4, 64, 0, dev, peak, min(Kx, 50), tile_from=16)
            extras.update(side_measurements(cfg, B, dev, peak, conf, iou))
            extras["half_head_tensor"] = half_measurements(cfg, B, dev, peak, conf, iou, pred, host_pred)
    ranks.barrier()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_block(cfg, name, world, B),
                "pipeline": "the K timed steps are one CUDA graph (NmsPipeline.capture): K1 alternates between 2 streams, K2 on a "
                            "third, 2 workspaces -- K1 of step k+1 overlaps K2 of step k and the drain of K1 of step k",
                "warmup_steps_run": leg["warmup_steps_run"],
                "host_submit_us_per_step": leg["host_submit_us_per_step"],
                "eager": {"ms_per_step": eager_ms, "value": world * B / eager_ms * 1e3, "host_submit_us_per_step": eager_host_us,
                          "steps": Kl, "what": "the same pipeline submitted step by step from Python (lp_nms_pipelined_f32 per step)"},
                "p50_batch_latency_ms": main_lat["p50_ms"], "p95_batch_latency_ms": main_lat["p95_ms"],
                "serial_stage_ms": {"filter_avg": lat_filter, "nms_avg": lat_nms},
                "latency": latency,
                "detections_per_image": sum(counts.tolist()) / B, "candidates_per_image": float(cand.sum()) / B,
                "roofline": roofline, "e2e": e2e, "e2e_device_inputs": e2e_dev,
                "clocks": dict(clocks.summary(), sustained=clocks_long.summary()), "extras": extras or None,
                "gpu_launches": K * NmsPlan.KERNELS_PER_CALL}
        if world == 1 and not args.no_cpu_baseline:
            restore_affinity(prev_affinity)
            use_all_host_threads()
            fn, kind, what = reference_callable()
            sample = cpu_sample(cfg, cpu_sample_size(args.config))
            cpu_pass(fn, sample, cfg)                  # warm-up
            spent, images, passes = 0.0, 0, 0
            while spent < 10.0 and passes < 1000:     # ~10 s of CPU work on the bounded sample
                dt, n = cpu_pass(fn, sample, cfg)
                spent, images, passes = spent + dt, images + n, passes + 1
            line["cpu_baseline"] = {
                "value": images / spent, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                "host_cpus": os.cpu_count(),
                "sample": f"{passes} passes over {sample.shape[0]} images of {name} ({spent:.1f} s), "
                          f"{CPU_CHUNK}-image calls, {what}"}
            # second baseline of SURVEY 8-d: the reference's own GPU route (what tools/infer.py --device 0
            # does today) -- the same function on CUDA tensors: ~45 ATen launches per image + torchvision's
            # generic CUDA nms kernel.  Not used for parity (it is not bit-identical to the CPU kernel).
            try:
                gsample = sample.to(dev)
                cpu_pass(fn, gsample, cfg)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                n_img = 0
                for _ in range(3):
                    _dt, n = cpu_pass(fn, gsample, cfg)
                    n_img += n
                torch.cuda.synchronize(dev)
                line["reference_cuda_route"] = {
                    "value": n_img / (time.perf_counter() - t0), "unit": UNIT,
                    "what": f"{what} on CUDA tensors of the same B200 (ATen kernels + torchvision CUDA nms), "
                            "input clone included, 8-image calls"}
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
