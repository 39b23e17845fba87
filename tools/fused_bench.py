#!/usr/bin/env python3
"""Fused raw-levels -> detections path: KF roofline and serial / pipelined throughput vs the
unfused decode -> K1 -> K2 chain on the same level tensors.

    python tools/fused_bench.py [B] [img] [conf]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_lp_b200 import synth
from yolo_lp_b200.head import DecodePlan, PostprocessPlan
from yolo_lp_b200.nms import NmsPlan

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
img = int(sys.argv[2]) if len(sys.argv) > 2 else 640
conf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
iou, max_det = 0.45, 300
dev = torch.device("cuda:0")
levels = synth.synth_levels(B, img, img, dev, seed=0)
HALF = os.environ.get("LEVELS_HALF") == "1"   # fp16 level tensors: the fused path reads halves, the unfused chain the upcast copy
if HALF:
    half_levels = [{k: v.half() for k, v in lv.items()} for lv in levels]
    levels = [{k: v.float() for k, v in lv.items()} for lv in half_levels]
A = sum(h * w for h, w in synth.level_shapes(img, img))
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, k=100):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k


fused = PostprocessPlan(half_levels if HALF else levels, (8, 16, 32), max_det)
dec = DecodePlan(levels, (8, 16, 32))
nms = NmsPlan(B, A, max_det, dev)
t_kf = timed(lambda: fused.run_filter(conf))
t_fused = timed(lambda: fused.run(conf, iou))
t_dec = timed(dec.run)
t_unfused = timed(lambda: (dec.run(), nms.run(dec.out, conf, iou)))
fused.run(conf, iou); nms.run(dec.out, conf, iou); torch.cuda.synchronize()
assert torch.equal(fused.counts, nms.counts)
for b in range(B):
    k = int(fused.counts[b])
    assert torch.equal(fused.out[b, :k], nms.out[b, :k])
algo = B * A * 277 * (2 if HALF else 4)
print(json.dumps({
    "B": B, "img": img, "A": A, "conf": conf, "levels_dtype": "f16" if HALF else "f32", "candidates_per_image": float(fused.candidate_counts().float().mean()),
    "kept_per_image": float(fused.counts.float().mean()),
    "KF_ms": t_kf, "KF_algorithmic_bytes": algo, "KF_gbs": algo / t_kf / 1e6, "KF_frac_of_measured_hbm": algo / t_kf / 1e6 / peak,
    "fused_serial_ms": t_fused, "fused_images_per_s": B / t_fused * 1e3,
    "decode_ms": t_dec, "unfused_serial_ms": t_unfused, "unfused_images_per_s": B / t_unfused * 1e3,
    "speedup_fused_vs_unfused": t_unfused / t_fused, "bit_identical": True}))
