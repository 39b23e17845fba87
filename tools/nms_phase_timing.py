#!/usr/bin/env python3
"""Per-phase clock64 breakdown of K2 (lp::nms_kernel) for one BASELINE config (debug aid).

    python tools/nms_phase_timing.py [config id]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_lp_b200 import _abi, synth
from yolo_lp_b200.nms import NmsPlan

cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = synth.CONFIGS[cid]
B = min(cfg["B"], 32)
dev = torch.device("cuda:0")
pred = synth.synth_head(B, cfg["A"], cfg["img"], cfg["n_plates"], cfg["n_pos"], cfg["seed"]).to(dev)
plan = NmsPlan(B, cfg["A"], cfg["max_det"], dev)
buf = torch.zeros((B, 16), dtype=torch.int64, device=dev)
for _ in range(3):
    plan.run(pred, cfg["conf"], cfg["iou"])
plan.opts = _abi.opts(timing=buf.data_ptr())
plan.run(pred, cfg["conf"], cfg["iou"])
torch.cuda.synchronize()
plan.opts = None
t = buf.cpu()
names = ["order (sort / histogram)", "first segment + window", "nms", "gather"]
d = torch.stack([t[:, 2] - t[:, 0], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3], t[:, 5] - t[:, 4]], 1).double()
print(f"cfg{cid}: B={B}  candidates/img={plan.candidate_counts().float().mean().item():.0f}  kept/img={plan.counts.float().mean().item():.0f}")
for i, n in enumerate(names):
    print(f"  {n:26s} mean {d[:, i].mean():9.0f} clk   max {d[:, i].max():9.0f} clk")
if int(t[:, 7].sum()) > 0:   # segmented ordering: split "first segment + window"
    print(f"    first segment: {t[:, 14].double().mean():.0f} keys; compaction {(t[:, 1] - t[:, 2]).double().mean():.0f} clk, "
          f"sort {(t[:, 7] - t[:, 1]).double().mean():.0f} clk, window staging {(t[:, 3] - t[:, 7]).double().mean():.0f} clk")
g1 = (t[:, 6] - t[:, 4]).double()
print(f"  gather: first batch of rows staged after {g1.mean():.0f} clk, its groups re-scored after another {(t[:, 15] - t[:, 6]).double().mean():.0f} clk")
q = t[:, 8:14].double()
if float(q[:, 5].sum()) > 0:   # library built with LPNMS_NVCC_EXTRA=-DLP_NMS_PROFILE
    n = q[:, 5].clamp(min=1)
    print(f"  last warp: {n.mean():.1f} chunk steps, window staging {q[:, 0].mean():.0f} clk in all; per step: test {(q[:, 1] / n).mean():.0f}  "
          f"barrier {(q[:, 2] / n).mean():.0f}  settle (warp 0; others skip) {(q[:, 3] / n).mean():.0f}  barrier {(q[:, 4] / n).mean():.0f} clk")
tot = (t[:, 5] - t[:, 0]).double()
print(f"  total          mean {tot.mean():9.0f} clk   max {tot.max():9.0f} clk   (SM clock ~1.9 GHz -> {tot.max() / 1.9e3:.1f} us)")
