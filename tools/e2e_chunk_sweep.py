#!/usr/bin/env python3
"""Host-buffer (e2e) path: step time vs H2D chunk size (tuning aid for yolo_lp_b200/host.py)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_lp_b200 import synth
from yolo_lp_b200.host import HostPipeline

cfg = synth.CONFIGS[2]
B = cfg["B"]
pred = synth.synth_head(B, cfg["A"], cfg["img"], cfg["n_plates"], cfg["n_pos"], cfg["seed"], pin_memory=True)
dev = torch.device("cuda:0")
d = torch.empty_like(pred, device=dev)
for _ in range(3):
    d.copy_(pred, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    d.copy_(pred, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"raw H2D of the whole tensor: {dt * 1e3:.3f} ms = {pred.numel() * 4 / dt / 1e9:.1f} GB/s")
for chunk in (32, 16, 11, 8, 5, 3, 2):
    pipe = HostPipeline(B, cfg["A"], cfg["max_det"], chunk_images=chunk)
    for _ in range(3):
        pipe.run(pred, cfg["conf"], cfg["iou"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        pipe.run(pred, cfg["conf"], cfg["iou"])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"chunk {chunk:2d} images ({chunk * cfg['A'] * 1160 / 2**20:6.1f} MiB): {dt * 1e3:.3f} ms/step = {B / dt:8.0f} img/s")
