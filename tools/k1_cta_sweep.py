#!/usr/bin/env python3
"""K1 alone vs its CTA count, for a batch large enough that the default heuristic reserves many SMs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_lp_b200 import _abi
from yolo_lp_b200.nms import NmsPlan

B, A = 128, 8400
dev = torch.device("cuda:0")
pred = torch.rand((B, A, 290), device=dev) * 0.1
plan = NmsPlan(B, A, 300, dev)
for ctas in (0, 148, 116, 100, 84, 74, 60, 48):
    plan.opts = _abi.opts(filter_ctas=ctas)
    for _ in range(5):
        plan.run_filter(pred, 0.25)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(30):
        plan.run_filter(pred, 0.25)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 30
    print(f"K1 ctas={ctas or 'auto':>4}: {ms * 1e3:7.1f} us  {B * A * 1160 / ms / 1e6:7.0f} GB/s")
