#!/usr/bin/env python3
"""Exhaustive check on the GPU that the device sigmoid (csrc/common.cuh sigmoid_f32) is monotone
non-decreasing over every finite fp32 input -- the property that makes
max_j sigmoid(x_j) == sigmoid(max_j x_j) exact in the fused path -- and its worst relative error
against float64 on a dense sample.  (The same sweep runs as a gpu test.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_lp_b200 import _abi

dev = torch.device("cuda:0")
CH = 1 << 26
stream = torch.cuda.current_stream(dev).cuda_stream
bad, worst = 0, 0.0
for sign in (0, 1):   # positive floats ascend with their bit pattern, negative floats descend
    prev = None
    for lo in range(0, 0x7f800000, CH):
        hi = min(lo + CH, 0x7f800000)
        bits = torch.arange(lo, hi, dtype=torch.int64, device=dev) + (0x80000000 if sign else 0)
        x = bits.to(torch.uint32).view(torch.float32)
        y = torch.empty_like(x)
        _abi.call("lp_debug_sigmoid_f32", x.data_ptr(), x.numel(), y.data_ptr(), stream)
        d = y[1:] - y[:-1]
        bad += int(((d < 0) if not sign else (d > 0)).sum())
        if prev is not None:
            bad += int((y[0] < prev) if not sign else (y[0] > prev))
        prev = y[-1].clone()
        xs = x[::4096].double()
        ref = 1.0 / (1.0 + torch.exp(-xs))
        ok = ref > 1e-37
        rel = ((y[::4096].double() - ref).abs() / ref)[ok]
        if rel.numel():
            worst = max(worst, float(rel.max()))
print(f"monotonicity violations: {bad}; worst relative error vs float64 on the sample: {worst:.3e}")
sys.exit(1 if bad else 0)
