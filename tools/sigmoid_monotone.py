#!/usr/bin/env python3
"""Exhaustive check on the GPU that the decode kernel's sigmoid is monotone non-decreasing over
every finite fp32 input (needed for max(sigmoid(x_j)) == sigmoid(max x_j) in the fused path), and
its worst relative error against float64 on a dense sample."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_lp_b200 import _abi

dev = torch.device("cuda:0")
CH = 1 << 26
stream = torch.cuda.current_stream(dev).cuda_stream


def sig(x):
    y = torch.empty_like(x)
    _abi.call("lp_debug_sigmoid_f32", x.data_ptr(), x.numel(), y.data_ptr(), stream)
    return y


bad = 0
worst = 0.0
# positive floats: bit patterns 0 .. 0x7f7fffff ascending; negative floats: 0x80000000.. descending in value
for sign in (0, 1):
    prev_last = None
    for lo in range(0, 0x7f800000, CH):
        hi = min(lo + CH, 0x7f800000)
        bits = torch.arange(lo, hi, dtype=torch.int64, device=dev)
        if sign:
            bits = bits + 0x80000000
        x = bits.to(torch.int32 if not sign else torch.int64)
        x = (bits & 0xffffffff).to(torch.uint32).view(torch.float32) if hasattr(torch, "uint32") else None
        y = sig(x)
        d = y[1:] - y[:-1]
        # positive side: x ascending -> y must not decrease; negative side: x descending -> y must not increase
        viol = (d < 0) if not sign else (d > 0)
        bad += int(viol.sum())
        if prev_last is not None:
            bad += int((y[0] < prev_last) if not sign else (y[0] > prev_last))
        prev_last = y[-1].clone()
        if lo % (CH * 8) == 0:
            xs = x[:: 4096].double()
            ref = 1.0 / (1.0 + torch.exp(-xs))
            ok = ref > 1e-37
            rel = ((y[:: 4096].double() - ref).abs() / ref)[ok]
            if rel.numel():
                worst = max(worst, float(rel.max()))
print(f"monotonicity violations: {bad}; worst relative error vs float64 on the sample: {worst:.3e}")
sys.exit(1 if bad else 0)
