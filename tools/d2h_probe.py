#!/usr/bin/env python3
"""Device->host copy time of one batch's detections (out[32,300,28] fp32 = 1.07 MB, counts[32]) into pinned
memory: what the D2H leg of the device-inputs flow costs per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = torch.device("cuda:0")
out = torch.rand(32, 300, 28, device=dev)
cnt = torch.zeros(32, dtype=torch.int32, device=dev)
oh = torch.empty(out.shape, pin_memory=True)
ch = torch.empty(cnt.shape, dtype=torch.int32, pin_memory=True)
s = torch.cuda.Stream(dev)
def once():
    with torch.cuda.stream(s):
        ch.copy_(cnt, non_blocking=True)
        oh.copy_(out, non_blocking=True)
for _ in range(10): once()
s.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(s):
    a.record(s)
    for _ in range(100): once()
    b.record(s)
s.synchronize()
print("D2H of out (1.075 MB) + counts, device time per step: %.1f us" % (a.elapsed_time(b) * 10))
big = torch.rand(64 << 20, device=dev); bh = torch.empty(big.shape, pin_memory=True)
with torch.cuda.stream(s):
    bh.copy_(big, non_blocking=True); a.record(s); bh.copy_(big, non_blocking=True); b.record(s)
s.synchronize()
print("D2H 256 MB: %.1f GB/s" % (big.numel() * 4 / a.elapsed_time(b) / 1e6))
