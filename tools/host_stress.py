#!/usr/bin/env python3
"""Race hunt for the host-buffer path (chunked H2D on a copy stream overlapped with the kernels of the
previous chunk): alternating pinned inputs, every result compared with the device path's.

    python tools/host_stress.py [iterations]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolo_lp_b200 as lp
from yolo_lp_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cfg = synth.CONFIGS[2]
bad = 0
for dtype in (torch.float32, torch.float16):
    hosts, refs = [], []
    for s in (1, 2, 3):
        h = synth.synth_head(40, cfg["A"], 640, 24, 300, seed=70 + s, pin_memory=True)   # 40 images: a ragged last chunk
        if dtype == torch.float16:
            hh = torch.empty(h.shape, dtype=torch.float16, pin_memory=True)
            hh.copy_(h)
            h = hh
        hosts.append(h)
        refs.append([r.cpu() for r in lp.non_max_suppression(h.cuda(), 0.25, 0.45)])
    for i in range(N):
        got = lp.non_max_suppression(hosts[i % 3], 0.25, 0.45)
        want = refs[i % 3]
        if len(got) != len(want) or any(not torch.equal(g, w) for g, w in zip(got, want)):
            bad += 1
    print(f"host path {dtype}: {N} calls, mismatching calls so far: {bad}")
sys.exit(1 if bad else 0)
