#!/usr/bin/env python3
"""One warm and one measured launch of the decode kernel and of KF on the cfg2 shape, for
    ncu --set full --clock-control none --import-source on -k regex:'decode_tma|levels_filter_tma' \\
        --launch-skip 2 -c 2 -o gpurun_out/x python tools/ncu_targets.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_lp_b200 import synth
from yolo_lp_b200.head import DecodePlan, PostprocessPlan

B, img = 32, 640
dev = torch.device("cuda:0")
levels = synth.synth_levels(B, img, img, dev, seed=0)
dec = DecodePlan(levels, (8, 16, 32))
fused = PostprocessPlan(levels, (8, 16, 32), 300)
for _ in range(2):
    dec.run()
    fused.run(0.25, 0.45)     # serial entry: KF on #SMs - #SMs/6 CTAs
    torch.cuda.synchronize()
print("ok")
