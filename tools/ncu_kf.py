#!/usr/bin/env python3
"""A few launches of KF alone (lp_detect_filter_f32 on the serial entry's CTA count) for
    ncu --set full --clock-control none --import-source on -k regex:levels_filter_tma -s 3 -c 1 \\
        -o gpurun_out/kf python tools/ncu_kf.py [B] [img] [conf]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_lp_b200 import _abi, synth
from yolo_lp_b200.head import PostprocessPlan

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
img = int(sys.argv[2]) if len(sys.argv) > 2 else 640
conf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
dev = torch.device("cuda:0")
levels = synth.synth_levels(B, img, img, dev, seed=0)
plan = PostprocessPlan(levels, (8, 16, 32), 300)
sms = torch.cuda.get_device_properties(dev).multi_processor_count
plan.opts = _abi.opts(filter_ctas=sms - min(B, sms // 6))
for _ in range(5):
    plan.run_filter(conf)
    torch.cuda.synchronize()
print("ok")
