#!/usr/bin/env python3
"""A few launches of K1 (lp_nms_filter_f32) and K2 on one BASELINE config, for
    ncu --set full --clock-control none --import-source on -k regex:'filter_kernel|nms_kernel' -s 4 -c 2 \\
        -o gpurun_out/k1_cfgN python tools/ncu_k1.py N
The DRAM bytes of the captured K1 launch go into profiles/roofline_traffic.json (bench.py's roofline.traffic).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_lp_b200 import synth
from yolo_lp_b200.nms import NmsPlan

cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = synth.CONFIGS[cid]
B = cfg["B"] if cid != 3 else 32
dev = torch.device("cuda:0")
unique = min(B, 16)
pred = synth.synth_head(unique, cfg["A"], cfg["img"], cfg["n_plates"], cfg["n_pos"], cfg["seed"]).to(dev)
if unique < B:
    pred = pred.repeat((B + unique - 1) // unique, 1, 1)[:B].contiguous()
plan = NmsPlan(B, cfg["A"], cfg["max_det"], dev)
for _ in range(4):
    plan.run_filter(pred, cfg["conf"])
    plan.run_suppress(pred, cfg["iou"])
    torch.cuda.synchronize()
print("ok", cid, B, int(plan.candidate_counts().sum()))
