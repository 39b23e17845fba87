#!/usr/bin/env python3
"""Pipelined step time as a CUDA graph (what bench.py times) vs K1's CTA count: python tools/graph_cta_sweep.py [cfg]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_lp_b200 import _abi, synth
from yolo_lp_b200.nms import NmsPipeline

cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = synth.CONFIGS[cid]
B = min(cfg["B"], 64)
dev = torch.device("cuda:0")
pred = synth.synth_head(min(B, 16), cfg["A"], cfg["img"], cfg["n_plates"], cfg["n_pos"], cfg["seed"]).to(dev)
pred = pred.repeat((B + 15) // 16, 1, 1)[:B].contiguous()
K = 100
for ctas in (0, 76, 84, 92, 100, 108, 116, 124, 132, 148):
    pipe = NmsPipeline(B, cfg["A"], cfg["max_det"], dev)
    for pl in pipe.plans:
        pl.opts = _abi.opts(filter_ctas=ctas)
    g = pipe.capture(pred, cfg["conf"], cfg["iou"], K)
    g.launch(); g.launch()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        g.launch()
        a.record(); g.launch(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / K * 1e3)
    print(f"cfg{cid} K1 ctas={ctas or 'auto':>4}: {best:6.2f} us per step  {B / best * 1e6:9.0f} img/s")
