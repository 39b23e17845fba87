#!/usr/bin/env python3
"""Serial vs two-stream pipelined step time for one config, sweeping the K1 CTA limit (tuning aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_lp_b200 import _abi, synth
from yolo_lp_b200.nms import NmsPlan, NmsPipeline

cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = 200
cfg = synth.CONFIGS[cid]
B = min(cfg["B"], 64)
dev = torch.device("cuda:0")
pred = synth.synth_head(B, cfg["A"], cfg["img"], cfg["n_plates"], cfg["n_pos"], cfg["seed"]).to(dev)
plan = NmsPlan(B, cfg["A"], cfg["max_det"], dev)


def timed(fn, k=K):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k * 1e3


for ctas in (0, 148, 132, 116, 100, 84):
    plan.opts = _abi.opts(filter_ctas=ctas)
    t_f = timed(lambda: plan.run_filter(pred, cfg["conf"]))
    t_s = timed(lambda: plan.run(pred, cfg["conf"], cfg["iou"]))
    pipe = NmsPipeline(B, cfg["A"], cfg["max_det"], dev)
    for pl in pipe.plans:
        pl.opts = plan.opts
    ref = plan.counts.clone()

    def go():
        pipe.start()
        for _ in range(K):
            pipe.submit(pred, cfg["conf"], cfg["iou"])
        pipe.finish()
    go()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    go()
    b.record()
    torch.cuda.synchronize()
    t_p = a.elapsed_time(b) / K * 1e3
    ok = all(torch.equal(p.counts, ref) for p in pipe.plans)
    print(f"cfg{cid} K1 ctas={ctas or 148:3d}: filter alone {t_f:7.1f} us  serial step {t_s:7.1f} us  pipelined step {t_p:7.1f} us"
          f"  ({B / t_p * 1e6:9.0f} img/s)  counts_ok={ok}")
