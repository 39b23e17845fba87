#!/usr/bin/env python3
"""Fused path: KF alone / serial step / two-stream pipelined step vs the KF CTA limit (tuning aid).

    python tools/fused_pipeline_sweep.py [B] [img] [conf] [iou] [cta,cta,...]   (0 = the library's own policy)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_lp_b200 import _abi, synth
from yolo_lp_b200.head import PostprocessPlan, PostprocessPipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
img = int(sys.argv[2]) if len(sys.argv) > 2 else 640
conf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
iou = float(sys.argv[4]) if len(sys.argv) > 4 else 0.45
CTAS = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0, 148, 132, 124, 120, 116, 112, 108, 100]
K = 100
dev = torch.device("cuda:0")
levels = synth.synth_levels(B, img, img, dev, seed=1)


def timed(fn, k=K):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k * 1e3


for ctas in CTAS:
    plans = [PostprocessPlan(levels, (8, 16, 32), 300) for _ in range(2)]
    for pl in plans:
        pl.opts = _abi.opts(filter_ctas=ctas)
    t_kf = timed(lambda: plans[0].run_filter(conf))
    t_s = timed(lambda: plans[0].run(conf, iou))
    pipe = PostprocessPipeline(plans)

    def burst():
        pipe.start()
        for _ in range(K):
            pipe.submit(conf, iou)
        pipe.finish()
    t_p = timed(burst, 3) / K
    print(f"KF ctas={ctas or 'auto':>4}: KF alone {t_kf:6.1f} us  serial step {t_s:6.1f} us  pipelined step {t_p:6.1f} us ({B / t_p * 1e6:8.0f} img/s)")
