#!/usr/bin/env python3
"""Host-side cost per call of the entry points (perf_counter round the enqueue, tiny shapes so the GPU is
never the limit): what a caller without a CUDA graph pays per step, and what the per-call tensor-map
encoding of the TMA kernels adds (~1.5 us for 24 maps)."""
import sys, time, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_lp_b200 import synth, _abi
from yolo_lp_b200.head import PostprocessPlan, PostprocessPipeline, DecodePlan
from yolo_lp_b200.nms import NmsPlan, NmsPipeline
dev = torch.device("cuda:0")
levels = synth.synth_levels(32, 640, 640, dev, seed=1)
plans = [PostprocessPlan(levels, (8, 16, 32), 300) for _ in range(2)]
pred = synth.synth_head(32, 8400, 640, 24, 300, 1).to(dev)
nplan = NmsPlan(32, 8400, 300, dev)
def host_us(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    dt = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    return dt
print("NmsPlan.run_filter host us (queue may back up -> includes blocking):", host_us(lambda: nplan.run_filter(pred, 0.25), 50))
lsu = PostprocessPlan(levels, (8, 16, 32), 300); lsu.opts = _abi.opts(no_tma=True)
# tiny problem so that the GPU is never the limit: 1 image 64x64
small = synth.synth_levels(1, 64, 64, dev, seed=1)
sp = PostprocessPlan(small, (8, 16, 32), 300)
sl = PostprocessPlan(small, (8, 16, 32), 300); sl.opts = _abi.opts(no_tma=True)
sd = DecodePlan(small, (8, 16, 32))
print("fused filter (TMA maps encoded per call) host us:", host_us(lambda: sp.run_filter(0.25)))
print("fused filter (no TMA)                    host us:", host_us(lambda: sl.run_filter(0.25)))
print("fused run (filter+suppress)              host us:", host_us(lambda: sp.run(0.25, 0.45)))
print("decode (TMA)                             host us:", host_us(sd.run))
pipe = PostprocessPipeline([PostprocessPlan(small, (8, 16, 32), 300) for _ in range(2)])
pipe.start()
print("fused pipelined submit                   host us:", host_us(lambda: pipe.submit(0.25, 0.45)))
pipe.finish()
sn = NmsPlan(1, 84, 300, dev); spred = torch.rand(1, 84, 290, device=dev)
print("nms run                                  host us:", host_us(lambda: sn.run(spred, 0.25, 0.45)))
