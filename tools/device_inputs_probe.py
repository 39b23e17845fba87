#!/usr/bin/env python3
"""What bounds the device-inputs flow (fused path + D2H of the detections + host wait per batch)?
Step time of the eager fused pipeline (a) alone, (b) with the D2H copies queued but nobody waiting,
(c) with the host waiting for every batch two steps behind, (d) the same with 3 / 4 batches in flight."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_lp_b200 import synth
from yolo_lp_b200.head import PostprocessPlan, PostprocessPipeline

dev = torch.device("cuda:0")
B, K = 32, 200
levels = synth.synth_levels(B, 640, 640, dev, seed=1)


def bench(depth, copies, waits):
    plans = [PostprocessPlan(levels, (8, 16, 32), 300) for _ in range(depth)]
    pipe = PostprocessPipeline(plans)
    oh = [torch.empty((B, 300, 28), pin_memory=True) for _ in plans]
    ch = [torch.empty((B,), dtype=torch.int32, pin_memory=True) for _ in plans]
    ev = [torch.cuda.Event() for _ in plans]
    cs = torch.cuda.Stream(dev)

    def run(n):
        infl = []
        pipe.start()
        for _ in range(n):
            if waits and len(infl) == depth - 1:
                ev[infl.pop(0)].synchronize()
            slot = pipe.n % depth
            if copies:
                pipe.submit_to_host(0.25, 0.45, oh[slot], ch[slot], cs, ev[slot])
            else:
                pipe.submit(0.25, 0.45)
            infl.append(slot)
        pipe.finish()
        torch.cuda.synchronize(dev)
    run(20)
    t0 = time.perf_counter()
    run(K)
    return (time.perf_counter() - t0) / K * 1e6


for depth, copies, waits in ((2, False, False), (2, True, False), (2, True, True), (3, True, True), (4, True, True), (4, True, False)):
    print(f"depth {depth} copies {copies} host-waits {waits}: {bench(depth, copies, waits):.1f} us per step")
