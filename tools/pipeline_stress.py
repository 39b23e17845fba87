#!/usr/bin/env python3
"""Long-running race hunt for the pipelined entries (two alternating filter streams, re-armed
workspaces): N submissions cycling through three different batches; every result is compared on the
device with the serial path's result for that batch, mismatches are counted without synchronising.

    python tools/pipeline_stress.py [submissions] [config id]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_lp_b200 import synth
from yolo_lp_b200.head import PostprocessPlan, PostprocessPipeline
from yolo_lp_b200.nms import NmsPlan, NmsPipeline

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
cid = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = synth.CONFIGS[cid]
B, A, dev = min(cfg["B"], 32), cfg["A"], torch.device("cuda:0")
conf, iou, md = cfg["conf"], cfg["iou"], cfg["max_det"]
bad = torch.zeros((), dtype=torch.int64, device=dev)
side = torch.cuda.Stream(dev)
junk = torch.rand(1 << 22, device=dev)


def perturb(i):
    """Foreign work on a third stream: kernels of varying size that move onto whatever SMs are free."""
    with torch.cuda.stream(side):
        n = 1 << (10 + i % 13)
        junk[:n].mul_(1.0001).add_(0.5)


def mark(out, counts, ref):
    ro, rc = ref
    live = torch.arange(out.shape[1], device=dev)[None, :, None] < rc[:, None, None]
    bad.add_(((out != ro) & live).any().long() + (counts != rc).any().long())


for dtype in (() if os.environ.get("SKIP_NMS") else (torch.float32, torch.float16)):
    preds = [synth.synth_head(B, A, cfg["img"], cfg["n_plates"], cfg["n_pos"], seed=90 + i).to(dev).to(dtype) for i in range(3)]
    plan = NmsPlan(B, A, md, dev)
    refs = []
    for p in preds:
        o, c = plan.run(p, conf, iou)
        refs.append((o.clone(), c.clone()))
    pipe = NmsPipeline(B, A, md, dev)
    pipe.start()
    for i in range(N):
        slot, out, counts = pipe.submit(preds[i % 3], conf, iou)
        perturb(i)
        # compare on the NMS stream, right behind this step's K2 and before the slot is reused
        with torch.cuda.stream(pipe.s_nms):
            mark(out, counts, refs[i % 3])
    pipe.finish()
    torch.cuda.synchronize()
    print(f"NmsPipeline {dtype}: {N} submissions, mismatching steps: {int(bad)}")

levels = [synth.synth_levels(B, cfg["img"], cfg["img"], dev, seed=s) for s in (1, 2)]
for half in (False, True):
    lv = [[{k: (v.half() if half else v) for k, v in l.items()} for l in ls] for ls in levels]
    plans = [PostprocessPlan(l, (8, 16, 32), md) for l in lv]
    refs = []
    for pl in plans:
        o, c = pl.run(conf, iou)
        refs.append((o.clone(), c.clone()))
    pipe = PostprocessPipeline(plans)
    pipe.start()
    for i in range(N):
        slot, out, counts = pipe.submit(conf, iou)
        perturb(i)
        with torch.cuda.stream(pipe.s_nms):
            mark(out, counts, refs[slot])
    pipe.finish()
    torch.cuda.synchronize()
    print(f"PostprocessPipeline half={half}: {N} submissions, mismatching steps so far: {int(bad)}")
# the warp-specialised decode kernel, same treatment
from yolo_lp_b200.head import DecodePlan
dec = DecodePlan(levels[0], (8, 16, 32))
ref = dec.run().clone()
for i in range(N):
    out = dec.run()
    bad.add_((out.view(torch.int32) != ref.view(torch.int32)).any().long())
    perturb(i)
torch.cuda.synchronize()
print(f"decode: {N} launches, mismatching steps so far: {int(bad)}")
sys.exit(1 if int(bad) else 0)
