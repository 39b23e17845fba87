#!/usr/bin/env python3
"""Extended GPU-vs-oracle fuzz (the parametrised test in tests/ runs 24 seeds; this runs any range):

    python tools/fuzz_extended.py 100 400      # seeds 100..399, fp32 and fp16 storage
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from yolo_lp_b200 import synth
from yolo_lp_b200.nms import non_max_suppression_with_index
from oracle import lp_oracle

lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(lo, hi):
    rng = np.random.default_rng(seed)
    B = int(rng.integers(1, 5))
    big = os.environ.get("BIG") == "1"     # BIG=1: 1280x1280-sized inputs, thousands of candidates
    A = int(rng.choice([8400, 16000, 33600] if big else [33, 315, 2100, 5040, 8400, 12000]))
    n_pos = int(rng.integers(0, min(A, 9000 if big else 3000)))
    conf = float(rng.choice([0.0, 0.01, 0.05, 0.25, 0.6]))
    iou = float(rng.choice([0.0, 0.2, 0.45, 0.65, 1.0]))
    max_det = int(rng.choice([1, 7, 64, 300, 1000]))
    quant = int(rng.choice([0, 0, 4, 16]))
    plates = int(rng.choice([1, 3, 12, 40]))
    pred = synth.synth_head(B, A, 640, plates, n_pos, seed=5000 + seed, quant=quant or None)
    if seed % 3 == 0:
        g = torch.Generator().manual_seed(seed)
        pred[..., 4] = torch.rand(pred.shape[:2], generator=g)
    for half in (False, True):
        x = pred.half() if half else pred
        want, widx = lp_oracle.non_max_suppression(x.float().numpy(), conf, iou, max_det=max_det, return_index=True)
        rows, idx = non_max_suppression_with_index(x.cuda(), conf, iou, max_det)
        for b in range(B):
            ok = np.array_equal(idx[b].cpu().numpy(), widx[b]) and np.array_equal(
                rows[b].cpu().numpy().view(np.uint32), want[b].view(np.uint32))
            if not ok:
                bad += 1
                print(f"MISMATCH seed {seed} half {half} image {b}: B{B} A{A} n_pos{n_pos} conf{conf} iou{iou} md{max_det} q{quant} plates{plates}")
print(f"seeds {lo}..{hi - 1}: {bad} mismatching images")
sys.exit(1 if bad else 0)
