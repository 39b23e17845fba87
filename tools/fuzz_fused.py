#!/usr/bin/env python3
"""Fused raw-levels path vs decode -> K1 -> K2 on random shapes (TMA and non-TMA ones), thresholds and
logit quantisations (mass ties in sigmoid space); both must agree bit for bit.

    python tools/fuzz_fused.py 0 200
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import yolo_lp_b200 as lp
from yolo_lp_b200 import synth

lo, hi = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
bad = 0
for seed in range(lo, hi):
    rng = np.random.default_rng(seed)
    B = int(rng.integers(1, 9))
    H = int(rng.choice([96, 160, 224, 320, 352, 416, 608, 640]))
    W = int(rng.choice([96, 160, 224, 320, 352, 416, 608, 640]))
    conf = float(rng.choice([0.0, 0.02, 0.1, 0.25]))
    iou = float(rng.choice([0.3, 0.45, 0.65]))
    max_det = int(rng.choice([7, 50, 300]))
    quant = float(rng.choice([0, 0, 0.5, 0.125]))
    levels = synth.synth_levels(B, H, W, dev, seed=seed, pos_frac=float(rng.choice([0.0, 0.03, 0.3])))
    if quant:
        for lv in levels:
            for k in lv:
                if k not in ("reg", "cor"):
                    lv[k] = torch.round(lv[k] / quant) * quant
    half = bool(rng.integers(0, 2))   # fp16 level tensors: fused path on the halves, unfused chain on the upcast copy
    if half:
        h16 = [{k: v.half() for k, v in lv.items()} for lv in levels]
        levels = [{k: v.float() for k, v in lv.items()} for lv in h16]
    fused = lp.detect_postprocess(h16 if half else levels, (8, 16, 32), conf, iou, max_det)
    unfused = lp.non_max_suppression(lp.detect_decode(levels, (8, 16, 32)), conf, iou, max_det=max_det)
    for b, (f, u) in enumerate(zip(fused, unfused)):
        if not torch.equal(f, u):
            bad += 1
            print(f"MISMATCH seed {seed} image {b}: B{B} {H}x{W} conf{conf} iou{iou} md{max_det} q{quant} half{half}")
print(f"seeds {lo}..{hi - 1}: {bad} mismatching images")
sys.exit(1 if bad else 0)
