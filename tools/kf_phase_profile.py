#!/usr/bin/env python3
"""Per-role cycle profile of lp::levels_filter_tma_kernel (library built with
LPNMS_NVCC_EXTRA=-DLP_KF_PROFILE):  python tools/kf_phase_profile.py B img conf"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_lp_b200 import synth, _abi
from yolo_lp_b200.head import PostprocessPlan
B, img, conf = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
dev = torch.device("cuda:0")
levels = synth.synth_levels(B, img, img, dev, seed=0)
plan = PostprocessPlan(levels, (8, 16, 32), 300)
for _ in range(3): plan.run_filter(conf)
buf = torch.zeros((148, 23, 4), dtype=torch.int64, device=dev)
plan.opts = _abi.opts(timing=buf.data_ptr())
plan.run_filter(conf); torch.cuda.synchronize()
plan.opts = None
used = buf[buf.sum((1, 2)) > 0].double()
tiles = B * sum((h * w + 31) // 32 for h, w in synth.level_shapes(img, img)) / used.shape[0]
print("CTAs", used.shape[0], "tiles/CTA %.1f" % tiles, "total cycles/tile/CTA %.0f" % (used[:, 0].sum(1).mean() / tiles))
m = used.mean(0) / tiles
print("scanner (per tile): other %.0f wait-full %.0f scan %.0f" % tuple(m[2, :3].tolist()))
print("finisher (per own tile): other %.0f wait-bar %.0f finish %.0f" % tuple((m[16, :3] * 5).tolist()))
print("producer (per own tile): other %.0f wait-empty %.0f issue %.0f" % tuple((m[21, :3] * 5 / 3).tolist()))
