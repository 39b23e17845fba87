#!/usr/bin/env python3
"""Time lp::decode_kernel (Detect eval tail) on a BASELINE shape and report achieved HBM bandwidth.

    python tools/decode_bench.py [B] [img]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolo_lp_b200 as lp
from yolo_lp_b200 import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
img = int(sys.argv[2]) if len(sys.argv) > 2 else 640
dev = torch.device("cuda:0")
names, widths = ("pro", "alp", "ad0", "ad1", "ad2", "ad3", "ad4", "ad5"), (31, 24, 37, 37, 37, 37, 37, 37)
levels = []
for h, w in synth.level_shapes(img, img):
    lv = {n: torch.randn(B, c, h, w, device=dev) * 2 - 3 for n, c in zip(names, widths)}
    lv["reg"], lv["cor"] = torch.rand(B, 4, h, w, device=dev) * 6, torch.rand(B, 8, h, w, device=dev) * 6 - 1
    levels.append(lv)
A = sum(h * w for h, w in synth.level_shapes(img, img))
out = torch.empty((B, A, 290), device=dev)
from yolo_lp_b200 import _abi
from yolo_lp_b200.head import DecodePlan
plan = DecodePlan(levels, (8, 16, 32), out)
if os.environ.get("DEC_TMA") == "0":
    plan.opts = _abi.opts(no_tma=True)     # force the cp.async load path
for _ in range(5):
    plan.run()
torch.cuda.synchronize()
K = 50
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(K):
    plan.run()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / K
algo = B * A * (289 + 290) * 4
peak = 6529.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
if os.environ.get("DEC_PROFILE"):   # library built with LPNMS_NVCC_EXTRA=-DLP_DEC_PROFILE
    buf = torch.zeros((148, 3, 4), dtype=torch.int64, device=dev)
    plan.opts = _abi.opts(no_tma=os.environ.get("DEC_TMA") == "0", timing=buf.data_ptr())
    plan.run()
    torch.cuda.synchronize()
    tiles = B * sum((h * w + 31) // 32 for h, w in synth.level_shapes(img, img)) / 148
    t = (buf.double().mean(0) / tiles).tolist()
    print("cycles per tile  load producer: other %.0f wait-empty %.0f issue %.0f" % tuple(t[0][:3]))
    print("                 store producer: other %.0f wait-ofull %.0f issue %.0f wait-read %.0f" % tuple(t[1]))
    print("                 consumer t0: wait-oempty %.0f wait-full %.0f transposition %.0f release %.0f" % tuple(t[2]))
print(json.dumps({"kernel": "lp::decode_tma_kernel", "B": B, "img": img, "A": A, "ms": ms, "algorithmic_bytes": algo,
                  "achieved_gbs": algo / ms / 1e6, "frac_of_measured_hbm": algo / ms / 1e6 / peak,
                  "images_per_s": B / ms * 1e3}))
