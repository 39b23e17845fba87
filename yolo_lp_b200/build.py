"""In-tree build of ``liblpnms.so`` (hand-written sm_100a CUDA behind a C ABI).

``python -m yolo_lp_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU; the resulting ``.so`` sits next to this file (git-ignored, but it
travels to the GPU box with the gpurun snapshot) and links cudart statically so it
does not care which CUDA runtime torch bundles.

Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   Blackwell only, SASS only
  -fmad=false                               no FMA contraction: kept sets are bit-exact
  -lineinfo                                 ncu source page maps to these files
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "filter.cu", "filter_half.cu", "nms.cu", "decode.cu", "decode_tma.cu", "fused.cu", "fused_tma.cu", "geometry.cu"]
HEADERS = ["common.cuh", "kernels.cuh", "level_tiles.cuh", os.path.join("..", "..", "include", "lpnms.h")]
LIB = os.path.join(HERE, "liblpnms.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "--cudart", "static", "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; liblpnms.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile ``liblpnms.so`` if missing or older than its sources; return its path."""
    if not force and not is_stale():
        return LIB
    extra = os.environ.get("LPNMS_NVCC_EXTRA", "").split()   # e.g. -DLP_NMS_PROFILE for tools/nms_phase_timing.py
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-o", LIB + ".tmp"] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(LIB + ".tmp", LIB)
    log = os.path.join(HERE, "build_ptxas.log")
    with open(log, "w") as f:
        f.write(proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
