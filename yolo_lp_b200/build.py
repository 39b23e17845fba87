"""In-tree build of ``liblpnms.so`` (hand-written sm_100a CUDA behind a C ABI).

``python -m yolo_lp_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU; the resulting ``.so`` sits next to this file (git-ignored, but it
travels to the GPU box with the gpurun snapshot) and links cudart statically so it
does not care which CUDA runtime torch bundles.

Every ``csrc/*.cu`` is its own translation unit (no relocatable device code: kernels
never call across files), compiled in parallel into ``csrc/_obj/`` and linked once.

Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   Blackwell only, SASS only
  -fmad=false                               no FMA contraction: kept sets are bit-exact
  -lineinfo                                 ncu source page maps to these files
"""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
HEADER = os.path.join(HERE, "..", "include", "lpnms.h")
LIB = os.path.join(HERE, "liblpnms.so")
LOG = os.path.join(OBJ, "ptxas.log")     # registers / spills of every kernel (-Xptxas -v), not tracked

COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "-Xptxas", "-v",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "--cudart", "static", "-shared"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; liblpnms.so cannot be built")


def sources() -> list:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def dependencies() -> list:
    """Everything a translation unit may include: all of csrc/ plus the public header."""
    return sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [HEADER, os.path.abspath(__file__)]


def extra_flags() -> list:
    return os.environ.get("LPNMS_NVCC_EXTRA", "").split()   # e.g. -DLP_NMS_PROFILE for tools/nms_phase_timing.py


def _stamp() -> str:
    """Fingerprint of the build inputs that are not file contents (flags)."""
    return hashlib.sha1(" ".join(COMPILE_FLAGS + extra_flags()).encode()).hexdigest()


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    stamp_file = os.path.join(OBJ, "flags.sha1")
    if not os.path.exists(stamp_file) or open(stamp_file).read().strip() != _stamp():
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in dependencies())


ASSERT_LIB = os.path.join(HERE, "liblpnms_kfassert.so")   # the same library with -DLP_KF_ASSERT (see fused_tma.cu)


def build_assert_lib() -> str:
    return build(force=True, lib=ASSERT_LIB, extra=["-DLP_KF_ASSERT"])


def build(force: bool = False, verbose: bool = False, lib: str = LIB, extra: list | None = None) -> str:
    """Compile ``liblpnms.so`` if missing or older than its sources; return its path."""
    if not force and lib == LIB and not is_stale():
        return lib
    nvcc, extra = nvcc_path(), extra_flags() + list(extra or [])
    os.makedirs(OBJ, exist_ok=True)
    tag = "" if lib == LIB else "." + os.path.basename(lib)

    def compile_one(src: str):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + tag + ".o")
        cmd = [nvcc, *COMPILE_FLAGS, *extra, "-c", src, "-o", obj]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        return obj, proc.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, sources()))
    cmd = [nvcc, *LINK_FLAGS, "-o", lib + ".tmp"] + [o for o, _ in results]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(lib + ".tmp", lib)
    log = "".join(err for _, err in results)
    if lib == LIB:
        with open(LOG, "w") as f:
            f.write(log)
        with open(os.path.join(OBJ, "flags.sha1"), "w") as f:
            f.write(_stamp())
    if verbose:
        print(log)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
