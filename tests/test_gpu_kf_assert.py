"""-m gpu: the fused-path kernel built with -DLP_KF_ASSERT (yolo_lp_b200/liblpnms_kfassert.so, made by
``yolo_lp_b200.build.build_assert_lib()`` / ``__graft_entry__.build()``): every hand-off between the
producer, scanner and finisher warps of lp::levels_filter_tma_kernel carries the tile number it belongs
to and a warp that finds another tile's tag traps (csrc/fused_tma.cu).  compute-sanitizer is closed on
this GPU pool; this is the in-kernel phase check in its place, run under the schedule perturbations that
exposed the three historical bugs (cold workspace, foreign kernels sharing the SMs)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSERT_LIB = os.path.join(ROOT, "yolo_lp_b200", "liblpnms_kfassert.so")

STRESS = r'''
import sys, torch
sys.path.insert(0, %r)
from yolo_lp_b200 import _abi, synth
from yolo_lp_b200.head import PostprocessPlan, PostprocessPipeline
assert _abi.LIB_PATH.endswith("liblpnms_kfassert.so")
DEV = "cuda:0"
for half in (False, True):
    for (B, img, conf) in ((32, 640, 0.25), (6, 1280, 0.25), (20, 640, 0.001)):
        levels = synth.synth_levels(B, img, img, DEV, seed=B)
        if half:
            levels = [{k: v.half() for k, v in lv.items()} for lv in levels]
        lsu = PostprocessPlan(levels, (8, 16, 32), 300)
        lsu.opts = _abi.opts(no_tma=True)        # the register-resident kernel (no ring): the reference result
        if half:
            lsu = PostprocessPlan([{k: v.float() for k, v in lv.items()} for lv in levels], (8, 16, 32), 300)
            lsu.opts = _abi.opts(no_tma=True)
        ref_out, ref_counts = (t.clone() for t in lsu.run(conf, 0.45))
        live = torch.arange(300, device=DEV)[None, :, None] < ref_counts[:, None, None]
        side = torch.cuda.Stream(DEV)
        rc = torch.randint(0, 300, (32,), device=DEV, dtype=torch.int32)
        bad = torch.zeros((), dtype=torch.int64, device=DEV)
        for poison in (0x00, 0xFF):              # cold / poisoned workspace: slow finishers
            plan = PostprocessPlan(levels, (8, 16, 32), 300)
            plan.workspace.fill_(poison)
            for ctas in (0, 148, 37, 5):         # few CTAs: many ring wraps per CTA
                plan.opts = _abi.opts(filter_ctas=ctas)
                for _ in range(40):
                    out, counts = plan.run(conf, 0.45)
                    bad.add_(((out != ref_out) & live).any().long() + (counts != ref_counts).any().long())
                    with torch.cuda.stream(side):
                        _ = torch.arange(300, device=DEV)[None, :, None] < rc[:, None, None]
        plans = [PostprocessPlan(levels, (8, 16, 32), 300) for _ in range(2)]
        pipe = PostprocessPipeline(plans)
        pipe.start()
        for _ in range(200):
            slot, out, counts = pipe.submit(conf, 0.45)
            with torch.cuda.stream(side):
                _ = torch.arange(300, device=DEV)[None, :, None] < rc[:, None, None]
        pipe.finish()
        bad.add_(((out != ref_out) & live).any().long() + (counts != ref_counts).any().long())
        torch.cuda.synchronize()                 # a trap surfaces here as a CUDA error
        assert int(bad) == 0, (half, B, img, conf, int(bad))
print("ok")
'''


def test_kf_phase_assertions_hold_under_perturbed_schedules():
    if not os.path.exists(ASSERT_LIB):
        pytest.fail("yolo_lp_b200/liblpnms_kfassert.so is missing: __graft_entry__.build() builds it")
    env = dict(os.environ, LPNMS_LIB=ASSERT_LIB)
    out = subprocess.run([sys.executable, "-c", STRESS % ROOT], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and "ok" in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])


def test_assert_library_really_carries_the_checks():
    """The instrumented build differs from the product library exactly by the trap paths."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not on this box")

    def traps(lib):
        sass = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN2lp24levels_filter_tma_kernelILb0EEEvNS_18LevelsFilterParamsENS_10DecodeMapsE", lib],
                              capture_output=True, text=True, timeout=300).stdout
        return sass.count("BPT.TRAP")
    assert traps(ASSERT_LIB) > traps(os.path.join(ROOT, "yolo_lp_b200", "liblpnms.so"))
