"""CPU checks round the staged reference (oracle/_ref, see oracle/stage_ref.py): the staged files are
the reference's bytes, the torch port used when the staged copy is absent equals it, and
``patch.install()`` rebinds / falls back / uninstalls correctly on the real package."""
import os
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

from oracle import stage_ref, torch_port
from yolo_lp_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_SRC = os.path.isdir("/root/reference/yolov6")


@pytest.fixture(scope="module")
def ref():
    if HAVE_SRC:
        stage_ref.stage()
    if not stage_ref.is_staged():
        pytest.skip("oracle/_ref not staged and /root/reference absent")
    return stage_ref.reference()


def test_staged_copy_is_the_reference_byte_for_byte(ref):
    assert stage_ref.verify() == []
    if HAVE_SRC:
        assert stage_ref.verify(src="/root/reference") == []
    assert os.path.abspath(ref.nms.__file__).startswith(os.path.join(ROOT, "oracle", "_ref"))


def test_staged_copy_is_ignored_by_git_but_travels_with_gpurun():
    ignore = open(os.path.join(ROOT, ".gitignore")).read().split()
    assert "oracle/_ref/" in ignore
    gpurunignore = os.path.join(ROOT, ".gpurunignore")
    if os.path.exists(gpurunignore):
        assert "oracle/_ref" not in open(gpurunignore).read()


@pytest.mark.parametrize("cid,idx", [(2, [0, 7]), (4, [3]), (5, [1])])
def test_torch_port_equals_the_staged_reference(ref, cid, idx):
    cfg = synth.CONFIGS[cid]
    pred = torch.cat([synth.synth_head(1, cfg["A"], cfg["img"], cfg["n_plates"], cfg["n_pos"], cfg["seed"], first_index=i)
                      for i in idx])
    want = ref.non_max_suppression(pred.clone(), cfg["conf"], cfg["iou"], max_det=cfg["max_det"])
    got = torch_port.non_max_suppression(pred.clone(), cfg["conf"], cfg["iou"], max_det=cfg["max_det"])
    assert len(want) == len(got)
    for w, g in zip(want, got):
        assert w.shape[0] > 0 and torch.equal(w, g)


def test_head_runs_on_b200_predicate():
    from yolo_lp_b200.patch import head_runs_on_b200
    cuda = [types.SimpleNamespace(is_cuda=True)]
    cpu = [types.SimpleNamespace(is_cuda=False)]
    lp_head = types.SimpleNamespace(training=False, use_dfl=False, grid_cell_offset=0.5)
    assert head_runs_on_b200(lp_head, cuda)
    assert not head_runs_on_b200(lp_head, cpu)
    assert not head_runs_on_b200(types.SimpleNamespace(training=True, use_dfl=False, grid_cell_offset=0.5), cuda)
    # the upstream yolov6m/l heads (reg_max=16: softmax + proj_conv, effidehead.py:248-250) stay with the reference
    assert not head_runs_on_b200(types.SimpleNamespace(training=False, use_dfl=True, grid_cell_offset=0.5), cuda)
    assert not head_runs_on_b200(types.SimpleNamespace(training=False, use_dfl=False, grid_cell_offset=0.0), cuda)


def test_install_rebinds_falls_back_and_uninstalls(ref):
    """INTEGRATION.md route 1 on the real package, in a fresh interpreter: names rebound, a use_dfl=True
    head and CPU tensors still take the reference's own code (same results as unpatched), uninstall()
    restores the originals."""
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from oracle import stage_ref
ref = stage_ref.reference()
import yolov6.utils.nms as n, yolov6.core.inferer as i, yolov6.models.effidehead as e
orig_nms, orig_fwd, orig_rescale = n.non_max_suppression, e.Detect.forward, i.Inferer.rescale
# a use_dfl=True head as the upstream m/l configs build it, and an LP head
ch = [0] * 11
ch[6], ch[8], ch[10] = 16, 32, 64
torch.manual_seed(0)
dfl = e.Detect(31, 24, 37, 3, head_layers=e.build_effidehead_layer(ch, 1, 31, 24, 37, reg_max=16, num_layers=3), use_dfl=True, reg_max=16).eval()
feats = [torch.rand(1, c, 64 // s, 64 // s) for c, s in zip((16, 32, 64), (8, 16, 32))]
with torch.no_grad():
    want_dfl = dfl([f.clone() for f in feats])
rows = torch.rand(5, 12) * 600
want_rows = i.Inferer.rescale((640, 640), rows.clone(), (1080, 1920, 3))

import yolo_lp_b200
from yolo_lp_b200 import patch
done = yolo_lp_b200.install()
from yolo_lp_b200.nms import non_max_suppression as ours
assert n.non_max_suppression is ours and i.non_max_suppression is ours
assert e.Detect.forward.__module__ == "yolo_lp_b200.patch" and i.Inferer.rescale.__module__ == "yolo_lp_b200.patch"
assert {"yolov6.utils.nms.non_max_suppression", "yolov6.core.inferer.Inferer.rescale",
        "yolov6.models.effidehead.Detect.forward"} <= set(done)
with torch.no_grad():
    got_dfl = dfl([f.clone() for f in feats])          # use_dfl (and CPU) -> the reference's own forward
assert torch.equal(got_dfl, want_dfl)
got_rows = i.Inferer.rescale((640, 640), rows.clone(), (1080, 1920, 3))   # CPU rows -> the reference's own rescale
assert torch.equal(got_rows, want_rows)
patch.uninstall()
assert n.non_max_suppression is orig_nms and i.non_max_suppression is orig_nms
assert e.Detect.forward is orig_fwd and i.Inferer.rescale is orig_rescale
print("ok")
''' % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-3000:]
