"""Stage the UNMODIFIED reference for the GPU box (test infrastructure, never product).

``/root/reference`` exists only in the build container.  The reference is Python, so "building"
it is copying: this recipe copies the reference's importable package (``yolov6/**/*.py`` and the
two LP configs) byte for byte into ``oracle/_ref/`` -- git-ignored (reference sources never enter
this repo's history) but not gpurun-ignored, so the copy travels to the GPU box like the built
``.so``.  ``MANIFEST.json`` records the SHA-256 of every staged file; the CPU suite re-checks the
staged files against ``/root/reference`` whenever that is present.

Who may use it (and only as the checker / the timed CPU baseline): ``tests/``,
``__graft_entry__.smoke()``, ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline``.

    python oracle/stage_ref.py            # (re)stage from /root/reference
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(REF_DST, "MANIFEST.json")
EXTRA = ("configs/yololps.py", "configs/yololpn.py", "LICENSE")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def _wanted(src: str) -> list:
    out = []
    for root, _dirs, files in os.walk(os.path.join(src, "yolov6")):
        for fn in files:
            if fn.endswith(".py"):
                out.append(os.path.relpath(os.path.join(root, fn), src))
    out += [e for e in EXTRA if os.path.exists(os.path.join(src, e))]
    return sorted(out)


def stage(src: str = REF_SRC, dst: str = REF_DST) -> str | None:
    """Copy the reference package into ``oracle/_ref``; returns the path, or None when the
    reference is not on this machine (the GPU box: the staged copy is used as it came)."""
    if not os.path.isdir(os.path.join(src, "yolov6")):
        return dst if is_staged(dst) else None
    manifest = {}
    for rel in _wanted(src):
        target = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(target), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), target)
        manifest[rel] = _sha(target)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=0, sort_keys=True)
    return dst


def is_staged(dst: str = REF_DST) -> bool:
    return os.path.exists(os.path.join(dst, "yolov6", "utils", "nms.py")) and os.path.exists(os.path.join(dst, "MANIFEST.json"))


def verify(dst: str = REF_DST, src: str | None = None) -> list:
    """Files whose staged bytes differ from the manifest (and from ``src`` when given)."""
    bad = []
    files = json.load(open(os.path.join(dst, "MANIFEST.json")))["files"]
    for rel, digest in files.items():
        p = os.path.join(dst, rel)
        if not os.path.exists(p) or _sha(p) != digest:
            bad.append(rel)
        elif src is not None and _sha(os.path.join(src, rel)) != digest:
            bad.append(rel)
    return bad


def reference_path() -> str | None:
    """Where the unmodified reference can be imported from: the staged copy, else None."""
    return REF_DST if is_staged() else None


class Reference:
    """The reference's own callables for this path, imported from the staged copy."""

    def __init__(self):
        path = reference_path()
        if path is None:
            raise RuntimeError("oracle/_ref is not staged: run `python oracle/stage_ref.py` in the build container")
        if path not in sys.path:
            sys.path.insert(0, path)
        self.path = path
        self.nms = importlib.import_module("yolov6.utils.nms")
        self.general = importlib.import_module("yolov6.utils.general")
        self.anchor_generator = importlib.import_module("yolov6.assigners.anchor_generator")
        self.effidehead = importlib.import_module("yolov6.models.effidehead")
        self.inferer = importlib.import_module("yolov6.core.inferer")
        if not os.path.abspath(self.nms.__file__).startswith(os.path.abspath(path)):
            raise RuntimeError(f"yolov6 was imported from {self.nms.__file__}, not from the staged copy")
        self.non_max_suppression = self.nms.non_max_suppression
        self.Detect = self.effidehead.Detect
        self.build_effidehead_layer = self.effidehead.build_effidehead_layer
        self.rescale = self.inferer.Inferer.rescale

    def build_head(self, channels=(64, 128, 256), rerandomise=True, seed=0, cls_scale=8.0, reg_scale=4.0, reg_bias=1.5):
        """An LP Detect head (effidehead.py:15,304) built like tests/golden/make_golden.py does: LP-s widths by
        default, deterministic weights on any host (torch.rand from a seeded generator).  ``rerandomise``
        replaces the zero-initialised prediction convs (the true random init is degenerate: every class
        score 0.01, every distance 1.0) by uniform weights scaled so that class scores spread over
        (0, 1) and neighbouring boxes overlap -- i.e. the filter and the NMS both have work to do."""
        import torch
        g = torch.Generator().manual_seed(seed)
        ch_list = [0] * 11
        ch_list[6], ch_list[8], ch_list[10] = channels
        layers = self.build_effidehead_layer(ch_list, 1, 31, 24, 37, reg_max=0, num_layers=3)
        head = self.Detect(31, 24, 37, 3, head_layers=layers, use_dfl=False, reg_max=0)
        head.initialize_biases()
        with torch.no_grad():
            for p in head.parameters():
                if p.dim() > 1:
                    p.copy_((torch.rand(p.shape, generator=g) - 0.5) * (2.0 / max(1, p[0].numel()) ** 0.5))
            if rerandomise:
                for name in ("pro_preds", "alp_preds", "ad0_preds", "ad1_preds", "ad2_preds", "ad3_preds", "ad4_preds",
                             "ad5_preds"):
                    for conv in getattr(head, name):
                        conv.weight.copy_((torch.rand(conv.weight.shape, generator=g) - 0.5) * cls_scale)
                for conv in head.reg_preds:
                    conv.weight.copy_((torch.rand(conv.weight.shape, generator=g) - 0.5) * reg_scale)
                    conv.bias.fill_(reg_bias)
                for conv in head.cor_preds:
                    conv.weight.copy_((torch.rand(conv.weight.shape, generator=g) - 0.5) * reg_scale)
            else:
                for name in ("pro_preds", "alp_preds", "ad0_preds", "ad1_preds", "ad2_preds", "ad3_preds", "ad4_preds",
                             "ad5_preds", "reg_preds", "cor_preds"):
                    for conv in getattr(head, name):
                        conv.weight.zero_()     # as initialize_biases leaves them (effidehead.py:94-154)
        return head.eval()


_ref = None


def reference() -> Reference:
    global _ref
    if _ref is None:
        _ref = Reference()
    return _ref


if __name__ == "__main__":
    where = stage()
    print("staged:", where, "files:", len(json.load(open(MANIFEST))["files"]) if where else 0)
