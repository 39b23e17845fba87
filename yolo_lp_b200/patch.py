"""Monkey-patch an importable YOLO-LP checkout so its drivers use this library.

``install()`` rebinds, in the reference's own modules, the names its callers
resolved at import time:

  yolov6.utils.nms.non_max_suppression        + the copies imported by name in
  yolov6.core.inferer (inferer.py:20) and yolov6.core.evaler (evaler.py:16)
  yolov6.core.inferer.Inferer.rescale         (staticmethod, inferer.py:203)
  yolov6.models.effidehead.Detect.forward     (eval branch only, effidehead.py:214)
  yolov6.core.evaler.Evaler.eval              (the LP metric, evaler.py:153; only if the module imports)

``uninstall()`` puts the originals back.  Training code paths (losses, assigners, the
train branch of ``Detect.forward`` and of ``generate_anchors``) are left untouched, and
every patched callable hands anything outside the B200 path back to the original it
replaced:

  * ``Detect.forward``: training mode, CPU input, ``use_dfl=True`` heads (the upstream
    yolov6m/l configs: softmax + ``proj_conv``, effidehead.py:248-250) and a
    ``grid_cell_offset`` other than 0.5 run the reference's own forward;
  * ``Inferer.rescale``: CPU tensors run the reference's own (in-place) arithmetic.
``non_max_suppression`` needs no such branch: a CPU prediction is streamed through the
GPU kernels and comes back as CPU rows (yolo_lp_b200.host).
"""
from __future__ import annotations

import importlib

_originals: list = []   # (owner object, attribute name, original value) in install order


def _rebind(owner, name: str, value) -> None:
    _originals.append((owner, name, owner.__dict__[name] if name in getattr(owner, "__dict__", {}) else getattr(owner, name)))
    setattr(owner, name, value)


def head_runs_on_b200(detect, x) -> bool:
    """Whether the patched ``Detect.forward`` takes the fused decode kernel for this call."""
    return bool(not detect.training and x[0].is_cuda and not getattr(detect, "use_dfl", False)
                and float(getattr(detect, "grid_cell_offset", 0.5)) == 0.5)


def install(nms: bool = True, rescale: bool = True, detect: bool = True, evaler: bool = True) -> list:
    """Returns the list of patched ``module.attribute`` names."""
    from . import head, inferer as _rescale, nms as _nms
    if _originals:
        uninstall()
    done = []
    if nms:
        for modname in ("yolov6.utils.nms", "yolov6.core.inferer", "yolov6.core.evaler"):
            try:
                mod = importlib.import_module(modname)
            except Exception:  # evaler needs pycocotools; skip what cannot be imported
                continue
            if hasattr(mod, "non_max_suppression"):
                _rebind(mod, "non_max_suppression", _nms.non_max_suppression)
                done.append(modname + ".non_max_suppression")
    if rescale:
        try:
            inferer = importlib.import_module("yolov6.core.inferer")
            original_rescale = inferer.Inferer.__dict__["rescale"].__func__

            def rescale_(ori_shape, boxes_and_cors, target_shape):
                if not boxes_and_cors.is_cuda:
                    return original_rescale(ori_shape, boxes_and_cors, target_shape)
                return _rescale.rescale(ori_shape, boxes_and_cors, target_shape)

            _rebind(inferer.Inferer, "rescale", staticmethod(rescale_))
            done.append("yolov6.core.inferer.Inferer.rescale")
        except Exception:
            pass
    if detect:
        try:
            eff = importlib.import_module("yolov6.models.effidehead")
            original = eff.Detect.forward

            def forward(self, x):
                if not head_runs_on_b200(self, x):
                    return original(self, x)
                return head.detect_forward_eval(self, x)

            _rebind(eff.Detect, "forward", forward)
            done.append("yolov6.models.effidehead.Detect.forward")
        except Exception:
            pass
    if evaler:
        try:
            ev = importlib.import_module("yolov6.core.evaler")   # needs pycocotools at import time
            from .evaler import lp_eval

            def eval_(self, preds, targets, model, task):
                self.eval_speed(task)                              # evaler.py:155, unchanged
                return lp_eval(preds, targets)

            _rebind(ev.Evaler, "eval", eval_)
            done.append("yolov6.core.evaler.Evaler.eval")
        except Exception:
            pass
    return done


def uninstall() -> None:
    """Restore every name ``install()`` rebound."""
    while _originals:
        owner, name, value = _originals.pop()
        setattr(owner, name, value)
