"""Monkey-patch an importable YOLO-LP checkout so its drivers use this library.

``install()`` rebinds, in the reference's own modules, the names its callers
resolved at import time:

  yolov6.utils.nms.non_max_suppression        + the copies imported by name in
  yolov6.core.inferer (inferer.py:20) and yolov6.core.evaler (evaler.py:16)
  yolov6.core.inferer.Inferer.rescale         (staticmethod, inferer.py:203)
  yolov6.models.effidehead.Detect.forward     (eval branch only, effidehead.py:214)
  yolov6.core.evaler.Evaler.eval              (the LP metric, evaler.py:153; only if the module imports)

Training code paths (losses, assigners, the train branch of ``Detect.forward``
and of ``generate_anchors``) are left untouched.
"""
from __future__ import annotations

import importlib


def install(nms: bool = True, rescale: bool = True, detect: bool = True, evaler: bool = True) -> list:
    """Returns the list of patched ``module.attribute`` names."""
    from . import head, inferer as _rescale, nms as _nms
    done = []
    if nms:
        for modname in ("yolov6.utils.nms", "yolov6.core.inferer", "yolov6.core.evaler"):
            try:
                mod = importlib.import_module(modname)
            except Exception:  # evaler needs pycocotools; skip what cannot be imported
                continue
            if hasattr(mod, "non_max_suppression"):
                mod.non_max_suppression = _nms.non_max_suppression
                done.append(modname + ".non_max_suppression")
    if rescale:
        try:
            inferer = importlib.import_module("yolov6.core.inferer")
            inferer.Inferer.rescale = staticmethod(_rescale.rescale)
            done.append("yolov6.core.inferer.Inferer.rescale")
        except Exception:
            pass
    if detect:
        try:
            eff = importlib.import_module("yolov6.models.effidehead")
            original = eff.Detect.forward

            def forward(self, x):
                if self.training or not x[0].is_cuda:
                    return original(self, x)
                return head.detect_forward_eval(self, x)

            eff.Detect.forward = forward
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

            ev.Evaler.eval = eval_
            done.append("yolov6.core.evaler.Evaler.eval")
        except Exception:
            pass
    return done
