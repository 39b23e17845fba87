"""yolo_lp_b200 -- B200-native (sm_100a) post-processing for YOLO-LP.

Drop-in replacements, same names and signatures as the reference callables
(file:line into KyleHuang9/YOLO-LP):

  non_max_suppression, xywh2xyxy      yolov6/utils/nms.py:31, :21
  generate_anchors                    yolov6/assigners/anchor_generator.py:4 (eval branch)
  dist2bbox, dist2cor                 yolov6/utils/general.py:29, :51
  detect_decode / detect_forward_eval yolov6/models/effidehead.py:214-301 (eval branch)
  rescale                             yolov6/core/inferer.py:203 (Inferer.rescale)

Everything computes in ``liblpnms.so`` (hand-written CUDA behind the C ABI of
``include/lpnms.h``); there is no CPU / PyTorch fallback.  Submodules import
lazily so that ``yolo_lp_b200.synth`` works on hosts without the library.
"""
from __future__ import annotations

__all__ = ["non_max_suppression", "xywh2xyxy", "generate_anchors", "dist2bbox", "dist2cor", "detect_decode",
           "detect_forward_eval", "DetectEval", "rescale", "NmsPlan", "install"]

_LAZY = {
    "non_max_suppression": "nms", "xywh2xyxy": "nms", "NmsPlan": "nms", "non_max_suppression_with_index": "nms",
    "generate_anchors": "head", "dist2bbox": "head", "dist2cor": "head", "detect_decode": "head",
    "detect_forward_eval": "head", "DetectEval": "head", "detect_postprocess": "head", "detect_forward_nms": "head",
    "DecodePlan": "head", "PostprocessPlan": "head",
    "rescale": "inferer", "rescale_batch": "inferer", "rescale_table": "inferer",
    "txt_records": "inferer", "txt_lines": "inferer",
    "lp_eval": "evaler", "eval_counts": "evaler",
    "install": "patch",
}


def __getattr__(name):
    mod = _LAZY.get(name)
    if mod is None:
        raise AttributeError(name)
    import importlib
    return getattr(importlib.import_module(f"{__name__}.{mod}"), name)
