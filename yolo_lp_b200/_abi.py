"""ctypes binding of ``liblpnms.so`` (declared in ``include/lpnms.h``).

There is no CPU or PyTorch fallback: if the shared library is missing or a call
fails, this module raises.  PyTorch is used only by the callers for device
memory and streams; nothing here takes a torch type.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# LPNMS_LIB: another build of the same ABI (the -DLP_KF_ASSERT library of tests/test_gpu_kf_assert.py)
LIB_PATH = os.environ.get("LPNMS_LIB") or os.path.join(HERE, "liblpnms.so")

ROW = 290
OUT = 28
MAX_LEVELS = 4
MAX_NMS = 30000  # yolov6/utils/nms.py:62

LP_E_THRESHOLD = -5


class LpLevel(ctypes.Structure):
    """``lp_level_t``"""
    _fields_ = [("cls", c_void_p * 8), ("reg", c_void_p), ("cor", c_void_p),
                ("h", c_int), ("w", c_int), ("stride", c_float)]


class LpOpts(ctypes.Structure):
    """``lp_opts_t``: per-call tuning / debug knobs (NULL = production defaults)."""
    _fields_ = [("filter_ctas", c_int), ("no_tma", c_int), ("timing", c_void_p)]


def opts(filter_ctas: int = 0, no_tma: bool = False, timing: int | None = None) -> LpOpts:
    """``timing``: device address of an int64 ``[B,16]`` buffer (``tensor.data_ptr()``)."""
    return LpOpts(int(filter_ctas), int(bool(no_tma)), timing)


#: name -> (restype, argtypes); must list every LP_API symbol of include/lpnms.h
SIGNATURES = {
    "lp_version": (c_int, []),
    "lp_error_string": (c_char_p, [c_int]),
    "lp_nms_workspace_bytes": (c_int, [c_int, c_int, c_int, POINTER(c_size_t)]),
    "lp_nms_f32": (c_int, [c_void_p, c_int, c_int, c_double, c_double, c_int, c_int, c_void_p, c_size_t,
                           c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, POINTER(LpOpts)]),
    "lp_nms_pipelined_f32": (c_int, [c_void_p, c_int, c_int, c_double, c_double, c_int, c_int, c_void_p, c_size_t,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, POINTER(LpOpts)]),
    "lp_nms_filter_f32": (c_int, [c_void_p, c_int, c_int, c_double, c_void_p, c_size_t, c_void_p, POINTER(LpOpts)]),
    "lp_nms_suppress_f32": (c_int, [c_void_p, c_int, c_int, c_double, c_int, c_int, c_void_p, c_size_t,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, POINTER(LpOpts)]),
    "lp_nms_f16": (c_int, [c_void_p, c_int, c_int, c_double, c_double, c_int, c_int, c_void_p, c_size_t,
                           c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, POINTER(LpOpts)]),
    "lp_nms_pipelined_f16": (c_int, [c_void_p, c_int, c_int, c_double, c_double, c_int, c_int, c_void_p, c_size_t,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, POINTER(LpOpts)]),
    "lp_nms_filter_f16": (c_int, [c_void_p, c_int, c_int, c_double, c_void_p, c_size_t, c_void_p, POINTER(LpOpts)]),
    "lp_nms_suppress_f16": (c_int, [c_void_p, c_int, c_int, c_double, c_int, c_int, c_void_p, c_size_t,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, POINTER(LpOpts)]),
    "lp_debug_sigmoid_f32": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p]),
    "lp_detect_decode_f32": (c_int, [POINTER(LpLevel), c_int, c_int, c_void_p, c_void_p, POINTER(LpOpts)]),
    "lp_detect_decode_half_scores_f32": (c_int, [POINTER(LpLevel), c_int, c_int, c_void_p, c_void_p, POINTER(LpOpts)]),
    "lp_detect_postprocess_f32": (c_int, [POINTER(LpLevel), c_int, c_int, c_double, c_double, c_int, c_int, c_void_p,
                                          c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, POINTER(LpOpts)]),
    "lp_detect_pipelined_f32": (c_int, [POINTER(LpLevel), c_int, c_int, c_double, c_double, c_int, c_int, c_void_p,
                                        c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(LpOpts)]),
    "lp_detect_pipelined_to_host_f32": (c_int, [POINTER(LpLevel), c_int, c_int, c_double, c_double, c_int, c_int, c_void_p,
                                                c_size_t, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(LpOpts)]),
    "lp_detect_pipelined_to_host_f16": (c_int, [POINTER(LpLevel), c_int, c_int, c_double, c_double, c_int, c_int, c_void_p,
                                                c_size_t, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(LpOpts)]),
    "lp_detect_workspace_bytes": (c_int, [c_int, c_int, c_int, POINTER(c_size_t)]),
    "lp_detect_filter_f32": (c_int, [POINTER(LpLevel), c_int, c_int, c_double, c_int, c_void_p, c_size_t, c_void_p, POINTER(LpOpts)]),
    "lp_detect_postprocess_f16": (c_int, [POINTER(LpLevel), c_int, c_int, c_double, c_double, c_int, c_int, c_void_p,
                                          c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, POINTER(LpOpts)]),
    "lp_detect_pipelined_f16": (c_int, [POINTER(LpLevel), c_int, c_int, c_double, c_double, c_int, c_int, c_void_p,
                                        c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(LpOpts)]),
    "lp_detect_filter_f16": (c_int, [POINTER(LpLevel), c_int, c_int, c_double, c_int, c_void_p, c_size_t, c_void_p, POINTER(LpOpts)]),
    "lp_detect_suppress_f32": (c_int, [POINTER(LpLevel), c_int, c_int, c_double, c_int, c_int, c_void_p, c_size_t,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, POINTER(LpOpts)]),
    "lp_generate_anchors_f32": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_float), c_int, c_float,
                                        c_void_p, c_void_p, c_void_p]),
    "lp_dist2bbox_f32": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_int, c_void_p, c_void_p]),
    "lp_dist2cor_f32": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_void_p, c_void_p]),
    "lp_xywh2xyxy_f32": (c_int, [c_void_p, c_longlong, c_longlong, c_void_p, c_longlong, c_void_p]),
    "lp_rescale_f32": (c_int, [c_void_p, c_longlong, c_longlong, c_float, c_float, c_float, c_float, c_float,
                               c_int, c_void_p]),
    "lp_txt_records_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "lp_txt_lines_host": (c_int, [c_void_p, c_longlong, ctypes.c_char_p, c_size_t, POINTER(c_size_t)]),
    "lp_prepare_targets_f32": (c_int, [c_void_p, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "lp_eval_match_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "lp_eval_accumulate_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "lp_rescale_batch_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
}

_OPTS_PTR = POINTER(LpOpts)   # ctypes caches pointer types: this is the object the table above holds
_lib = None


class LpError(RuntimeError):
    def __init__(self, fn: str, code: int, text: str):
        super().__init__(f"{fn} failed with code {code}: {text}")
        self.code = code


def load() -> ctypes.CDLL:
    """Load the CUDA library; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m yolo_lp_b200.build` "
                "(yolo_lp_b200 has no CPU or PyTorch fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(fn: str, code: int) -> None:
    if code == 0:
        return
    text = load().lp_error_string(code).decode()
    if code == LP_E_THRESHOLD:  # the reference's only error path (nms.py:57-58)
        raise AssertionError(text)
    if code < 0:
        raise ValueError(f"{fn}: {text} (code {code})")
    raise LpError(fn, code, text)


def call(fn: str, *args, opts: LpOpts | None = None) -> None:
    """Call ``fn``; entries whose last parameter is ``const lp_opts_t*`` get ``opts`` (NULL by default)."""
    if SIGNATURES[fn][1] and SIGNATURES[fn][1][-1] is _OPTS_PTR:
        args = args + (ctypes.byref(opts) if opts is not None else None,)
    check(fn, getattr(load(), fn)(*args))


def detect_workspace_bytes(B: int, A: int, max_det: int) -> int:
    n = c_size_t(0)
    call("lp_detect_workspace_bytes", B, A, max_det, ctypes.byref(n))
    return int(n.value)


def nms_workspace_bytes(B: int, A: int, max_det: int) -> int:
    n = c_size_t(0)
    call("lp_nms_workspace_bytes", B, A, max_det, ctypes.byref(n))
    return int(n.value)
