"""``Inferer.rescale`` (``yolov6/core/inferer.py:203-228``) on B200.

``rescale(ori_shape, boxes_and_cors, target_shape)`` maps the 12 box/corner
coordinates of each detection from the letterboxed input back to the source
image, IN PLACE, and returns its argument -- exactly the reference's contract.
Ratio and padding are computed here in Python doubles like the reference
(:206-207) and handed to the kernel rounded to fp32; the kernel does the fp32
subtract, TRUE division and clamp (:210-225).
"""
from __future__ import annotations

import torch

from . import _abi


def rescale_params(ori_shape, target_shape):
    """(pad_x, pad_y, ratio, W0, H0) as Python floats (inferer.py:206-207)."""
    ratio = min(ori_shape[0] / target_shape[0], ori_shape[1] / target_shape[1])
    pad_x = (ori_shape[1] - target_shape[1] * ratio) / 2
    pad_y = (ori_shape[0] - target_shape[0] * ratio) / 2
    return pad_x, pad_y, ratio, float(target_shape[1]), float(target_shape[0])


def rescale(ori_shape, boxes_and_cors, target_shape, do_round=False):
    '''Rescale the output to the original image shape (in place).  ``do_round=True`` also
    applies the caller's ``.round()`` of inferer.py:100 in the same launch.'''
    t = boxes_and_cors
    if not isinstance(t, torch.Tensor) or t.device.type != "cuda":
        raise RuntimeError("yolo_lp_b200.rescale needs a CUDA tensor (no CPU fallback)")
    if t.dim() != 2 or t.shape[1] != 12:
        raise ValueError(f"boxes_and_cors must be [k, 12], got {tuple(t.shape)}")
    if t.shape[0] == 0:
        return t
    pad_x, pad_y, ratio, w0, h0 = rescale_params(ori_shape, target_shape)
    direct = t.dtype == torch.float32 and t.stride(1) == 1 and t.stride(0) >= 12
    work = t if direct else t.float().contiguous()
    with torch.cuda.device(work.device):
        _abi.call("lp_rescale_f32", work.data_ptr(), work.shape[0], work.stride(0), pad_x, pad_y, ratio, w0, h0,
                  int(bool(do_round)), torch.cuda.current_stream(work.device).cuda_stream)
    if not direct:
        t.copy_(work)
    return t


def rescale_batch(det, counts, ori_shapes, target_shapes, do_round=True):
    """One launch for a whole batch: ``det[B,max_det,28]``, ``counts[B]`` (device int32),
    per-image letterboxed and source shapes.  Columns 0..11 are rescaled in place."""
    B = det.shape[0]
    params = torch.tensor([rescale_params(o, s) for o, s in zip(ori_shapes, target_shapes)],
                          dtype=torch.float64).to(torch.float32).to(det.device)
    with torch.cuda.device(det.device):
        _abi.call("lp_rescale_batch_f32", det.data_ptr(), counts.data_ptr(), B, det.shape[1], params.data_ptr(),
                  int(bool(do_round)), torch.cuda.current_stream(det.device).cuda_stream)
    return det


def rescale_table(ori_shapes, target_shapes, device):
    """``[B,5]`` fp32 device table for the fused rescale of ``lp_nms_f32``."""
    return torch.tensor([rescale_params(o, s) for o, s in zip(ori_shapes, target_shapes)],
                        dtype=torch.float64).to(torch.float32).to(device)


def txt_records(det, counts, src_shapes):
    """Batched ``--save-txt`` records of ``Inferer.infer`` (inferer.py:92-93,103-119) for
    ``det[B,max_det,28]`` (already rescaled + rounded), device ``counts[B]`` and per-image source
    shapes ``(H0, W0, ...)``: returns ``[B, max_det, 21]`` = 8 class ids | normalised xywh | 8
    normalised corners | conf (mean of groups 0..6, as the reference reports it)."""
    B, max_det = det.shape[0], det.shape[1]
    wh = torch.tensor([[float(s[1]), float(s[0])] for s in src_shapes], dtype=torch.float32).to(det.device)
    rec = torch.empty((B, max_det, 21), dtype=torch.float32, device=det.device)
    with torch.cuda.device(det.device):
        _abi.call("lp_txt_records_f32", det.data_ptr(), counts.data_ptr(), B, max_det, wh.data_ptr(), rec.data_ptr(),
                  torch.cuda.current_stream(det.device).cuda_stream)
    return rec


def txt_lines(records) -> str:
    """Text of ``n`` records (host tensor ``[n, 21]``) exactly as ``inferer.py:118-120`` writes it:
    ``('%g ' * 20).rstrip() % line`` + newline per detection (formatted natively)."""
    import ctypes
    r = records.detach().to("cpu", torch.float32).contiguous()
    n = r.shape[0]
    buf = ctypes.create_string_buffer(max(64, n * 20 * 16))
    written = ctypes.c_size_t(0)
    _abi.call("lp_txt_lines_host", r.data_ptr(), n, buf, len(buf), ctypes.byref(written))
    return buf.raw[:written.value].decode()
