"""LP evaluation metric of ``Evaler.eval`` (``yolov6/core/evaler.py:153-283``) on B200.

The reference walks every target of every image in a triple Python loop; here the per-target
matching runs in one kernel launch (one warp per target) and the order-dependent counter logic --
including the reference's stale-bin quirk for an IoU of exactly 1.0 -- in a native host function.
"""
from __future__ import annotations

import torch

from . import _abi

OUT = _abi.OUT


def _flatten(nested):
    """``preds`` / ``targets`` as ``Evaler.predict`` returns them: a list of batches, each a list of
    per-image tensors (a flat list of per-image tensors is accepted too)."""
    flat = []
    for item in nested:
        if isinstance(item, torch.Tensor):
            flat.append(item)
        else:
            flat.extend(item)
    return flat


def eval_counts(preds, targets, device=None):
    """Returns (counters[42] int64, summary[25] float64) -- layouts in ``include/lpnms.h``."""
    preds, targets = _flatten(preds), _flatten(targets)
    assert len(preds) == len(targets), 'predict imgs count is not match with targets!'   # evaler.py:157
    B = len(preds)
    if device is None:
        device = next((p.device for p in preds if p.is_cuda), torch.device("cuda", torch.cuda.current_device()))
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("yolo_lp_b200.evaler needs a CUDA device (no CPU fallback)")
    counts_host = torch.tensor([p.shape[0] for p in preds], dtype=torch.int32)
    max_det = max(1, int(counts_host.max())) if B else 1
    det = torch.zeros((B, max_det, OUT), dtype=torch.float32, device=device)
    for b, p in enumerate(preds):
        if p.shape[0]:
            det[b, :p.shape[0]] = p.to(device=device, dtype=torch.float32)
    tcount = [t.shape[0] for t in targets]
    T = sum(tcount)
    timg_host = torch.repeat_interleave(torch.arange(B, dtype=torch.int32), torch.tensor(tcount, dtype=torch.int64))
    match_host = torch.zeros((T, 4), dtype=torch.float32)
    if T:
        tgt = torch.cat([t.to(device=device, dtype=torch.float32).reshape(-1, 20) for t in targets if t.shape[0]])
        match = torch.empty((T, 4), dtype=torch.float32, device=device)
        counts_dev, timg_dev = counts_host.to(device), timg_host.to(device)   # named: must outlive the launch
        with torch.cuda.device(device):
            _abi.call("lp_eval_match_f32", det.data_ptr(), counts_dev.data_ptr(), B, max_det, tgt.data_ptr(),
                      timg_dev.data_ptr(), T, match.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
        match_host = match.cpu()
    counters = torch.zeros(42, dtype=torch.int64)
    summary = torch.zeros(25, dtype=torch.float64)
    _abi.call("lp_eval_accumulate_host", match_host.data_ptr(), timg_host.data_ptr(), counts_host.data_ptr(), B, T,
              counters.data_ptr(), summary.data_ptr())
    return counters, summary


def lp_eval(preds, targets, device=None):
    """Drop-in for the value ``Evaler.eval(preds, targets, model, task)`` returns:
    ``[mAP, mAP_50, mAP_75, mAP_50_95, recall, mAP_list, recall_list]``."""
    _, s = eval_counts(preds, targets, device)
    s = s.tolist()
    return [s[0], s[1], s[2], s[3], s[4], s[5:15], s[15:25]]


def prepare_targets(targets, w, h, batch_size):
    """Label prep of ``Evaler.predict`` (``evaler.py:119-127``): ``targets[T,21]`` (CUDA; image index |
    8 class ids | normalised xywh | 8 normalised corners) -> list of ``batch_size`` tensors ``[m,20]``
    (8 class ids | xyxy in pixels | 8 corners in pixels), rows of an image in their input order."""
    if not isinstance(targets, torch.Tensor) or targets.device.type != "cuda":
        raise RuntimeError("yolo_lp_b200.evaler.prepare_targets needs a CUDA tensor (no CPU fallback)")
    t = targets.to(torch.float32).contiguous()
    T = t.shape[0]
    out = torch.empty((T, 20), dtype=torch.float32, device=t.device)
    img = torch.empty((T,), dtype=torch.int32, device=t.device)
    with torch.cuda.device(t.device):
        _abi.call("lp_prepare_targets_f32", t.data_ptr(), T, float(w), float(h), out.data_ptr(), img.data_ptr(),
                  torch.cuda.current_stream(t.device).cuda_stream)
    img = img.long()
    return [out[img == b] for b in range(batch_size)]   # boolean mask keeps the input order
