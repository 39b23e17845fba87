"""Torch-CPU port of the reference post-processing path, used as the TIMED CPU
baseline (``bench.py`` ``cpu_baseline`` / ``--impl reference``) and as a second
checker.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``/root/reference``
does not exist on the GPU box, so the reference cannot be imported there; this
port issues the same ATen op sequence per image (including the dead
``pred_candidates`` computation, ``nms.py:48-55``, which costs the reference
~12 % of its time) and calls the same third-party kernel,
``torchvision.ops.nms`` (``nms.py:121``), so its timing stands in for the
reference's.  ``tests/test_oracle_golden.py`` checks it against the goldens.
"""
from __future__ import annotations

import torch
import torchvision

from .lp_oracle import GROUPS, MAX_NMS


def _sum_left(cols):
    s = cols[0] + cols[1]
    for c in cols[2:]:
        s = s + c
    return s


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None,
                        agnostic=False, multi_label=False, max_det=300):
    """Port of ``yolov6/utils/nms.py:31-130`` (same ops, same order, same
    in-place mutation of ``prediction``; no wall-clock time limit)."""
    # nms.py:47-55 -- dead code in the reference, kept for timing fidelity only.
    judge = torch.max(prediction[..., 13:47], dim=-1)[0] > conf_thres
    _ = torch.logical_and(prediction[..., 4] > conf_thres, judge)
    assert 0 <= conf_thres <= 1, f"conf_thresh must be in 0.0 to 1.0, however {conf_thres} is provided."
    assert 0 <= iou_thres <= 1, f"iou_thres must be in 0.0 to 1.0, however {iou_thres} is provided."

    empty = torch.zeros((0, 28), device=prediction.device)
    out = [empty] * prediction.shape[0]
    for b, x in enumerate(prediction):
        if not x.shape[0]:
            continue
        x[:, 13:] *= x[:, 4:5]                                           # :76
        half_w, half_h = x[:, 2] / 2, x[:, 3] / 2                         # :21-28
        box = torch.stack((x[:, 0] - half_w, x[:, 1] - half_h, x[:, 0] + half_w, x[:, 1] + half_h), 1)
        conf, arg = zip(*(torch.max(x[:, s:e], 1, keepdim=True) for s, e in GROUPS))   # :81-88
        sq = [c.squeeze() for c in conf]
        mask = (_sum_left(sq[:7] + [sq[6]]) / 8.0 >= conf_thres).squeeze()   # :90-91 (ad4 twice)
        det = torch.cat((box, x[:, 5:13]) + conf + arg, 1)[mask]          # :94-97
        n = det.shape[0]
        if not n:
            continue
        if n > MAX_NMS:                                                   # :115-116
            sc = _sum_left([det[:, 12 + k] for k in range(8)]) / 8.0
            det = det[sc.argsort(descending=True)[:MAX_NMS]]
        scores = _sum_left([det[:, 12 + k] for k in range(8)]) / 8.0      # :120
        keep = torchvision.ops.nms(det[:, :4], scores, iou_thres)[:max_det]   # :121-123
        out[b] = det[keep]
    return out


def rescale(ori_shape, boxes_and_cors, target_shape):
    """Port of ``Inferer.rescale`` (``yolov6/core/inferer.py:203-228``), in place."""
    ratio = min(ori_shape[0] / target_shape[0], ori_shape[1] / target_shape[1])
    pad = (ori_shape[1] - target_shape[1] * ratio) / 2, (ori_shape[0] - target_shape[0] * ratio) / 2
    boxes_and_cors[:, 0::2] -= pad[0]
    boxes_and_cors[:, 1::2] -= pad[1]
    boxes_and_cors[:, :] /= ratio
    boxes_and_cors[:, 0::2].clamp_(0, target_shape[1])
    boxes_and_cors[:, 1::2].clamp_(0, target_shape[0])
    return boxes_and_cors


def detect_decode(levels, strides):
    """Port of the eval tail of ``Detect.forward`` (``effidehead.py:247-301``)
    on raw per-level conv outputs (dicts of NCHW tensors, see lp_oracle)."""
    names = ("pro", "alp", "ad0", "ad1", "ad2", "ad3", "ad4", "ad5")
    B = levels[0]["reg"].shape[0]
    pts, strs = [], []
    for lv, s in zip(levels, strides):
        h, w = lv["reg"].shape[2:]
        sy, sx = torch.meshgrid(torch.arange(h) + 0.5, torch.arange(w) + 0.5, indexing="ij")
        pts.append(torch.stack([sx, sy], -1).float().reshape(-1, 2))
        strs.append(torch.full((h * w, 1), float(s)))
    ap, st = torch.cat(pts), torch.cat(strs)

    def flat(n, act=None):
        parts = [(act(lv[n]) if act else lv[n]).reshape(B, lv[n].shape[1], -1) for lv in levels]
        return torch.cat(parts, -1).permute(0, 2, 1)

    reg, cor = flat("reg"), flat("cor")
    x1y1, x2y2 = ap - reg[..., :2], ap + reg[..., 2:]
    box = torch.cat([(x1y1 + x2y2) / 2, x2y2 - x1y1], -1)
    ax, ay = ap[:, 0:1], ap[:, 1:2]
    corners = torch.cat([ap - cor[..., 0:2], ax - cor[..., 2:3], ay + cor[..., 3:4],
                         ap + cor[..., 4:6], ax + cor[..., 6:7], ay - cor[..., 7:8]], -1)
    box = box * st
    corners = corners * st
    ones = torch.ones((B, box.shape[1], 1))
    return torch.cat([box, ones, corners] + [flat(n, torch.sigmoid) for n in names], -1)
