"""Scalar-semantics fp32 restatement (numpy) of the YOLO-LP post-processing path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the
reference lines (relative to ``/root/reference``) it restates.  All arithmetic
is carried out in ``np.float32`` with the reference's association order and no
fused multiply-add, because kept-set parity is defined bit-for-bit.

Parity pin: ``tests/golden/*.npz`` (made by ``tests/golden/make_golden.py`` from
the imported reference + torchvision 0.26.0 CPU).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32

#: column groups of the 290-wide head row: province, alphabet, six characters
#: (``yolov6/utils/nms.py:81-88``).
GROUPS = ((13, 44), (44, 68), (68, 105), (105, 142), (142, 179), (179, 216), (216, 253), (253, 290))
ROW = 290
OUT = 28
MAX_NMS = 30000  # yolov6/utils/nms.py:62


# --------------------------------------------------------------------------- geometry
def generate_anchors_eval(level_hw, strides, grid_cell_offset=0.5):
    """``generate_anchors(..., is_eval=True, mode='af')``,
    ``yolov6/assigners/anchor_generator.py:11-31``.

    level_hw: [(h, w), ...].  Returns (anchor_points[A,2], stride_tensor[A,1]),
    level-major, row-major inside a level, points = (x+off, y+off) in grid units.
    """
    pts, strs = [], []
    for (h, w), s in zip(level_hw, strides):
        sx = np.arange(w, dtype=np.int64).astype(f32) + f32(grid_cell_offset)
        sy = np.arange(h, dtype=np.int64).astype(f32) + f32(grid_cell_offset)
        yy, xx = np.meshgrid(sy, sx, indexing="ij")
        pts.append(np.stack([xx, yy], -1).reshape(-1, 2).astype(f32))
        strs.append(np.full((h * w, 1), s, dtype=f32))
    return np.concatenate(pts, 0), np.concatenate(strs, 0)


def dist2bbox(distance, anchor_points, box_format="xyxy"):
    """``yolov6/utils/general.py:29-40``."""
    distance = np.asarray(distance, f32)
    lt, rb = distance[..., :2], distance[..., 2:4]
    x1y1 = (anchor_points - lt).astype(f32)
    x2y2 = (anchor_points + rb).astype(f32)
    if box_format == "xyxy":
        return np.concatenate([x1y1, x2y2], -1)
    if box_format == "xywh":
        c_xy = ((x1y1 + x2y2).astype(f32) / f32(2)).astype(f32)
        wh = (x2y2 - x1y1).astype(f32)
        return np.concatenate([c_xy, wh], -1)
    raise ValueError(box_format)


def dist2cor(distance, anchor_points):
    """``yolov6/utils/general.py:51-66``: 8 distances -> TL, BL, BR, TR corners."""
    d = np.asarray(distance, f32)
    ax, ay = anchor_points[..., 0], anchor_points[..., 1]
    out = np.stack(
        [ax - d[..., 0], ay - d[..., 1],   # TL = a - lt
         ax - d[..., 2], ay + d[..., 3],   # BL
         ax + d[..., 4], ay + d[..., 5],   # BR = a + rb
         ax + d[..., 6], ay - d[..., 7]],  # TR
        -1)
    return out.astype(f32)


def sigmoid(x):
    """torch.sigmoid in fp32.  Not bit-reproducible across libms; parity for
    this one op is held to 1e-5 relative (north_star)."""
    x = np.asarray(x, f32)
    return (f32(1) / (f32(1) + np.exp(-x, dtype=f32))).astype(f32)


CLS_NAMES = ("pro", "alp", "ad0", "ad1", "ad2", "ad3", "ad4", "ad5")


def detect_decode(levels, strides, half_scores=False):
    """Eval tail of ``Detect.forward``, ``yolov6/models/effidehead.py:247-301``
    (``use_dfl=False``): everything after the prediction convs.  ``half_scores``: the ``model.half()``
    forward -- level tensors are halves (upcast exactly here), ``torch.sigmoid`` rounds every score to
    half (:251-258) and ``torch.cat`` promotes it to fp32 with the geometry columns, which are computed
    in fp32 because the anchor points are fp32 (:283-301).

    levels: list (per FPN level) of dicts with raw conv outputs, NCHW fp32:
    ``pro[B,31,h,w] alp[B,24,h,w] ad0..ad5[B,37,h,w] reg[B,4,h,w] cor[B,8,h,w]``.
    Returns ``[B, A, 290]``.
    """
    hw = [lv["reg"].shape[2:] for lv in levels]
    ap, st = generate_anchors_eval(hw, strides)
    B = levels[0]["reg"].shape[0]

    def flat(name):  # effidehead.py:260-280  reshape [B,C,hw] -> cat levels -> permute
        return np.concatenate([np.asarray(lv[name], f32).reshape(B, lv[name].shape[1], -1) for lv in levels], -1).transpose(0, 2, 1)

    box = dist2bbox(flat("reg"), ap[None], "xywh")          # :283
    cor = dist2cor(flat("cor"), ap[None])                     # :284
    box = (box * st[None]).astype(f32)                        # :285
    cor = (cor * st[None]).astype(f32)                        # :286
    cls = [sigmoid(flat(n)) for n in CLS_NAMES]               # :251-258
    if half_scores:
        cls = [c.astype(np.float16).astype(f32) for c in cls]
    ones = np.ones((B, box.shape[1], 1), f32)                 # :290
    return np.concatenate([box, ones, cor] + cls, -1).astype(f32)   # :287-301


# --------------------------------------------------------------------------- NMS path
def xywh2xyxy(x):
    """``yolov6/utils/nms.py:21-28``."""
    x = np.asarray(x, f32)
    h0 = (x[:, 2] / f32(2)).astype(f32)
    h1 = (x[:, 3] / f32(2)).astype(f32)
    return np.stack([x[:, 0] - h0, x[:, 1] - h1, x[:, 0] + h0, x[:, 1] + h1], 1).astype(f32)


def _sum8(c, last):
    """Left-to-right fp32 sum of seven group confidences plus column ``last``."""
    s = (c[:, 0] + c[:, 1]).astype(f32)
    for k in (2, 3, 4, 5, 6):
        s = (s + c[:, k]).astype(f32)
    return (s + c[:, last]).astype(f32)


def score_rows(x):
    """``nms.py:76-97`` for one image ``x[A,290]``: returns (det[A,28], filter_mean[A]).

    filter_mean reproduces the reference's bug of adding ``ad4_conf`` twice and
    never ``ad5_conf`` (``nms.py:90-91``).
    """
    x = np.asarray(x, f32)
    cls = (x[:, 13:] * x[:, 4:5]).astype(f32)                                   # :76
    box = xywh2xyxy(x[:, :4])                                                   # :79
    conf = np.stack([cls[:, s - 13:e - 13].max(1) for s, e in GROUPS], 1)       # :81-88
    arg = np.stack([cls[:, s - 13:e - 13].argmax(1) for s, e in GROUPS], 1)     # first index on ties
    filt = (_sum8(conf, 6) / f32(8)).astype(f32)                                # :90-91
    det = np.concatenate([box, x[:, 5:13], conf, arg.astype(f32)], 1).astype(f32)  # :94-96
    return det, filt


def nms_score(det):
    """``nms.py:120``: mean of all eight group confidences, left to right."""
    return (_sum8(det[:, 12:20], 7) / f32(8)).astype(f32)


def greedy_nms(boxes, scores, iou_thres, limit=None):
    """``torchvision.ops.nms`` CPU kernel (0.26.0), call site ``nms.py:121``.

    Stable descending sort; areas precomputed in fp32; fp32 IoU with the union
    associated as ``(area_i + area_j) - inter``; suppression iff
    ``(double)iou > iou_thres``.  ``limit`` stops after that many keeps (legal:
    the reference truncates afterwards, ``nms.py:122-123``).
    """
    boxes = np.asarray(boxes, f32)
    scores = np.asarray(scores, f32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    order = np.argsort(-scores.astype(np.float64), kind="stable")
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    area = ((x2 - x1).astype(f32) * (y2 - y1).astype(f32)).astype(f32)
    supp = np.zeros(n, bool)
    keep = []
    thr = float(iou_thres)
    with np.errstate(invalid="ignore", divide="ignore"):
        for k, i in enumerate(order):
            if supp[i]:
                continue
            keep.append(i)
            if limit is not None and len(keep) >= limit:
                break
            r = order[k + 1:]
            w = np.maximum(f32(0), (np.minimum(x2[i], x2[r]) - np.maximum(x1[i], x1[r])).astype(f32))
            h = np.maximum(f32(0), (np.minimum(y2[i], y2[r]) - np.maximum(y1[i], y1[r])).astype(f32))
            inter = (w * h).astype(f32)
            ovr = (inter / ((area[i] + area[r]).astype(f32) - inter).astype(f32)).astype(f32)
            supp[r[ovr.astype(np.float64) > thr]] = True
    return np.asarray(keep, np.int64)


def nms_one_image(x, conf_thres, iou_thres, max_det=300, max_nms=MAX_NMS):
    """Loop body of ``non_max_suppression``, ``nms.py:68-125``, for one image.

    Returns (rows[k,28], anchor_index[k]).  When more than ``max_nms`` rows pass
    the filter the reference takes an *unstable* argsort (``nms.py:115-116``), so
    its result is implementation-defined on ties; the oracle (and the product)
    define the cut as (score descending, anchor ascending).
    """
    det, filt = score_rows(x)
    mask = filt >= f32(conf_thres)
    idx = np.nonzero(mask)[0]
    det = det[mask]                                                             # :97
    if det.shape[0] == 0:
        return np.zeros((0, OUT), f32), np.zeros((0,), np.int64)
    score = nms_score(det)
    if det.shape[0] > max_nms:                                                  # :115-116
        o = np.argsort(-score.astype(np.float64), kind="stable")[:max_nms]
        det, idx, score = det[o], idx[o], score[o]
    keep = greedy_nms(det[:, :4], score, iou_thres, limit=max_det)[:max_det]    # :121-123
    return det[keep], idx[keep]


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None,
                        agnostic=False, multi_label=False, max_det=300, return_index=False):
    """``yolov6/utils/nms.py:31-130``.  ``classes``/``agnostic``/``multi_label``
    are accepted and ignored exactly as in the reference; the 10 s wall-clock
    ``time_limit`` (``:126-128``) is deliberately not reproduced; the input is
    not mutated (the reference's ``x[:,13:] *= x[:,4:5]`` side effect, ``:76``)."""
    assert 0 <= conf_thres <= 1, f"conf_thresh must be in 0.0 to 1.0, however {conf_thres} is provided."
    assert 0 <= iou_thres <= 1, f"iou_thres must be in 0.0 to 1.0, however {iou_thres} is provided."
    prediction = np.asarray(prediction, f32)
    rows, idxs = [], []
    for x in prediction:
        r, i = nms_one_image(x, conf_thres, iou_thres, max_det)
        rows.append(r)
        idxs.append(i)
    return (rows, idxs) if return_index else rows


# --------------------------------------------------------------------------- rescale
def rescale_params(ori_shape, target_shape):
    """Python-double ratio and padding of ``Inferer.rescale``,
    ``yolov6/core/inferer.py:206-207``."""
    ratio = min(ori_shape[0] / target_shape[0], ori_shape[1] / target_shape[1])
    pad_x = (ori_shape[1] - target_shape[1] * ratio) / 2
    pad_y = (ori_shape[0] - target_shape[0] * ratio) / 2
    return ratio, pad_x, pad_y


def rescale(ori_shape, boxes_and_cors, target_shape, do_round=False):
    """``Inferer.rescale``, ``yolov6/core/inferer.py:203-228`` on ``[k,12]`` fp32.
    ``do_round`` applies the caller's ``.round()`` (``inferer.py:100``, half-to-even)."""
    ratio, pad_x, pad_y = rescale_params(ori_shape, target_shape)
    v = np.array(boxes_and_cors, f32, copy=True)
    v[:, 0::2] = (v[:, 0::2] - f32(pad_x)).astype(f32)      # :210
    v[:, 1::2] = (v[:, 1::2] - f32(pad_y)).astype(f32)      # :211
    v = (v / f32(ratio)).astype(f32)                        # :212  true division by the fp32-rounded scalar
    v[:, 0::2] = np.clip(v[:, 0::2], f32(0), f32(target_shape[1]))   # :214-225
    v[:, 1::2] = np.clip(v[:, 1::2], f32(0), f32(target_shape[0]))
    if do_round:
        v = np.rint(v).astype(f32)
    return v


# --------------------------------------------------------------------------- caller-side records
def txt_records(det, src_shape):
    """Per-detection ``--save-txt`` record of ``Inferer.infer``, ``yolov6/core/inferer.py:92-93,103-119``
    (after the rescale + round of ``:100``): ``[8 class ids | xywh / (W0,H0,W0,H0) | 8 corners /
    (W0,H0)x4 | conf]`` with ``xywh = box_convert(xyxy)`` (``inferer.py:309-316``) and
    ``conf = mean(row[12:19])`` (seven of the eight groups, ``:113``).  ``det``: [k,28] fp32;
    ``src_shape``: (H0, W0, ...).  Returns [k,21] fp32."""
    det = np.asarray(det, f32)
    h0, w0 = f32(src_shape[0]), f32(src_shape[1])
    gn = np.array([w0, h0, w0, h0], f32)
    x1, y1, x2, y2 = det[:, 0], det[:, 1], det[:, 2], det[:, 3]
    xywh = np.stack([((x1 + x2).astype(f32) / f32(2)).astype(f32), ((y1 + y2).astype(f32) / f32(2)).astype(f32),
                     (x2 - x1).astype(f32), (y2 - y1).astype(f32)], 1)
    xywh = (xywh / gn).astype(f32)
    cor = (det[:, 4:12] / np.tile(gn[:2], 4)).astype(f32)
    c = det[:, 12:19]
    s = c[:, 0]
    for k in range(1, 7):
        s = (s + c[:, k]).astype(f32)
    conf = (s / f32(7)).astype(f32)
    return np.concatenate([det[:, 20:28], xywh, cor, conf[:, None]], 1).astype(f32)


def txt_line(record):
    """The text line of ``inferer.py:118-120``: ``('%g ' * 20).rstrip() % line`` over the first 20 numbers."""
    vals = tuple(float(v) for v in record[:20])
    return ('%g ' * len(vals)).rstrip() % vals


# --------------------------------------------------------------------------- LP eval metric
def box_iou(box1, box2):
    """``yolov6/utils/general.py:93-115``: [N,4] x [M,4] -> [N,M], fp32,
    ``inter / ((area1 + area2) - inter)``."""
    b1, b2 = np.asarray(box1, f32), np.asarray(box2, f32)
    area1 = ((b1[:, 2] - b1[:, 0]).astype(f32) * (b1[:, 3] - b1[:, 1]).astype(f32)).astype(f32)
    area2 = ((b2[:, 2] - b2[:, 0]).astype(f32) * (b2[:, 3] - b2[:, 1]).astype(f32)).astype(f32)
    wh = np.clip((np.minimum(b1[:, None, 2:4], b2[None, :, 2:4]) - np.maximum(b1[:, None, :2], b2[None, :, :2])).astype(f32),
                 f32(0), None)
    inter = (wh[..., 0] * wh[..., 1]).astype(f32)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (inter / ((area1[:, None] + area2[None, :]).astype(f32) - inter).astype(f32)).astype(f32)


def eval_match(pred, target):
    """Per-target matching of ``Evaler.eval`` for one image, ``yolov6/core/evaler.py:183-229``.
    pred [n,28] (NMS rows), target [m,20] = 8 class ids | xyxy | 8 corners.  Returns per target
    (t_iou fp32, match index, is_cor, is_cls); t_iou = -1 where the image has no predictions."""
    pred, target = np.asarray(pred, f32), np.asarray(target, f32)
    m = target.shape[0]
    t_iou = np.full(m, -1, f32)
    match = np.zeros(m, np.int64)
    is_cor = np.zeros(m, bool)
    is_cls = np.zeros(m, bool)
    if pred.shape[0] == 0 or m == 0:
        return t_iou, match, is_cor, is_cls
    iou = box_iou(pred[:, :4], target[:, 8:12])                       # :189
    match = iou.argmax(0)                                             # :190 first index on ties
    t_iou = iou[match, np.arange(m)]
    for k in range(m):
        tp, tt = pred[match[k]], target[k]
        area = ((tt[10] - tt[8]).astype(f32) * (tt[11] - tt[9]).astype(f32)).astype(f32)   # :213
        d = np.abs((tp[4:12] - tt[12:20]).astype(f32))
        # torch.sum over 8 contiguous fp32 values (:218): the CPU kernel adds in the fixed order below
        ssum = torch_like_sum8(d)
        lhs = (ssum / f32(8)).astype(f32)
        rhs = (f32(0.1) * np.sqrt(area, dtype=f32)).astype(f32)
        is_cor[k] = bool(lhs < rhs)
        is_cls[k] = all(int(tp[20 + i]) == int(tt[i]) for i in range(8))                  # :223-226
    return t_iou, match, is_cor, is_cls


def torch_like_sum8(d):
    """Sum of 8 fp32 values as ``torch.sum`` computes it on CPU for a short contiguous vector:
    plain left-to-right accumulation (verified against the goldens)."""
    s = f32(d[0])
    for v in d[1:]:
        s = f32(s + f32(v))
    return s


IOU_LIST = [0.5 + i * 0.05 for i in range(10)]   # evaler.py:160, Python doubles


def eval_accumulate(per_image):
    """Counters of ``Evaler.eval`` (``evaler.py:160-245``) from the per-target matches, in image /
    target order.  Reproduces the reference's quirks: comparisons of the fp32 IoU against Python
    doubles happen in fp32; an IoU of exactly 1.0 falls into no bin, so the correctness counters
    reuse the PREVIOUS target's bin index (``iou_idx`` is a stale loop variable) and ``pred_cnts``
    skips it.  ``per_image``: list of (n_pred, t_iou, is_cor, is_cls)."""
    lo = [f32(v) for v in IOU_LIST]
    hi = [f32(v + 0.05) for v in IOU_LIST]
    c = dict(true_cnt=0, pred_cnt=0, pred_cnts=[0] * 10, cor_right=[0] * 10, cls_right=[0] * 10, right=[0] * 10)
    iou_idx = None
    for n_pred, t_iou, is_cor, is_cls in per_image:
        c["true_cnt"] += len(t_iou)
        if n_pred == 0 or len(t_iou) == 0:
            continue
        for k in range(len(t_iou)):
            v = f32(t_iou[k])
            if v < f32(0.5):
                continue
            if v >= f32(0.7):
                c["pred_cnt"] += 1
            for n in range(10):
                if v >= lo[n] and v < hi[n]:
                    iou_idx = n
                    break
            if iou_idx is None:
                raise NameError("iou_idx")   # the reference raises here too
            if is_cor[k]:
                c["cor_right"][iou_idx] += 1
            if is_cls[k]:
                c["cls_right"][iou_idx] += 1
            if is_cor[k] and is_cls[k]:
                c["right"][iou_idx] += 1
        for k in range(len(t_iou)):
            v = f32(t_iou[k])
            if v < f32(0.5):
                continue
            for n in range(10):
                if v >= lo[n] and v < hi[n]:
                    c["pred_cnts"][n] += 1
                    break
    return c


def eval_summary(c):
    """mAP / recall figures of ``evaler.py:247-283`` from the counters (Python doubles)."""
    right, pred_cnts = c["right"], c["pred_cnts"]
    mAP_list = [right[i] / pred_cnts[i] if pred_cnts[i] > 0 else -int(right[i] == pred_cnts[i]) for i in range(10)]
    valid = [v for v in mAP_list if v != -1]
    mAP_50_95 = sum(valid) / len(valid) if valid else 0.0
    right_50, pred_50 = sum(right), sum(pred_cnts)
    right_75 = sum(right[i] for i in range(10) if IOU_LIST[i] >= 0.75)
    pred_75 = sum(pred_cnts[i] for i in range(10) if IOU_LIST[i] >= 0.75)
    t_right = sum(right[i] for i in range(10) if IOU_LIST[i] >= 0.7)
    mAP_50 = right_50 / pred_50 if pred_50 > 0 else 0.0
    mAP_75 = right_75 / pred_75 if pred_75 > 0 else 0.0
    mAP = t_right / c["pred_cnt"] if c["pred_cnt"] > 0 else 0.0
    recall_list = [sum(right[: i + 1]) / c["true_cnt"] if c["true_cnt"] > 0 else 0.0 for i in range(10)]
    recall = sum(right) / c["true_cnt"]
    return [mAP, mAP_50, mAP_75, mAP_50_95, recall, mAP_list, recall_list]


def prepare_targets(targets, w, h, batch_size):
    """Label prep of ``Evaler.predict``, ``yolov6/core/evaler.py:119-127``: ``targets[T,21]`` =
    image index | 8 class ids | normalised xywh | 8 normalised corners  ->  per image ``[m,20]`` =
    8 class ids | xyxy (pixels of the letterboxed batch) | 8 corners (pixels)."""
    t = np.array(targets, f32, copy=True)
    if t.shape[0]:
        t[:, 9:13] = xywh2xyxy(t[:, 9:13])                      # :120
        for j in range(9, 21, 2):                               # :123-125
            t[:, j] = (t[:, j] * f32(w)).astype(f32)
            t[:, j + 1] = (t[:, j + 1] * f32(h)).astype(f32)
    out = [np.zeros((0, 20), f32) for _ in range(batch_size)]
    for row in t:                                               # :126, order preserved per image
        b = int(row[0])
        out[b] = np.concatenate([out[b], row[None, 1:]], 0)
    return out
