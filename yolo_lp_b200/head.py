"""Detect-head inference decode and its geometry helpers on B200.

Drop-in surfaces (same names / argument meaning as the reference):
  ``generate_anchors``  yolov6/assigners/anchor_generator.py:4   (eval, anchor-free branch :11-31)
  ``dist2bbox``         yolov6/utils/general.py:29-40
  ``dist2cor``          yolov6/utils/general.py:51-66
  ``detect_decode``     the eval tail of ``Detect.forward``, effidehead.py:247-301
  ``detect_forward_eval`` / ``DetectEval``  the whole eval branch (:214-301): the module's own
                        convs (cuDNN, untouched) followed by the fused decode kernel.
"""
from __future__ import annotations

import ctypes

import torch

from . import _abi

CLS_NAMES = ("pro", "alp", "ad0", "ad1", "ad2", "ad3", "ad4", "ad5")
CLS_WIDTH = (31, 24, 37, 37, 37, 37, 37, 37)
ROW = _abi.ROW


def _cuda_f32(t: torch.Tensor, what: str, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or t.device.type != "cuda":
        raise RuntimeError(f"yolo_lp_b200: {what} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def generate_anchors(feats, fpn_strides, grid_cell_size=5.0, grid_cell_offset=0.5, device='cpu', is_eval=False,
                     mode='af'):
    """Generate anchors from features (reference signature).  Only the inference
    branch (``is_eval=True, mode='af'``) is on the hot path and implemented here;
    the training branch stays with the reference."""
    assert feats is not None
    if not is_eval or mode != 'af':
        raise NotImplementedError("yolo_lp_b200.generate_anchors covers is_eval=True, mode='af' only")
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("yolo_lp_b200.generate_anchors needs a CUDA device (no CPU fallback)")
    n = len(fpn_strides)
    hs = (ctypes.c_int * n)(*[int(feats[i].shape[2]) for i in range(n)])
    ws = (ctypes.c_int * n)(*[int(feats[i].shape[3]) for i in range(n)])
    ss = (ctypes.c_float * n)(*[float(s) for s in fpn_strides])
    A = sum(h * w for h, w in zip(hs, ws))
    points = torch.empty((A, 2), dtype=torch.float32, device=device)
    strides = torch.empty((A, 1), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _abi.call("lp_generate_anchors_f32", hs, ws, ss, n, float(grid_cell_offset), points.data_ptr(),
                  strides.data_ptr(), _stream(device))
    return points, strides


def _broadcast_anchor_args(distance, anchor_points, width):
    d = _cuda_f32(distance, "distance")
    ap = _cuda_f32(anchor_points, "anchor_points")
    if d.shape[-1] != width or ap.dim() != 2 or ap.shape[1] != 2 or d.dim() < 2 or d.shape[-2] != ap.shape[0]:
        raise ValueError(f"expected distance [..., A, {width}] and anchor_points [A, 2], got "
                         f"{tuple(distance.shape)} and {tuple(anchor_points.shape)}")
    A = ap.shape[0]
    return d, ap, d.numel() // (A * width), A


def dist2bbox(distance, anchor_points, box_format='xyxy'):
    '''Transform distance(ltrb) to box(xywh or xyxy).'''
    if box_format not in ('xyxy', 'xywh'):
        raise ValueError(box_format)
    d, ap, n, A = _broadcast_anchor_args(distance, anchor_points, 4)
    out = torch.empty_like(d)
    with torch.cuda.device(d.device):
        _abi.call("lp_dist2bbox_f32", d.data_ptr(), ap.data_ptr(), n, A, int(box_format == 'xywh'),
                  out.data_ptr(), _stream(d.device))
    return out


def dist2cor(distance, anchor_points):
    '''Transform 8 corner distances to the four corner points (TL, BL, BR, TR).'''
    d, ap, n, A = _broadcast_anchor_args(distance, anchor_points, 8)
    out = torch.empty_like(d)
    with torch.cuda.device(d.device):
        _abi.call("lp_dist2cor_f32", d.data_ptr(), ap.data_ptr(), n, A, out.data_ptr(), _stream(d.device))
    return out


class _LevelTable:
    """Validated ``lp_level_t[]`` for a fixed set of level tensors."""

    def __init__(self, levels, strides, allow_half: bool = False):
        n = len(levels)
        if not 0 < n <= _abi.MAX_LEVELS or len(strides) != n:
            raise ValueError("1..4 levels with one stride each")
        # fp16 level tensors (model.half()) stay halves when the caller has an f16 entry for them and
        # the shapes suit it (TMA kernel only: h*w % 8 == 0 at every level); otherwise they are upcast
        every = [t for lv in levels for t in lv.values()]
        self.half = bool(allow_half and all(t.dtype == torch.float16 for t in every)
                         and all((lv["reg"].shape[2] * lv["reg"].shape[3]) % 8 == 0 for lv in levels))
        dt = torch.float16 if self.half else torch.float32
        self.arr = (_abi.LpLevel * n)()
        self.keep = []  # keeps (possibly copied) contiguous inputs alive
        self.n = n
        self.B = B = int(levels[0]["reg"].shape[0])
        self.device = levels[0]["reg"].device
        A = 0
        for i, (lv, s) in enumerate(zip(levels, strides)):
            reg = _cuda_f32(lv["reg"], "reg", dt)
            _, c, h, w = reg.shape
            if c != 4:
                raise ValueError("reg must have 4 channels (use_dfl=False, reg_max=0)")
            cor = _cuda_f32(lv["cor"], "cor", dt)
            if tuple(cor.shape) != (B, 8, h, w):
                raise ValueError(f"cor shape {tuple(cor.shape)}")
            self.keep += [reg, cor]
            for g, (name, width) in enumerate(zip(CLS_NAMES, CLS_WIDTH)):
                t = _cuda_f32(lv[name], name, dt)
                if tuple(t.shape) != (B, width, h, w):
                    raise ValueError(f"{name} shape {tuple(t.shape)} != {(B, width, h, w)}")
                self.keep.append(t)
                self.arr[i].cls[g] = t.data_ptr()
            self.arr[i].reg, self.arr[i].cor = reg.data_ptr(), cor.data_ptr()
            self.arr[i].h, self.arr[i].w, self.arr[i].stride = h, w, float(s)
            A += h * w
        self.A = A


class DecodePlan(_LevelTable):
    """Validated, pointer-resolved launch of the decode kernel for a fixed set of level tensors
    (e.g. the static output buffers of a CUDA-graphed head).  ``run()`` is a single C call."""

    def __init__(self, levels, strides=(8, 16, 32), out: torch.Tensor | None = None, half_scores: bool = False):
        super().__init__(levels, strides)
        # half_scores: class scores rounded to half like the reference's model.half() head tensor
        self._entry = "lp_detect_decode_half_scores_f32" if half_scores else "lp_detect_decode_f32"
        B, A = self.B, self.A
        if out is None:
            out = torch.empty((B, A, ROW), dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != (B, A, ROW) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous fp32 [B, A, 290] tensor")
        self.out = out
        self.opts = None   # _abi.opts(...): per-call knobs (tests, tools); None = production

    def run(self) -> torch.Tensor:
        with torch.cuda.device(self.device):
            _abi.call(self._entry, self.arr, self.n, self.B, self.out.data_ptr(), _stream(self.device), opts=self.opts)
        return self.out


class PostprocessPlan(_LevelTable):
    """Fused head tail + NMS (``lp_detect_postprocess_f32``): raw level tensors -> detections without
    materialising ``[B, A, 290]``.  Bit-identical to ``DecodePlan`` followed by ``NmsPlan``."""

    KERNELS_PER_CALL = 2  # lp::levels_filter_tma_kernel (or lp::levels_filter_kernel), lp::nms_kernel<true>

    def __init__(self, levels, strides=(8, 16, 32), max_det: int = 300, max_nms: int = _abi.MAX_NMS,
                 want_anchor: bool = False):
        super().__init__(levels, strides, allow_half=True)
        self.max_det, self.max_nms = int(max_det), int(max_nms)
        self._sfx = "_f16" if self.half else "_f32"
        nbytes = _abi.detect_workspace_bytes(self.B, self.A, self.max_det)
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.out = torch.empty((self.B, self.max_det, _abi.OUT), dtype=torch.float32, device=self.device)
        self.counts = torch.empty((self.B,), dtype=torch.int32, device=self.device)
        self.kept_anchor = (torch.empty((self.B, self.max_det), dtype=torch.int32, device=self.device)
                            if want_anchor else None)
        # True while the last thing enqueued on this workspace was a pipelined step (whose K2 leaves the
        # candidate counters zeroed): only then may the next pipelined step skip its memset
        self.armed = False
        self.opts = None   # _abi.opts(...): per-call knobs (tests, tools); None = production

    def run(self, conf_thres, iou_thres, rescale=None, do_round=False):
        self.armed = False
        with torch.cuda.device(self.device):
            _abi.call("lp_detect_postprocess" + self._sfx, self.arr, self.n, self.B, float(conf_thres), float(iou_thres),
                      self.max_det, self.max_nms, self.workspace.data_ptr(), self.workspace.numel(),
                      self.out.data_ptr(), self.counts.data_ptr(),
                      self.kept_anchor.data_ptr() if self.kept_anchor is not None else None,
                      rescale.data_ptr() if rescale is not None else None, int(bool(do_round)), _stream(self.device),
                      opts=self.opts)
        return self.out, self.counts

    def run_filter(self, conf_thres):
        self.armed = False
        with torch.cuda.device(self.device):
            _abi.call("lp_detect_filter" + self._sfx, self.arr, self.n, self.B, float(conf_thres), self.max_det,
                      self.workspace.data_ptr(), self.workspace.numel(), _stream(self.device), opts=self.opts)

    def run_suppress(self, iou_thres, rescale=None, do_round=False):
        with torch.cuda.device(self.device):
            _abi.call("lp_detect_suppress_f32", self.arr, self.n, self.B, float(iou_thres), self.max_det, self.max_nms,
                      self.workspace.data_ptr(), self.workspace.numel(), self.out.data_ptr(), self.counts.data_ptr(),
                      self.kept_anchor.data_ptr() if self.kept_anchor is not None else None,
                      rescale.data_ptr() if rescale is not None else None, int(bool(do_round)), _stream(self.device),
                      opts=self.opts)
        return self.out, self.counts

    def candidate_counts(self) -> torch.Tensor:
        return self.workspace[: 4 * self.B].view(torch.int32)


class PostprocessPipeline:
    """Two-stream pipeline of the fused path (native ``lp_detect_pipelined_f32``): KF of batch i+1
    overlaps K2 of batch i.  ``plans`` are ``PostprocessPlan`` objects (one per in-flight batch, each
    bound to its own level tensors / workspace); ``submit(k)`` enqueues plan ``k % depth``."""

    def __init__(self, plans, filter_streams: int = 2):
        self.plans = list(plans)
        self.device = self.plans[0].device
        depth = len(self.plans)
        filter_streams = max(1, min(filter_streams, depth))
        with torch.cuda.device(self.device):
            lo, hi = torch.cuda.Stream.priority_range()
            # KF launches alternate between streams so that consecutive KFs overlap drain and ramp-up
            # (see NmsPipeline); they never share a workspace: depth >= filter_streams
            self.s_filters = [torch.cuda.Stream(self.device, priority=lo) for _ in range(filter_streams)]
            self.s_filter = self.s_filters[0]
            self.s_nms = torch.cuda.Stream(self.device, priority=hi)
            self.filtered = [torch.cuda.Event() for _ in range(depth)]
            self.done = [torch.cuda.Event() for _ in range(depth)]
            for ev in self.filtered + self.done:   # force creation of the cudaEvent_t handles
                ev.record(self.s_nms)
        self.n = 0

    def start(self):
        cur = torch.cuda.current_stream(self.device)
        for sf in self.s_filters:
            sf.wait_stream(cur)
        self.s_nms.wait_stream(cur)

    def submit(self, conf_thres, iou_thres, timing=None):
        slot = self.n % len(self.plans)
        plan = self.plans[slot]
        s_filter = self.s_filters[self.n % len(self.s_filters)]
        if timing is not None:
            for ev in timing:
                if ev.cuda_event == 0:
                    ev.record(s_filter)
        armed, plan.armed = plan.armed, False   # re-armed only once the whole step has been queued
        _abi.call("lp_detect_pipelined" + plan._sfx, plan.arr, plan.n, plan.B, float(conf_thres), float(iou_thres),
                  plan.max_det, plan.max_nms, plan.workspace.data_ptr(), plan.workspace.numel(), plan.out.data_ptr(),
                  plan.counts.data_ptr(), None, None, 0, s_filter.cuda_stream, self.s_nms.cuda_stream,
                  self.done[slot].cuda_event if armed else None,
                  self.filtered[slot].cuda_event, self.done[slot].cuda_event,
                  timing[0].cuda_event if timing is not None else None,
                  timing[1].cuda_event if timing is not None else None, opts=plan.opts)
        plan.armed = True
        self.n += 1
        return slot, plan.out, plan.counts

    def submit_to_host(self, conf_thres, iou_thres, out_host: torch.Tensor, counts_host: torch.Tensor,
                       copy_stream: torch.cuda.Stream, copied: torch.cuda.Event):
        """:meth:`submit` plus the D2H copy of the step's ``out`` / ``counts`` into the pinned host tensors on
        ``copy_stream``, all in ONE native call (``lp_detect_pipelined_to_host_f32``); ``copied`` fires when
        the host may read them.  Returns the slot."""
        slot = self.n % len(self.plans)
        plan = self.plans[slot]
        if not (out_host.is_pinned() and counts_host.is_pinned()) or out_host.dtype != torch.float32 or \
                counts_host.dtype != torch.int32 or tuple(out_host.shape) != tuple(plan.out.shape) or counts_host.numel() != plan.B:
            raise ValueError("out_host / counts_host must be pinned fp32 [B,max_det,28] / int32 [B] tensors")
        if copied.cuda_event == 0:       # torch creates the cudaEvent_t lazily, on the first record
            copied.record(copy_stream)
        s_filter = self.s_filters[self.n % len(self.s_filters)]
        armed, plan.armed = plan.armed, False   # re-armed only once the whole step has been queued
        _abi.call("lp_detect_pipelined_to_host" + plan._sfx, plan.arr, plan.n, plan.B, float(conf_thres), float(iou_thres),
                  plan.max_det, plan.max_nms, plan.workspace.data_ptr(), plan.workspace.numel(), plan.out.data_ptr(),
                  plan.counts.data_ptr(), None, 0, s_filter.cuda_stream, self.s_nms.cuda_stream,
                  self.done[slot].cuda_event if armed else None, self.filtered[slot].cuda_event, self.done[slot].cuda_event,
                  out_host.data_ptr(), counts_host.data_ptr(), copy_stream.cuda_stream, copied.cuda_event, opts=plan.opts)
        plan.armed = True
        self.n += 1
        return slot

    def finish(self):
        cur = torch.cuda.current_stream(self.device)
        for sf in self.s_filters:
            cur.wait_stream(sf)
        cur.wait_stream(self.s_nms)

    def capture(self, conf_thres, iou_thres, steps: int):
        """``steps`` pipelined steps as one CUDA graph (yolo_lp_b200.nms.GraphedSteps); this pipeline
        then belongs to the graph."""
        from .nms import GraphedSteps
        return GraphedSteps(self, lambda p: p.submit(conf_thres, iou_thres), steps)


def detect_postprocess(levels, strides=(8, 16, 32), conf_thres=0.25, iou_thres=0.45, max_det=300):
    """``non_max_suppression(Detect.forward(x))`` from the raw prediction-conv outputs in two
    launches (KF + K2); returns the reference's ``list[Tensor[k, 28]]``."""
    assert 0 <= conf_thres <= 1, f'conf_thresh must be in 0.0 to 1.0, however {conf_thres} is provided.'
    assert 0 <= iou_thres <= 1, f'iou_thres must be in 0.0 to 1.0, however {iou_thres} is provided.'
    plan = PostprocessPlan(levels, strides, max_det)
    out, counts = plan.run(conf_thres, iou_thres)
    from .nms import rows_of
    return rows_of(out, counts.cpu().tolist())


def detect_decode(levels, strides=(8, 16, 32), out: torch.Tensor | None = None, half_scores: bool = False) -> torch.Tensor:
    """Eval tail of ``Detect.forward`` (effidehead.py:247-301, ``use_dfl=False``).
    ``half_scores``: the ``model.half()`` variant -- class scores rounded to the nearest half (fp16 level
    tensors are upcast exactly; the result is fp32 like the reference's, see :func:`detect_forward_eval`).

    ``levels``: per FPN level a dict of the raw prediction-conv outputs (NCHW, CUDA fp32):
    ``pro[B,31,h,w] alp[B,24,h,w] ad0..ad5[B,37,h,w] reg[B,4,h,w] cor[B,8,h,w]``.
    Returns the head tensor ``[B, A, 290]``.
    """
    return DecodePlan(levels, strides, out, half_scores).run()


_PRED_ATTRS = tuple(n + "_preds" for n in CLS_NAMES)


def detect_forward_eval(detect, x):
    """Eval branch of the reference ``Detect.forward`` (effidehead.py:214-301) for a module
    with the reference's attribute names: the stem / cls / reg / prediction convs run as they
    are (torch + cuDNN), everything after them is one fused kernel.  (The reference also overwrites
    the caller's list ``x[i]`` with the stem outputs, effidehead.py:237; that side effect is not kept.)"""
    if getattr(detect, "use_dfl", False):
        raise NotImplementedError("use_dfl=True (distillation heads) is outside the LP configs")
    levels = []
    for i in range(detect.nl):
        f = detect.stems[i](x[i])
        cls_feat = detect.cls_convs[i](f)
        reg_feat = detect.reg_convs[i](f)
        lv = {name: getattr(detect, attr)[i](cls_feat) for name, attr in zip(CLS_NAMES, _PRED_ATTRS)}
        lv["reg"] = detect.reg_preds[i](reg_feat)
        lv["cor"] = detect.cor_preds[i](reg_feat)
        levels.append(lv)
    # model.half() (inferer.py:46-50): the reference's head tensor is fp32 even then -- its anchors are
    # fp32, so dist2bbox / dist2cor promote (general.py:29-66) and torch.cat promotes the half sigmoids
    # with them (effidehead.py:288-301).  The decode upcasts the half conv outputs exactly and computes
    # in fp32, which IS the reference's arithmetic for the box / corner columns, and rounds the class
    # scores to half as torch.sigmoid on a half tensor does (lp_detect_decode_half_scores_f32).
    return detect_decode(levels, [float(s) for s in detect.stride], half_scores=x[0].dtype == torch.float16)


def detect_forward_nms(detect, x, conf_thres=0.25, iou_thres=0.45, max_det=300):
    """``non_max_suppression(detect(x), ...)`` for a reference-style ``Detect`` module in eval mode:
    the module's convs, then the fused KF + K2 kernels (no ``[B, A, 290]`` tensor in between)."""
    if getattr(detect, "use_dfl", False):
        raise NotImplementedError("use_dfl=True (distillation heads) is outside the LP configs")
    levels = []
    for i in range(detect.nl):
        f = detect.stems[i](x[i])
        cls_feat = detect.cls_convs[i](f)
        reg_feat = detect.reg_convs[i](f)
        lv = {name: getattr(detect, attr)[i](cls_feat) for name, attr in zip(CLS_NAMES, _PRED_ATTRS)}
        lv["reg"] = detect.reg_preds[i](reg_feat)
        lv["cor"] = detect.cor_preds[i](reg_feat)
        levels.append(lv)
    return detect_postprocess(levels, [float(s) for s in detect.stride], conf_thres, iou_thres, max_det)


class DetectEval(torch.nn.Module):
    """Wraps a reference ``Detect`` module: training mode defers to it, eval mode runs
    :func:`detect_forward_eval`."""

    def __init__(self, detect):
        super().__init__()
        self.detect = detect

    def forward(self, x):
        if self.training:
            return self.detect(x)
        return detect_forward_eval(self.detect, x)
