"""Drop-in ``non_max_suppression`` (``yolov6/utils/nms.py:31-130``) on B200.

Same signature, defaults and return type as the reference:
``list`` of ``B`` fp32 tensors ``[k_b, 28]`` on ``prediction.device``, rows =
``xyxy | 8 corner coords | 8 group confidences | 8 group argmax (as float)``,
ordered by decreasing mean score (ties: ascending anchor).  All arithmetic runs
in ``liblpnms.so`` (K1 filter + K2 sort/NMS/gather); PyTorch only owns the
buffers and the stream.

Deliberate differences from the reference:
  * ``prediction`` is NOT mutated (reference: ``x[:,13:] *= x[:,4:5]``, nms.py:76);
  * no 10 s wall-clock ``time_limit`` early exit (nms.py:63,126-128);
  * with more than 30 000 candidates the cut is (score desc, anchor asc) where the
    reference's unstable argsort is implementation-defined (nms.py:115-116);
  * fp16 predictions (the reference's ``--half`` mode, inferer.py:46-50): the rows come
    back as fp16 like the reference's (its ``torch.cat`` keeps the prediction dtype,
    nms.py:94-96), but they are the fp32 path's rows on ``prediction.float()`` rounded
    once to half -- every load is an exact upcast and all arithmetic is fp32 -- not the
    reference's per-operation half arithmetic, which exists on CUDA only and sits behind
    an unstable sort of massively tied half scores (no CPU oracle can pin it);
  * a CPU ``prediction`` is not a fallback: it is streamed through the GPU kernels and
    the rows come back as CPU tensors.
"""
from __future__ import annotations

import torch

from . import _abi

ROW, OUT, MAX_NMS = _abi.ROW, _abi.OUT, _abi.MAX_NMS


def _entry(stem: str, pred: torch.Tensor) -> str:
    """C entry for the tensor's storage type: fp16 head tensors (the reference's ``--half`` mode,
    inferer.py:46-50) are read natively -- every value upcast exactly on load -- so the result is
    bit for bit that of the fp32 entry on ``pred.float()`` at half the bytes."""
    return stem + ("_f16" if pred.dtype == torch.float16 else "_f32")


def _check_thresholds(conf_thres, iou_thres):
    # nms.py:57-58, same messages
    assert 0 <= conf_thres <= 1, f'conf_thresh must be in 0.0 to 1.0, however {conf_thres} is provided.'
    assert 0 <= iou_thres <= 1, f'iou_thres must be in 0.0 to 1.0, however {iou_thres} is provided.'


class NmsPlan:
    """Pre-allocated buffers for repeated NMS calls of one shape on one device.

    ``run`` only enqueues work on the current stream (memset + 2 kernel launches)
    and returns device tensors; nothing synchronises.
    """

    KERNELS_PER_CALL = 2  # lp::filter_kernel, lp::nms_kernel

    def __init__(self, B: int, A: int, max_det: int = 300, device=None, max_nms: int = MAX_NMS,
                 want_anchor: bool = False):
        _abi.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("yolo_lp_b200 runs on CUDA devices only (no CPU fallback)")
        self.B, self.A, self.max_det, self.max_nms = int(B), int(A), int(max_det), int(max_nms)
        nbytes = _abi.nms_workspace_bytes(self.B, self.A, self.max_det)
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.out = torch.empty((self.B, self.max_det, OUT), dtype=torch.float32, device=self.device)
        self.counts = torch.empty((self.B,), dtype=torch.int32, device=self.device)
        self.kept_anchor = (torch.empty((self.B, self.max_det), dtype=torch.int32, device=self.device)
                            if want_anchor else None)
        # True while the last thing enqueued on this workspace was a pipelined step (whose K2 leaves the
        # candidate counters zeroed): only then may the next pipelined step skip its memset
        self.armed = False
        self.opts = None   # _abi.opts(...): per-call tuning / debug knobs (tests, tools); None = production
        self._counts_host = None

    def counts_to_host(self, counts: torch.Tensor | None = None) -> list:
        """The call's one host sync: ``counts`` -> a pinned buffer -> Python ints (a pinned target keeps the
        copy asynchronous and saves the pageable staging of ``.cpu()``)."""
        counts = self.counts if counts is None else counts
        if self._counts_host is None:
            self._counts_host = torch.empty((self.B,), dtype=torch.int32, pin_memory=True)
        self._counts_host.copy_(counts, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._counts_host.tolist()

    def run(self, pred: torch.Tensor, conf_thres: float, iou_thres: float, rescale: torch.Tensor | None = None,
            do_round: bool = False, out=None, counts=None):
        """Enqueue filter + NMS over ``pred[B,A,290]`` (CUDA, fp32 or fp16, contiguous)."""
        if pred.device != self.device or pred.dtype not in (torch.float32, torch.float16) or not pred.is_contiguous():
            raise ValueError("pred must be a contiguous fp32 / fp16 tensor on the plan's device")
        if tuple(pred.shape) != (self.B, self.A, ROW):
            raise ValueError(f"pred shape {tuple(pred.shape)} != {(self.B, self.A, ROW)}")
        out = self.out if out is None else out
        counts = self.counts if counts is None else counts
        self.armed = False
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _abi.call(_entry("lp_nms", pred), pred.data_ptr(), self.B, self.A, float(conf_thres), float(iou_thres),
                      self.max_det, self.max_nms, self.workspace.data_ptr(), self.workspace.numel(),
                      out.data_ptr(), counts.data_ptr(),
                      self.kept_anchor.data_ptr() if self.kept_anchor is not None else None,
                      rescale.data_ptr() if rescale is not None else None, int(bool(do_round)), stream, opts=self.opts)
        return out, counts

    def run_filter(self, pred: torch.Tensor, conf_thres: float):
        """Stage K1 only (lp_nms_filter_f32): candidates + counts are left in the workspace."""
        self.armed = False
        with torch.cuda.device(self.device):
            _abi.call(_entry("lp_nms_filter", pred), pred.data_ptr(), self.B, self.A, float(conf_thres),
                      self.workspace.data_ptr(), self.workspace.numel(),
                      torch.cuda.current_stream(self.device).cuda_stream, opts=self.opts)

    def run_suppress(self, pred: torch.Tensor, iou_thres: float, rescale=None, do_round=False, out=None, counts=None):
        """Stage K2 only (lp_nms_suppress_f32) on what :meth:`run_filter` left behind."""
        out = self.out if out is None else out
        counts = self.counts if counts is None else counts
        with torch.cuda.device(self.device):
            _abi.call(_entry("lp_nms_suppress", pred), pred.data_ptr(), self.B, self.A, float(iou_thres), self.max_det,
                      self.max_nms, self.workspace.data_ptr(), self.workspace.numel(), out.data_ptr(),
                      counts.data_ptr(), self.kept_anchor.data_ptr() if self.kept_anchor is not None else None,
                      rescale.data_ptr() if rescale is not None else None, int(bool(do_round)),
                      torch.cuda.current_stream(self.device).cuda_stream, opts=self.opts)
        return out, counts

    def candidate_counts(self) -> torch.Tensor:
        """Per-image candidate counts K1 left in the workspace (device int32 view)."""
        return self.workspace[: 4 * self.B].view(torch.int32)


class NmsPipeline:
    """Two-stream software pipeline over consecutive batches: K1 (filter, HBM-bound, all SMs it can
    get) of batch i+1 runs while K2 (sort/NMS/gather, latency-bound, one CTA per image) of batch i
    is still resolving.  ``depth`` plans alternate so a batch's candidates are not overwritten
    before its K2 has consumed them.  Results of ``submit`` are valid once ``done[slot]`` fired.
    """

    def __init__(self, B: int, A: int, max_det: int = 300, device=None, depth: int = 2, filter_streams: int = 2):
        self.plans = [NmsPlan(B, A, max_det, device) for _ in range(depth)]
        self.device = self.plans[0].device
        with torch.cuda.device(self.device):
            lo, hi = torch.cuda.Stream.priority_range()
            # K1 launches alternate between `filter_streams` streams: consecutive K1s then have no stream
            # order between them, so the first CTAs of batch i+1 move onto SMs as soon as K2 of batch
            # i-1 or K1 of batch i lets go of them -- K1's drain and ramp-up overlap instead of adding up
            # (cfg2: 48.2 -> 44.5 us per step).  They never share a workspace: depth >= filter_streams.
            assert 1 <= filter_streams <= depth
            self.s_filters = [torch.cuda.Stream(self.device, priority=lo) for _ in range(filter_streams)]
            self._fence = None   # filtered-event of a timed step: the next K1 must not start before it
            self.s_filter = self.s_filters[0]
            self.s_nms = torch.cuda.Stream(self.device, priority=hi)     # K2 CTAs are dispatched first
            self.filtered = [torch.cuda.Event() for _ in range(depth)]
            self.done = [torch.cuda.Event() for _ in range(depth)]
            for ev in self.filtered + self.done:   # force creation of the cudaEvent_t handles
                ev.record(self.s_nms)
        self.n = 0

    def start(self):
        """Order both streams after the caller's current stream."""
        cur = torch.cuda.current_stream(self.device)
        for sf in self.s_filters:
            sf.wait_stream(cur)
        self.s_nms.wait_stream(cur)

    def submit(self, pred: torch.Tensor, conf_thres: float, iou_thres: float, timing=None):
        """Enqueue one batch (one native call: lp_nms_pipelined_f32); returns (slot, out, counts).
        ``timing``: optional pair of CUDA events recorded round the K1 launch on its stream."""
        slot = self.n % len(self.plans)
        plan = self.plans[slot]
        s_filter = self.s_filters[self.n % len(self.s_filters)]
        if self._fence is not None:      # the previous step was a timed one: keep out of its way
            s_filter.wait_event(self._fence)
            self._fence = None
        if timing is not None:
            # A timed K1 is bracketed by events on its stream; for the bracket to be the kernel's own
            # duration it must neither queue behind the previous K1's CTAs nor share HBM with the next.
            if self.n > 0:
                s_filter.wait_event(self.filtered[(self.n - 1) % len(self.plans)])
            self._fence = self.filtered[slot]
            for ev in timing:  # torch creates the cudaEvent_t lazily, on the first record
                if ev.cuda_event == 0:
                    ev.record(s_filter)
        # the workspace counts as re-armed only once the step has been queued in full: a call that
        # raises leaves plan.armed False, so the next step on this workspace memsets its counters again
        armed, plan.armed = plan.armed, False
        _abi.call(_entry("lp_nms_pipelined", pred), pred.data_ptr(), plan.B, plan.A, float(conf_thres), float(iou_thres),
                  plan.max_det, plan.max_nms, plan.workspace.data_ptr(), plan.workspace.numel(), plan.out.data_ptr(),
                  plan.counts.data_ptr(), None, None, 0, s_filter.cuda_stream, self.s_nms.cuda_stream,
                  self.done[slot].cuda_event if armed else None,
                  self.filtered[slot].cuda_event, self.done[slot].cuda_event,
                  timing[0].cuda_event if timing is not None else None,
                  timing[1].cuda_event if timing is not None else None, opts=plan.opts)
        plan.armed = True
        self.n += 1
        return slot, plan.out, plan.counts

    def finish(self):
        """Make the caller's stream wait for everything submitted so far."""
        cur = torch.cuda.current_stream(self.device)
        for sf in self.s_filters:
            cur.wait_stream(sf)
        cur.wait_stream(self.s_nms)

    def capture(self, pred: torch.Tensor, conf_thres: float, iou_thres: float, steps: int) -> "GraphedSteps":
        """``steps`` pipelined steps over ``pred`` as one CUDA graph (see :class:`GraphedSteps`); this
        pipeline then belongs to the graph."""
        return GraphedSteps(self, lambda p: p.submit(pred, conf_thres, iou_thres), steps)


class GraphedSteps:
    """``steps`` consecutive pipelined steps captured ONCE into a CUDA graph and replayed with a
    single launch: the steady state of :class:`NmsPipeline` / ``PostprocessPipeline`` (K1 / KF
    alternating between the filter streams, K2 on the NMS stream, the events between them) becomes
    graph edges, so the GPU never waits for the host to submit the next step -- the per-step host
    cost (one Python -> ctypes call and ~6 CUDA API calls) is what capped 8-GPU scaling in round 1.

    The pipeline passed in is private to the graph from then on (its events were recorded under
    capture).  ``submit_one(pipe)`` must enqueue exactly one step on it.  Results of the last
    ``depth`` steps are in ``pipe.plans[i].out / .counts`` after a replay.
    """

    def __init__(self, pipe, submit_one, steps: int):
        self.pipe, self.steps = pipe, int(steps)
        dev = pipe.device
        for plan in pipe.plans:      # no event recorded outside the capture may be waited on inside it
            plan.armed = False
        pipe.n = 0
        with torch.cuda.device(dev):
            # once eagerly: lazy per-device initialisation (kernel attributes, tensor-map encoder) must
            # not happen under capture
            pipe.start()
            submit_one(pipe)
            pipe.finish()
            torch.cuda.synchronize(dev)
            for plan in pipe.plans:
                plan.armed = False
            pipe.n = 0
            self.graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(dev)
            with torch.cuda.graph(self.graph, stream=side, capture_error_mode="thread_local"):
                pipe.start()
                for _ in range(self.steps):
                    submit_one(pipe)
                pipe.finish()
        for plan in pipe.plans:      # eager use of these plans afterwards starts from a memset again
            plan.armed = False

    def launch(self):
        """Replay the ``steps`` steps on the current stream (one cudaGraphLaunch)."""
        self.graph.replay()


def rows_of(out: torch.Tensor, ks) -> list:
    """``[out[b, :k] for b, k in enumerate(ks)]`` as views of the padded ``[B, max_det, 28]`` buffer, built
    by ONE ``split_with_sizes`` instead of B Python indexing operations (which cost more host time than the
    kernels for a 32-image batch)."""
    B, M, W = out.shape
    sizes = [0] * (2 * B)
    sizes[0::2] = ks
    sizes[1::2] = [M - k for k in ks]
    return list(out.view(B * M, W).split_with_sizes(sizes)[0::2])


_plans: dict = {}


def _plan_for(B, A, max_det, device, want_anchor=False) -> NmsPlan:
    key = (B, A, max_det, device.index, torch.cuda.current_stream(device).cuda_stream, want_anchor)
    plan = _plans.get(key)
    if plan is None:
        if len(_plans) > 16:
            _plans.clear()
        plan = _plans[key] = NmsPlan(B, A, max_det, device, want_anchor=want_anchor)
    return plan


def _device_input(prediction: torch.Tensor) -> torch.Tensor:
    p = prediction
    if p.dtype not in (torch.float32, torch.float16):
        p = p.float()
    if not p.is_contiguous() or p.data_ptr() % 16:
        p = p.contiguous()
        if p.data_ptr() % 16:
            p = p.clone()
    return p


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, max_det=300):
    """Runs Non-Maximum Suppression on inference results (reference signature).

    ``classes``, ``agnostic`` and ``multi_label`` are accepted and ignored, as in
    the reference.  A CPU ``prediction`` is streamed to
    the current CUDA device in chunks (H2D overlapped with the kernels) and the
    detections come back as CPU tensors.
    """
    _check_thresholds(conf_thres, iou_thres)
    if prediction.dim() != 3 or prediction.shape[2] != ROW:
        raise ValueError(f"prediction must be [B, A, {ROW}], got {tuple(prediction.shape)}")
    B, A, _ = prediction.shape
    if B == 0:
        return []
    if A == 0 or max_det <= 0:
        return [torch.zeros((0, OUT), device=prediction.device)] * B
    if prediction.device.type == "cpu":
        from .host import host_pipeline
        dtype = torch.float16 if prediction.dtype == torch.float16 else torch.float32
        rows = host_pipeline(B, A, max_det, dtype).run(prediction, conf_thres, iou_thres)
        return [r.half() for r in rows] if dtype == torch.float16 else rows
    pred = _device_input(prediction)
    plan = _plan_for(B, A, int(max_det), pred.device)
    out = torch.empty((B, plan.max_det, OUT), dtype=torch.float32, device=pred.device)
    plan.run(pred, conf_thres, iou_thres, out=out)
    ks = plan.counts_to_host()  # the one host sync of the call
    if pred.dtype == torch.float16:   # rows in the prediction's dtype, like the reference's torch.cat
        out = out.half()
    return rows_of(out, ks)


def non_max_suppression_with_index(prediction, conf_thres=0.25, iou_thres=0.45, max_det=300):
    """Like :func:`non_max_suppression` but also returns the anchor index of every kept row
    (test / debugging aid; CUDA input only)."""
    _check_thresholds(conf_thres, iou_thres)
    pred = _device_input(prediction)
    B, A, _ = pred.shape
    plan = _plan_for(B, A, int(max_det), pred.device, want_anchor=True)
    out = torch.empty((B, plan.max_det, OUT), dtype=torch.float32, device=pred.device)
    plan.run(pred, conf_thres, iou_thres, out=out)
    ks = plan.counts_to_host()
    idx = plan.kept_anchor.clone()
    return rows_of(out, ks), [idx[b, :k].long() for b, k in enumerate(ks)]


def xywh2xyxy(x):
    """``yolov6/utils/nms.py:21-28`` for a CUDA tensor ``[n, 4]`` (any row stride)."""
    if not isinstance(x, torch.Tensor) or x.device.type != "cuda":
        raise RuntimeError("yolo_lp_b200.xywh2xyxy needs a CUDA tensor (no CPU fallback)")
    src = x if (x.dtype == torch.float32 and x.stride(-1) == 1) else x.float().contiguous()
    y = torch.empty((src.shape[0], 4), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        _abi.call("lp_xywh2xyxy_f32", src.data_ptr(), src.shape[0], src.stride(0), y.data_ptr(), 4,
                  torch.cuda.current_stream(src.device).cuda_stream)
    return y
