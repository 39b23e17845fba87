"""Host-buffer path: ``non_max_suppression`` on a CPU head tensor.

The input lives in host memory (pinned for full PCIe rate); images are streamed
to the GPU in chunks on a copy stream while the previous chunk is being filtered
and suppressed on the compute stream, and the (small) detections are copied back
into pinned buffers.  This is the ``e2e`` leg of ``bench.py``: H2D and D2H are
inside the call.  It is not a CPU fallback -- every image still goes through the
CUDA kernels.
"""
from __future__ import annotations

import torch

from . import _abi
from .nms import NmsPlan, OUT, ROW, rows_of

_CHUNK_BYTES = 48 << 20  # ~48 MiB per H2D chunk: long enough for full PCIe rate, short pipeline fill


class HostPipeline:
    def __init__(self, B: int, A: int, max_det: int, device=None, chunk_images: int | None = None,
                 dtype: torch.dtype = torch.float32):
        _abi.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.B, self.A, self.max_det = B, A, max_det
        self.dtype = dtype                      # fp16 head tensors travel and are read as halves
        per_image = A * ROW * (2 if dtype == torch.float16 else 4)
        c = chunk_images or max(1, min(B, _CHUNK_BYTES // per_image))
        self.chunk = c
        self.n_chunks = (B + c - 1) // c
        with torch.cuda.device(self.device):
            self.stage = [torch.empty((c, A, ROW), dtype=dtype, device=self.device) for _ in range(2)]
            self.plans = [NmsPlan(c, A, max_det, self.device) for _ in range(2)]
            self.tail_plan = None
            if B % c:
                self.tail_plan = NmsPlan(B % c, A, max_det, self.device)
            self.out_dev = torch.empty((B, max_det, OUT), dtype=torch.float32, device=self.device)
            self.counts_dev = torch.empty((B,), dtype=torch.int32, device=self.device)
            self.copy_stream = torch.cuda.Stream(self.device)
            self.compute_stream = torch.cuda.Stream(self.device)
            self.h2d_done = [torch.cuda.Event() for _ in range(2)]
            self.stage_free = [torch.cuda.Event() for _ in range(2)]
        self.out_host = torch.empty((B, max_det, OUT), dtype=torch.float32, pin_memory=True)
        self.counts_host = torch.empty((B,), dtype=torch.int32, pin_memory=True)
        self.h2d_bytes = B * per_image
        self.d2h_bytes = self.out_host.numel() * 4 + self.counts_host.numel() * 4

    def run(self, prediction: torch.Tensor, conf_thres: float, iou_thres: float):
        """prediction: CPU fp32 (or fp16, for a pipeline built with ``dtype=torch.float16``)
        ``[B, A, 290]``.  Returns a list of CPU tensors ``[k_b, 28]``."""
        if prediction.dtype != self.dtype or not prediction.is_contiguous():
            prediction = prediction.to(self.dtype).contiguous()
        c = self.chunk
        with torch.cuda.device(self.device):
            caller = torch.cuda.current_stream(self.device)
            self.copy_stream.wait_stream(caller)
            self.compute_stream.wait_stream(caller)
            for i in range(self.n_chunks):
                s = i & 1
                lo, hi = i * c, min(self.B, (i + 1) * c)
                n = hi - lo
                with torch.cuda.stream(self.copy_stream):
                    if i >= 2:
                        self.copy_stream.wait_event(self.stage_free[s])
                    self.stage[s][:n].copy_(prediction[lo:hi], non_blocking=True)
                    self.h2d_done[s].record(self.copy_stream)
                with torch.cuda.stream(self.compute_stream):
                    self.compute_stream.wait_event(self.h2d_done[s])
                    plan = self.plans[s] if n == c else self.tail_plan
                    plan.run(self.stage[s][:n], conf_thres, iou_thres,
                             out=self.out_dev[lo:hi], counts=self.counts_dev[lo:hi])
                    self.stage_free[s].record(self.compute_stream)
            with torch.cuda.stream(self.compute_stream):
                self.counts_host.copy_(self.counts_dev, non_blocking=True)
                self.out_host.copy_(self.out_dev, non_blocking=True)
            self.compute_stream.synchronize()
        ks = self.counts_host.tolist()
        # the pinned buffer is reused by the next call: hand out copies -- ONE cat of the used rows and one
        # split instead of B clones (which cost ~2 % of a PCIe-bound batch)
        if sum(ks) == 0:
            return [self.out_host.new_zeros((0, OUT)) for _ in ks]
        return list(torch.cat(rows_of(self.out_host, ks)).split_with_sizes(ks))


_pipes: dict = {}


def host_pipeline(B: int, A: int, max_det: int, dtype: torch.dtype = torch.float32) -> HostPipeline:
    key = (B, A, max_det, torch.cuda.current_device(), dtype)
    pipe = _pipes.get(key)
    if pipe is None:
        if len(_pipes) > 4:
            _pipes.clear()
        pipe = _pipes[key] = HostPipeline(B, A, max_det, dtype=dtype)
    return pipe
