"""Image-sharded multi-GPU post-processing.

Images are independent (the loop body of nms.py:68 touches one image only), so
GPU ``g`` of ``G`` owns the contiguous range ``[g*B/G, (g+1)*B/G)`` end to end.
There is NO data-path collective: each GPU's detections (<= max_det*28 floats per
image) go D2H and the host concatenates the per-image lists in image order.

Two launch styles:
  * one process per GPU (``torchrun``): :func:`gather_detections` collects the
    per-rank lists on rank 0 through ``torch.distributed.gather_object`` (host-side
    control-plane traffic only -- works on gloo or nccl process groups);
  * one process driving several devices: :class:`ShardedNms`.
"""
from __future__ import annotations

import torch

from .synth import shard_range


def gather_detections(local: list, world_size: int, rank: int, group=None, dst: int = 0):
    """Concatenate per-rank detection lists in rank (== image) order on ``dst``.

    ``local`` is this rank's ``list[Tensor[k,28]]``; tensors are moved to host
    memory first.  Returns the full list on ``dst`` and ``None`` elsewhere.
    """
    import torch.distributed as dist
    local = [t.detach().cpu() for t in local]
    if world_size == 1:
        return local
    bucket = [None] * world_size if rank == dst else None
    dist.gather_object(local, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    out = []
    for part in bucket:
        out += part
    return out


class ShardedNms:
    """Single-process driver: one stream + plan per visible device, no NCCL."""

    def __init__(self, B: int, A: int, max_det: int = 300, devices=None):
        from .nms import NmsPlan
        devices = list(range(torch.cuda.device_count())) if devices is None else list(devices)
        if not devices:
            raise RuntimeError("no CUDA device")
        self.B, self.devices = B, devices
        self.ranges = [shard_range(B, g, len(devices)) for g in range(len(devices))]
        self.plans = []
        for d, (lo, hi) in zip(devices, self.ranges):
            with torch.cuda.device(d):
                self.plans.append(NmsPlan(hi - lo, A, max_det, torch.device("cuda", d)) if hi > lo else None)

    def run(self, shards, conf_thres, iou_thres):
        """``shards[g]``: ``[B_g, A, 290]`` resident on device ``g``.  Returns the detections of
        all images, in image order, as host tensors."""
        pending = []
        for plan, pred in zip(self.plans, shards):
            if plan is None:
                continue
            with torch.cuda.device(plan.device):
                out, counts = plan.run(pred, conf_thres, iou_thres)
                pending.append((out.to("cpu", non_blocking=True), counts.to("cpu", non_blocking=True), plan.device))
        res = []
        for out, counts, dev in pending:
            torch.cuda.synchronize(dev)
            res += [out[b, :k] for b, k in enumerate(counts.tolist())]
        return res
