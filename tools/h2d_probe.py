#!/usr/bin/env python3
"""Host->device copy rate of one rank's cfg2 batch (311.8 MB) from different kinds of pinned host
memory, with every rank copying at once (run under torchrun for N > 1):

    torch pinned (cudaHostAlloc default)  |  cudaHostAllocWriteCombined  |  cudaHostAllocPortable

Answers VERDICT r1 item 7: is the 8-GPU e2e limit (23 GB/s per rank) a property of the buffers?
"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NBYTES = 32 * 8400 * 1160
dst = torch.empty(NBYTES, dtype=torch.uint8, device=dev)
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]


def alloc(flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), NBYTES, flags)
    assert rc == 0, rc
    ctypes.memset(p, 1, NBYTES)
    return p


def rate(src_ptr, reps=8):
    s = torch.cuda.Stream(dev)
    rt.cudaMemcpyAsync(dst.data_ptr(), src_ptr, NBYTES, 1, s.cuda_stream)
    s.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(reps):
        rt.cudaMemcpyAsync(dst.data_ptr(), src_ptr, NBYTES, 1, s.cuda_stream)
    s.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return NBYTES / float(t.item()) / 1e9


host = torch.empty(NBYTES, dtype=torch.uint8, pin_memory=True).fill_(1)
res = {"torch_pinned": rate(host.data_ptr())}
for name, flags in (("default", 0), ("write_combined", 4), ("portable", 1), ("wc_portable", 5)):
    res["cudaHostAlloc_" + name] = rate(alloc(flags))
if int(os.environ.get("RANK", "0")) == 0:
    print(f"N={world}: GB/s per rank (slowest rank), all ranks copying at once:", {k: round(v, 1) for k, v in res.items()})
if world > 1:
    dist.destroy_process_group()
