"""CPU-only host logic: synthetic generator determinism, shard partitioning, rescale scalars,
the reference-signature shims' argument checks, and the world_size-2 (gloo) gather path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from yolo_lp_b200 import synth
from yolo_lp_b200.inferer import rescale_params
from oracle import lp_oracle


def test_synth_is_deterministic_and_shard_independent():
    cfg = synth.CONFIGS[2]
    a = synth.synth_head(3, 256, cfg["img"], 6, 20, seed=9, first_index=4)
    b = synth.synth_head(1, 256, cfg["img"], 6, 20, seed=9, first_index=5)
    assert torch.equal(a[1], b[0])
    assert not torch.equal(a[0], a[1])
    assert a.shape == (3, 256, 290) and a.dtype == torch.float32
    assert torch.all(a[..., 4] == 1.0)
    q = synth.synth_head(1, 256, cfg["img"], 6, 20, seed=9, quant=16)
    assert torch.equal(q[..., 13:] * 16, torch.round(q[..., 13:] * 16))


def test_synth_configs_match_baseline():
    assert synth.CONFIGS[2]["B"] == 32 and synth.CONFIGS[2]["A"] == 8400
    assert synth.CONFIGS[3]["B"] == 256
    assert synth.CONFIGS[4]["conf"] == 0.001 and synth.CONFIGS[4]["B"] == 64
    assert synth.CONFIGS[5]["A"] == 33600 == sum(h * w for h, w in synth.level_shapes(1280, 1280))
    assert 8400 == sum(h * w for h, w in synth.level_shapes(640, 640))


@pytest.mark.parametrize("B,G", [(256, 8), (256, 4), (32, 2), (7, 4), (3, 8), (1, 1)])
def test_shard_ranges_partition_the_batch(B, G):
    parts = [synth.shard_range(B, g, G) for g in range(G)]
    assert parts[0][0] == 0 and parts[-1][1] == B
    for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
        assert a1 == b0 and a0 <= a1
    assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1


def test_rescale_params_match_oracle():
    for ori, tgt in [((640, 416), (1160, 720, 3)), ((384, 640), (1080, 1920, 3)), ((640, 640), (640, 640, 3))]:
        px, py, r, w0, h0 = rescale_params(ori, tgt)
        r2, px2, py2 = lp_oracle.rescale_params(ori, tgt)
        assert (px, py, r) == (px2, py2, r2) and (w0, h0) == (tgt[1], tgt[0])


def test_shims_reject_cpu_tensors_loudly():
    import yolo_lp_b200 as y
    with pytest.raises(RuntimeError):
        y.rescale((640, 640), torch.zeros(3, 12), (640, 640, 3))
    with pytest.raises(RuntimeError):
        y.dist2bbox(torch.zeros(1, 8, 4), torch.zeros(8, 2))
    with pytest.raises(RuntimeError):
        y.generate_anchors([torch.zeros(1, 1, 2, 2)], [8], device="cpu", is_eval=True)
    with pytest.raises(NotImplementedError):
        y.generate_anchors([torch.zeros(1, 1, 2, 2)], [8], device="cuda", is_eval=False)
    with pytest.raises(AssertionError):
        y.non_max_suppression(torch.zeros(1, 8, 290), conf_thres=1.5)
    with pytest.raises(ValueError):
        y.non_max_suppression(torch.zeros(1, 8, 289))
    assert y.non_max_suppression(torch.zeros(0, 8, 290)) == []


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_worker(rank, world, port, B, q):
    import torch.distributed as dist
    from yolo_lp_b200.shard import gather_detections
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = synth.shard_range(B, rank, world)
    # this rank's "detections": the oracle on its own image shard (CPU stand-in for the kernels)
    pred = synth.synth_head(hi - lo, 64, 128, 4, 10, seed=11, first_index=lo).numpy()
    local = [torch.from_numpy(r) for r in lp_oracle.non_max_suppression(pred, 0.2, 0.45)]
    full = gather_detections(local, world, rank)
    if rank == 0:
        q.put([t.numpy() for t in full])
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_preserves_image_order():
    B, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    pred = synth.synth_head(B, 64, 128, 4, 10, seed=11).numpy()
    want = lp_oracle.non_max_suppression(pred, 0.2, 0.45)
    assert len(got) == B
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_bench_reference_arm_prints_one_json_line_on_cpu():
    """`bench.py --impl reference` is CPU-only (the staged reference, else its torch port, on all host threads): it must run
    without a GPU and put exactly one JSON line with the contract's keys on stdout."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "post-proc images/sec" and d["unit"] == "images/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1
    from oracle import stage_ref
    assert d["cpu_baseline"]["kind"] == ("reference" if stage_ref.is_staged() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("cfg2")
