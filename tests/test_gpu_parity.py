"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against
(1) the golden outputs of the reference and (2) the oracle on seeded inputs.

Bar (north_star): kept-detection indices and counts bit-exact; box/score/keypoint values
within 1e-5 relative -- here asserted BIT-EXACT for everything on the NMS path, and within
1e-5 relative only for the sigmoid columns of the decode stage (GPU expf != CPU expf).
"""
import numpy as np
import pytest
import torch

import yolo_lp_b200 as lp
from yolo_lp_b200 import synth
from yolo_lp_b200.nms import non_max_suppression_with_index, NmsPlan
from oracle import lp_oracle
from _util import (golden, golden_names, split_rows, seeded_inputs, assert_rows_equal, iou_band_pairs,
                   IOU_BAND_THRESHOLDS)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SEEDED = golden_names("nms_cfg") + golden_names("nms_eval")
EDGES = golden_names("nms_edge_")
DECODE = [n for n in golden_names("decode_") if "half" not in n]


def _knobs(g):
    return float(g["conf"]), float(g["iou"]), int(g["max_det"])


def _gpu_nms(pred, conf, iou, max_det):
    rows, idx = non_max_suppression_with_index(pred.to(DEV), conf, iou, max_det)
    return [r.cpu().numpy() for r in rows], [i.cpu().numpy() for i in idx]


def _check_against_oracle(pred, conf, iou, max_det, what):
    got, gidx = _gpu_nms(pred, conf, iou, max_det)
    want, widx = lp_oracle.non_max_suppression(pred.numpy(), conf, iou, max_det=max_det, return_index=True)
    assert len(got) == len(want)
    for b in range(len(want)):
        assert np.array_equal(gidx[b], widx[b]), f"{what}[{b}]: kept anchor indices differ"
        assert_rows_equal(got[b], want[b], f"{what}[{b}]")
    return got


# ------------------------------------------------------------------ reference goldens
@pytest.mark.parametrize("name", SEEDED)
def test_nms_seeded_goldens(name):
    g = golden(name)
    conf, iou, max_det = _knobs(g)
    pred = seeded_inputs(g)
    got = _check_against_oracle(pred, conf, iou, max_det, name)
    for b, w in enumerate(split_rows(g["counts"], g["rows"])):
        assert_rows_equal(got[b], w, f"{name}[{b}] vs reference")


@pytest.mark.parametrize("name", EDGES)
def test_nms_edge_goldens(name):
    g = golden(name)
    conf, iou, max_det = _knobs(g)
    pred = torch.from_numpy(g["pred"])
    pred = pred[None] if pred.dim() == 2 else pred
    got = _check_against_oracle(pred, conf, iou, max_det, name)
    for b, w in enumerate(split_rows(g["counts"], g["rows"])):
        assert_rows_equal(got[b], w, f"{name}[{b}] vs reference")


# ------------------------------------------------------------------ oracle on seeded random inputs
@pytest.mark.parametrize("B,A,n_pos,conf,iou,max_det,quant", [
    (1, 1, 1, 0.0, 0.45, 300, None),        # a single anchor
    (3, 31, 8, 0.05, 0.45, 300, None),      # A < one tile, tiles straddle images, odd B*A
    (5, 33, 10, 0.0, 0.5, 300, 8),          # odd everything, mass ties
    (2, 525, 60, 0.1, 0.45, 300, None),     # 160x160 input
    (4, 2100, 150, 0.25, 0.45, 300, None),  # 320x320
    (2, 2100, 150, 0.0, 0.65, 50, 16),      # dense + ties + max_det cut
    (2, 5040, 300, 0.03, 0.65, 300, None),  # 384x640 letterbox (non-square)
    (1, 8400, 300, 0.0, 0.45, 1000, None),  # all 8400 pass, max_det 1000 (tools/infer.py default)
    (1, 8400, 300, 0.0, 0.3, 2000, 4),      # > KEPT_SMEM kept boxes not reached but > 1 window
])
def test_nms_vs_oracle(B, A, n_pos, conf, iou, max_det, quant):
    pred = synth.synth_head(B, A, 640, 12, n_pos, seed=100 + A, quant=quant)
    _check_against_oracle(pred, conf, iou, max_det, f"B{B}A{A}")


def test_many_kept_boxes_spill_past_shared_kept_cache():
    # widely spread tiny boxes: nothing overlaps, every candidate is kept -> > 1024 kept rows
    A = 3000
    pred = synth.synth_head(1, A, 640, 12, 100, seed=5)
    g = torch.Generator().manual_seed(3)
    pred[0, :, 0:2] = torch.rand((A, 2), generator=g) * 4000.0
    pred[0, :, 2:4] = 1.0
    got = _check_against_oracle(pred, 0.0, 0.45, 2500, "spread")
    assert got[0].shape[0] == 2500


def test_global_sort_path_and_max_nms_cut():
    # more candidates than the 16384-key shared-memory sort holds; the plan's max_nms cut is
    # exercised with a small max_nms (the reference's 30000 needs A > 30000 dense)
    A = 20000
    pred = synth.synth_head(1, A, 1280, 48, 500, seed=8, quant=32)
    dev = pred.to(DEV)
    want, widx = lp_oracle.nms_one_image(pred[0].numpy(), 0.0, 0.45, max_det=300)
    rows, idx = non_max_suppression_with_index(dev, 0.0, 0.45, 300)
    assert np.array_equal(idx[0].cpu().numpy(), widx)
    assert_rows_equal(rows[0].cpu().numpy(), want, "global sort")
    plan = NmsPlan(1, A, 300, torch.device(DEV), max_nms=5000, want_anchor=True)
    out, counts = plan.run(dev, 0.0, 0.45)
    k = int(counts.cpu()[0])
    want, widx = lp_oracle.nms_one_image(pred[0].numpy(), 0.0, 0.45, max_det=300, max_nms=5000)
    assert np.array_equal(plan.kept_anchor[0, :k].cpu().numpy(), widx)
    assert_rows_equal(out[0, :k].cpu().numpy(), want, "max_nms cut")


# ------------------------------------------------------------------ BASELINE full sizes: properties
def _iou_matrix(b):
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.maximum(b[:, None, :2], b[None, :, :2])
    rb = torch.minimum(b[:, None, 2:4], b[None, :, 2:4])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area[:, None] + area[None, :] - inter)


@pytest.mark.parametrize("cid", [2, 3, 4, 5])
def test_full_size_properties(cid):
    cfg = synth.CONFIGS[cid]
    B = cfg["B"] if cid != 3 else 64          # config 3 is 256 images sharded over GPUs: one shard's worth x2
    first = 0 if cid != 3 else 96
    pred = synth.synth_head(B, cfg["A"], cfg["img"], cfg["n_plates"], cfg["n_pos"], cfg["seed"], first_index=first)
    dev = pred.to(DEV)
    before = dev.clone()
    rows = lp.non_max_suppression(dev, cfg["conf"], cfg["iou"], max_det=cfg["max_det"])
    assert torch.equal(dev, before), "prediction must not be mutated"
    assert len(rows) == B
    # determinism + per-image independence: any sub-batch gives the same rows
    again = lp.non_max_suppression(dev, cfg["conf"], cfg["iou"], max_det=cfg["max_det"])
    sub = lp.non_max_suppression(dev[B // 2:B // 2 + 3].clone(), cfg["conf"], cfg["iou"], max_det=cfg["max_det"])
    for b in range(B):
        assert torch.equal(rows[b], again[b])
    for j in range(3):
        assert torch.equal(rows[B // 2 + j], sub[j])
    for b in range(B):
        r = rows[b]
        k = r.shape[0]
        assert 0 < k <= cfg["max_det"]
        score = r[:, 12:20].cpu().numpy().astype(np.float32)
        s = lp_oracle._sum8(score, 7) / np.float32(8)
        assert np.all(s[:-1] >= s[1:]), "rows must be in decreasing score order"
        iou = _iou_matrix(r[:, :4])
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= cfg["iou"] + 1e-6, "two kept boxes overlap above the threshold"
        arg = r[:, 20:28]
        assert torch.all(arg == arg.round()) and torch.all(arg >= 0) and torch.all(arg < 37)
    # exact parity on a sample of the batch (the oracle needs ~0.1-1 s per image)
    for b in sorted({0, B // 3, B - 1}):
        want, _ = lp_oracle.nms_one_image(pred[b].numpy(), cfg["conf"], cfg["iou"], max_det=cfg["max_det"])
        assert_rows_equal(rows[b].cpu().numpy(), want, f"cfg{cid}[{b}]")


# ------------------------------------------------------------------ API behaviour
def test_input_views_and_dtypes():
    pred = synth.synth_head(3, 525, 160, 6, 40, seed=21)
    want = lp_oracle.non_max_suppression(pred.numpy(), 0.1, 0.45)
    dev = pred.to(DEV)
    wide = torch.zeros((3, 525, 300), device=DEV)
    wide[..., 5:295] = dev
    for variant in (dev[:, :, :], wide[..., 5:295], dev.double(), dev[1:3]):
        got = lp.non_max_suppression(variant, 0.1, 0.45)
        off = 1 if variant.shape[0] == 2 else 0
        for b, r in enumerate(got):
            assert r.device.type == "cuda" and r.dtype == torch.float32
            assert_rows_equal(r.cpu().numpy(), want[b + off], "view")


def test_thresholds_assert_like_reference():
    dev = torch.zeros((1, 64, 290), device=DEV)
    with pytest.raises(AssertionError):
        lp.non_max_suppression(dev, conf_thres=-0.1)
    with pytest.raises(AssertionError):
        lp.non_max_suppression(dev, iou_thres=1.01)
    out = lp.non_max_suppression(dev, 0.25, 0.45, classes=[0], agnostic=True, multi_label=True)
    assert len(out) == 1 and tuple(out[0].shape) == (0, 28)


def test_host_buffer_path_matches_device_path():
    pred = synth.synth_head(7, 2100, 320, 8, 100, seed=33)
    want = lp_oracle.non_max_suppression(pred.numpy(), 0.2, 0.45)
    from yolo_lp_b200.host import HostPipeline
    pipe = HostPipeline(7, 2100, 300, chunk_images=3)     # 3 chunks incl. a ragged tail
    for pinned in (False, True):
        src = pred.pin_memory() if pinned else pred
        got = pipe.run(src, 0.2, 0.45)
        for b in range(7):
            assert got[b].device.type == "cpu"
            assert_rows_equal(got[b].numpy(), want[b], f"host[{b}]")
    got = lp.non_max_suppression(pred, 0.2, 0.45)          # public API with a CPU tensor
    for b in range(7):
        assert_rows_equal(got[b].numpy(), want[b], f"api-host[{b}]")


# ------------------------------------------------------------------ rescale
def test_rescale_goldens():
    g = golden("rescale")
    for n in range(int(g["n"])):
        hi, wi, h0, w0 = g[f"shape{n}"].tolist()
        src = torch.from_numpy(g[f"in{n}"]).to(DEV)
        t = src.clone()
        ret = lp.rescale((hi, wi), t, (h0, w0, 3))
        assert ret is t
        assert np.array_equal(t.cpu().numpy().view(np.uint32), g[f"out{n}"].view(np.uint32)), n
        # on a [k,28] row view, as Inferer.infer calls it (inferer.py:100), with the fused round
        det = torch.zeros((src.shape[0], 28), device=DEV)
        det[:, :12] = src
        det[:, 12:] = 7.0
        lp.rescale((hi, wi), det[:, :12], (h0, w0, 3), do_round=True)
        assert np.array_equal(det[:, :12].cpu().numpy(), g[f"round{n}"]), n
        assert torch.all(det[:, 12:] == 7.0)


def test_fused_rescale_in_nms_matches_separate_call():
    from yolo_lp_b200.inferer import rescale_table, rescale_batch
    pred = synth.synth_head(4, 2100, 320, 8, 100, seed=44)
    dev = pred.to(DEV)
    ori = [(320, 320), (320, 320), (320, 192), (192, 320)]
    tgt = [(1160, 720, 3), (640, 640, 3), (1080, 608, 3), (375, 1242, 3)]
    plan = NmsPlan(4, 2100, 300, torch.device(DEV))
    out_a, cnt_a = plan.run(dev, 0.2, 0.45)
    out_a, cnt_a = out_a.clone(), cnt_a.clone()
    rescale_batch(out_a, cnt_a, ori, tgt, do_round=True)
    out_b, cnt_b = plan.run(dev, 0.2, 0.45, rescale=rescale_table(ori, tgt, DEV), do_round=True)
    want = lp_oracle.non_max_suppression(pred.numpy(), 0.2, 0.45)
    for b, k in enumerate(cnt_b.cpu().tolist()):
        assert k == want[b].shape[0] == int(cnt_a[b])
        assert torch.equal(out_a[b, :k], out_b[b, :k])
        ref = want[b].copy()
        ref[:, :12] = lp_oracle.rescale(ori[b], ref[:, :12], tgt[b], do_round=True)
        assert_rows_equal(out_b[b, :k].cpu().numpy(), ref, f"fused rescale[{b}]")


# ------------------------------------------------------------------ geometry + decode
def test_geometry_goldens():
    g = golden("geometry")
    feats = [torch.zeros((1, 1, h, w), device=DEV) for h, w in g["hw"].tolist()]
    ap, st = lp.generate_anchors(feats, torch.tensor([8, 16, 32]), 5.0, 0.5, device=DEV, is_eval=True, mode="af")
    assert np.array_equal(ap.cpu().numpy(), g["anchor_points"]) and np.array_equal(st.cpu().numpy(), g["stride_tensor"])
    dist, cdist = torch.from_numpy(g["dist"]).to(DEV), torch.from_numpy(g["cdist"]).to(DEV)
    assert np.array_equal(lp.dist2bbox(dist, ap, "xyxy").cpu().numpy(), g["bbox_xyxy"])
    assert np.array_equal(lp.dist2bbox(dist, ap, "xywh").cpu().numpy(), g["bbox_xywh"])
    assert np.array_equal(lp.dist2cor(cdist, ap).cpu().numpy(), g["corners"])
    x = torch.rand((100, 4), device=DEV) * 50
    assert np.array_equal(lp.xywh2xyxy(x).cpu().numpy(), lp_oracle.xywh2xyxy(x.cpu().numpy()))


def _levels(g, device):
    return [{k: torch.from_numpy(g[f"l{l}_{k}"]).to(device)
             for k in ("pro", "alp", "ad0", "ad1", "ad2", "ad3", "ad4", "ad5", "reg", "cor")} for l in range(3)]


@pytest.mark.parametrize("name", DECODE)
def test_decode_goldens(name):
    g = golden(name)
    want = g["out"]
    got = lp.detect_decode(_levels(g, DEV), (8, 16, 32)).cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got[..., :13].view(np.uint32), want[..., :13].view(np.uint32)), "box/obj/corner columns"
    np.testing.assert_allclose(got[..., 13:], want[..., 13:], rtol=1e-5, atol=0)   # sigmoid: 1e-5 relative
    # stage-wise protocol: feed the GPU-decoded tensor to both NMS implementations
    conf, iou, max_det = _knobs(g)
    _check_against_oracle(torch.from_numpy(got), conf, iou, max_det, name + " decode->nms")


def test_decode_half_mode_golden():
    """The reference's model.half() forward (golden made by running the unmodified Detect in half on the
    CPU, tests/golden/make_golden.py make_decode_half): half conv outputs in, fp32 head tensor out whose
    class scores are sigmoids rounded to half.  lp_detect_decode_half_scores_f32: geometry columns
    bit-exact; scores equal to the reference's except where the 1e-5 sigmoid difference straddles a
    half rounding boundary (then one half ulp apart); then the stage-wise NMS protocol."""
    g = golden("decode_half_96x160")
    want = g["out"]
    levels = _levels(g, DEV)
    assert all(v.dtype == torch.float16 for lv in levels for v in lv.values())
    got = lp.detect_decode(levels, (8, 16, 32), half_scores=True)
    assert got.dtype == torch.float32
    got = got.cpu().numpy()
    assert np.array_equal(got[..., :13].view(np.uint32), want[..., :13].view(np.uint32)), "box/obj/corner columns"
    sc, ws = got[..., 13:], want[..., 13:]
    assert np.array_equal(sc, sc.astype(np.float16).astype(np.float32)), "scores are not half-representable"
    assert (sc == ws).mean() >= 0.99
    np.testing.assert_allclose(sc, ws, rtol=2.0 ** -10, atol=0)          # at most one half ulp
    conf, iou, max_det = _knobs(g)
    _check_against_oracle(torch.from_numpy(got), conf, iou, max_det, "half decode->nms")
    # ... and the plain entry on the same (upcast) tensors differs from it by the rounding only
    plain = lp.detect_decode(levels, (8, 16, 32)).cpu().numpy()
    assert np.array_equal(plain[..., :13], got[..., :13])
    assert np.array_equal(plain[..., 13:].astype(np.float16).astype(np.float32), sc)


def test_decode_full_size_vs_oracle():
    torch.manual_seed(0)
    B, widths = 2, (31, 24, 37, 37, 37, 37, 37, 37)
    names = ("pro", "alp", "ad0", "ad1", "ad2", "ad3", "ad4", "ad5")
    levels = []
    for h, w in synth.level_shapes(384, 640):
        lv = {n: torch.randn(B, c, h, w) * 3 for n, c in zip(names, widths)}
        lv["reg"] = torch.rand(B, 4, h, w) * 8
        lv["cor"] = torch.rand(B, 8, h, w) * 8 - 1
        levels.append(lv)
    want = lp_oracle.detect_decode([{k: v.numpy() for k, v in lv.items()} for lv in levels], (8, 16, 32))
    got = lp.detect_decode([{k: v.to(DEV) for k, v in lv.items()} for lv in levels], (8, 16, 32)).cpu().numpy()
    assert got.shape == (B, 5040, 290)
    assert np.array_equal(got[..., :13].view(np.uint32), want[..., :13].view(np.uint32))
    np.testing.assert_allclose(got[..., 13:], want[..., 13:], rtol=1e-5, atol=0)


@pytest.mark.parametrize("B,H,W", [(32, 640, 640), (3, 1280, 1280), (5, 416, 640), (1, 64, 96), (2, 608, 608)])
def test_decode_tma_path_equals_lsu_path_and_repeats(B, H, W):
    """The warp-specialised TMA kernel (decode_tma.cu) and the cp.async kernel (decode.cu, forced with
    the per-call knob lp_opts_t.no_tma) share the arithmetic: bit-identical outputs; 20 repeats of the TMA kernel over a
    poisoned output catch ring / barrier races.  608x608 has a 19x19 level (rows not 16-byte aligned):
    both settings then run the 4-byte cp.async path."""
    from yolo_lp_b200 import _abi
    levels = synth.synth_levels(B, H, W, DEV, seed=B + H)
    lsu = lp.DecodePlan(levels, (8, 16, 32))
    lsu.opts = _abi.opts(no_tma=True)
    ref = lsu.run().clone()
    plan = lp.DecodePlan(levels, (8, 16, 32))
    for _ in range(20):
        plan.out.fill_(float("nan"))
        got = plan.run()
        assert torch.equal(got.view(torch.int32), ref.view(torch.int32))


@pytest.mark.parametrize("B,H,W,conf", [(32, 640, 640, 0.25), (40, 640, 640, 0.001), (6, 1280, 1280, 0.25)])
def test_fused_tma_path_equals_lsu_path_many_tiles_per_cta(B, H, W, conf):
    """Shapes with dozens of tiles per persistent CTA, so the ring, the per-slot barriers and the
    producer / scanner / finisher hand-offs of fused_tma.cu all wrap many times.  First run on a
    poisoned workspace (a cold workspace makes the finishers slow -- that once let a fast finisher
    take over a barrier that belonged to a slow one), then back-to-back repeats; reference = the
    register-resident kernel of fused.cu (forced with lp_opts_t.no_tma), itself pinned to
    decode -> K1 -> K2 by test_fused_postprocess_equals_decode_then_nms."""
    from yolo_lp_b200 import _abi
    levels = synth.synth_levels(B, H, W, DEV, seed=7)
    ref = lp.PostprocessPlan(levels, (8, 16, 32), 300)
    ref.opts = _abi.opts(no_tma=True)
    ref_out, ref_counts = ref.run(conf, 0.45)
    torch.cuda.synchronize()

    def same(out, counts):
        return torch.equal(counts, ref_counts) and all(
            torch.equal(out[b, :int(counts[b])], ref_out[b, :int(counts[b])]) for b in range(B))

    for poison in (0x00, 0xFF):
        plan = lp.PostprocessPlan(levels, (8, 16, 32), 300)
        plan.workspace.fill_(poison)
        out, counts = plan.run(conf, 0.45)
        torch.cuda.synchronize()
        assert same(out, counts), f"first run on a workspace filled with {poison:#x}"
    for _ in range(10):
        out, counts = plan.run(conf, 0.45)
    torch.cuda.synchronize()
    assert same(out, counts)


def test_detect_forward_eval_runs_module_convs_then_kernel():
    """A stand-in module with the reference Detect's attribute names (the reference itself is
    not on the GPU box): convs run in torch, the tail in the kernel."""
    import torch.nn as nn
    from yolo_lp_b200.head import detect_forward_eval, CLS_NAMES, CLS_WIDTH
    torch.manual_seed(1)
    chans = (16, 32, 64)

    class Head(nn.Module):
        def __init__(self):
            super().__init__()
            self.nl, self.use_dfl, self.stride = 3, False, torch.tensor([8, 16, 32])
            mk = lambda f: nn.ModuleList([f(c) for c in chans])
            self.stems = mk(lambda c: nn.Conv2d(c, c, 1))
            self.cls_convs = mk(lambda c: nn.Conv2d(c, c, 3, padding=1))
            self.reg_convs = mk(lambda c: nn.Conv2d(c, c, 3, padding=1))
            for n, wd in zip(CLS_NAMES, CLS_WIDTH):
                setattr(self, n + "_preds", mk(lambda c, wd=wd: nn.Conv2d(c, wd, 1)))
            self.reg_preds = mk(lambda c: nn.Conv2d(c, 4, 1))
            self.cor_preds = mk(lambda c: nn.Conv2d(c, 8, 1))

    head = Head().to(DEV).eval()
    feats = [torch.rand(2, c, 96 // s, 160 // s, device=DEV) for c, s in zip(chans, (8, 16, 32))]
    with torch.no_grad():
        out = detect_forward_eval(head, feats)
        levels = []
        for i in range(3):
            f = head.stems[i](feats[i])
            cf, rf = head.cls_convs[i](f), head.reg_convs[i](f)
            lv = {n: getattr(head, n + "_preds")[i](cf).cpu().numpy() for n in CLS_NAMES}
            lv["reg"], lv["cor"] = head.reg_preds[i](rf).cpu().numpy(), head.cor_preds[i](rf).cpu().numpy()
            levels.append(lv)
    want = lp_oracle.detect_decode(levels, (8, 16, 32))
    assert tuple(out.shape) == (2, 315, 290)
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-5, atol=0)


# ------------------------------------------------------------------ two-stream pipeline
def test_pipeline_matches_serial_path_on_alternating_batches():
    """K1 of batch i+1 overlapping K2 of batch i must not change any result: two different
    batches are submitted alternately through the depth-2 pipeline and compared with the oracle."""
    from yolo_lp_b200.nms import NmsPipeline
    preds = [synth.synth_head(6, 2100, 320, 8, 120, seed=s) for s in (51, 52, 53)]
    wants = [lp_oracle.non_max_suppression(p.numpy(), 0.15, 0.5) for p in preds]
    devs = [p.to(DEV) for p in preds]
    pipe = NmsPipeline(6, 2100, 300, torch.device(DEV))
    pipe.start()
    results = []
    for i in range(9):
        slot, out, counts = pipe.submit(devs[i % 3], 0.15, 0.5)
        pipe.done[slot].synchronize()          # this batch's results are final once its K2 is done
        results.append((i % 3, out.clone(), counts.clone()))
    pipe.finish()
    torch.cuda.synchronize()
    for which, out, counts in results:
        for b, k in enumerate(counts.cpu().tolist()):
            assert_rows_equal(out[b, :k].cpu().numpy(), wants[which][b], f"pipeline batch {which}[{b}]")
    # K2 of a pipelined step re-arms the workspace (zeroed candidate counters + tile counter), which is
    # what lets the next pipelined step on it skip the memset node; any other entry still zeroes itself
    for plan in pipe.plans:
        assert int(plan.workspace[: 4 * (plan.B + 1)].view(torch.int32).abs().sum()) == 0
    out, counts = pipe.plans[0].run(devs[1], 0.15, 0.5)
    for b, k in enumerate(counts.cpu().tolist()):
        assert_rows_equal(out[b, :k].cpu().numpy(), wants[1][b], f"serial call on a pipeline workspace [{b}]")
    # ... and after that serial call (which leaves the counters dirty) the pipeline must zero them again
    pipe.start()
    for i in range(4):
        slot, out, counts = pipe.submit(devs[2], 0.15, 0.5)
    pipe.finish()
    torch.cuda.synchronize()
    for b, k in enumerate(counts.cpu().tolist()):
        assert_rows_equal(out[b, :k].cpu().numpy(), wants[2][b], f"pipeline after a serial call [{b}]")


def test_filter_cta_limit_does_not_change_results():
    from yolo_lp_b200 import _abi
    pred = synth.synth_head(3, 8400, 640, 24, 300, seed=61)
    want = lp_oracle.non_max_suppression(pred.numpy(), 0.25, 0.45)
    dev = pred.to(DEV)
    plan = lp.NmsPlan(3, 8400, 300, DEV)
    for ctas in (1, 7, 148, 0):
        plan.opts = _abi.opts(filter_ctas=ctas)
        out, counts = plan.run(dev, 0.25, 0.45)
        for b, k in enumerate(counts.cpu().tolist()):
            assert_rows_equal(out[b, :k].cpu().numpy(), want[b], f"ctas={ctas}[{b}]")


# ------------------------------------------------------------------ fused raw-levels -> detections
def _random_levels(B, H, W, seed, scale=3.0, shift=-2.0):
    g = torch.Generator().manual_seed(seed)
    names, widths = ("pro", "alp", "ad0", "ad1", "ad2", "ad3", "ad4", "ad5"), (31, 24, 37, 37, 37, 37, 37, 37)
    levels = []
    for h, w in synth.level_shapes(H, W):
        lv = {n: torch.randn((B, c, h, w), generator=g) * scale + shift for n, c in zip(names, widths)}
        lv["reg"] = torch.rand((B, 4, h, w), generator=g) * 6
        lv["cor"] = torch.rand((B, 8, h, w), generator=g) * 6 - 1
        levels.append(lv)
    return levels


@pytest.mark.parametrize("B,H,W,conf,iou,max_det,quant", [
    (2, 96, 160, 0.05, 0.45, 300, None),    # A = 315 (odd), ragged tiles
    (3, 320, 320, 0.10, 0.45, 300, None),
    (2, 384, 640, 0.02, 0.65, 300, None),   # non-square letterbox, dense
    (1, 640, 640, 0.0, 0.5, 300, None),     # every anchor passes -> segmented ordering
    (2, 160, 160, 0.0, 0.5, 50, 0.5),       # logits quantised to 0.5 -> mass ties in sigmoid space
])
def test_fused_postprocess_equals_decode_then_nms(B, H, W, conf, iou, max_det, quant):
    """Stage-wise protocol of SURVEY §7: the fused path must give bit-identical detections to our
    own decode kernel followed by K1 + K2 on the same level tensors."""
    from yolo_lp_b200.head import PostprocessPlan
    levels = _random_levels(B, H, W, seed=H + W)
    if quant:
        for lv in levels:
            for n in ("pro", "alp", "ad0", "ad1", "ad2", "ad3", "ad4", "ad5"):
                lv[n] = torch.round(lv[n] / quant) * quant
    dev_levels = [{k: v.to(DEV) for k, v in lv.items()} for lv in levels]
    head = lp.detect_decode(dev_levels, (8, 16, 32))
    want_rows, want_idx = non_max_suppression_with_index(head, conf, iou, max_det)
    plan = PostprocessPlan(dev_levels, (8, 16, 32), max_det, want_anchor=True)
    out, counts = plan.run(conf, iou)
    ks = counts.cpu().tolist()
    assert ks == [int(r.shape[0]) for r in want_rows]
    for b, k in enumerate(ks):
        assert torch.equal(plan.kept_anchor[b, :k].long(), want_idx[b]), f"kept anchors differ in image {b}"
        assert_rows_equal(out[b, :k].cpu().numpy(), want_rows[b].cpu().numpy(), f"fused[{b}]")
    # and the public wrapper
    rows = lp.detect_postprocess(dev_levels, (8, 16, 32), conf, iou, max_det)
    for b in range(B):
        assert torch.equal(rows[b], out[b, :ks[b]])


def _separated_levels(B, H, W, seed, conf):
    """Level tensors on which no decision sits within 1e-4 of a flip, so that the GPU's SFU sigmoid and the
    CPU's (they differ by <= 1e-5 relative) must give the SAME kept set: every confident anchor carries one
    logit L in all eight groups, the Ls are 0.004 apart (scores ~1e-3 apart relative) and keep 0.002 clear of
    the logit of the confidence threshold; the background sits at <= -6; boxes live on a half-cell grid."""
    g = torch.Generator().manual_seed(seed)
    names, widths = ("pro", "alp", "ad0", "ad1", "ad2", "ad3", "ad4", "ad5"), (31, 24, 37, 37, 37, 37, 37, 37)
    l_thr = float(np.log(conf / (1.0 - conf)))
    levels, rank = [], 0
    for h, w in synth.level_shapes(H, W):
        pos = torch.rand((B, 1, h, w), generator=g) < 0.15
        n = int(pos.sum())
        u = torch.arange(rank, rank + n, dtype=torch.float64)[torch.randperm(n, generator=g)]
        rank += n
        L = -1.6 + 0.004 * u
        L = torch.where((L - l_thr).abs() < 0.002, L + 0.004 * 0.5, L)          # clear of the threshold
        hot = torch.zeros((B, 1, h, w), dtype=torch.float32)
        hot[pos] = L.float()
        lv = {}
        for nme, c in zip(names, widths):
            x = -6.0 - torch.rand((B, c, h, w), generator=g) * 2.0
            j = torch.randint(c, (B, 1, h, w), generator=g)
            x = torch.where(pos.expand(B, c, h, w) & (torch.arange(c).view(1, c, 1, 1) == j), hot.expand(B, c, h, w), x)
            lv[nme] = x.contiguous()
        lv["reg"] = torch.randint(2, 11, (B, 4, h, w), generator=g).float() * 0.5
        lv["cor"] = torch.randint(-2, 11, (B, 8, h, w), generator=g).float() * 0.5
        levels.append(lv)
    return levels


@pytest.mark.parametrize("B,H,W", [(2, 256, 256), (3, 320, 192)])
def test_fused_postprocess_vs_cpu_oracle_end_to_end_exact_kept_sets(B, H, W):
    """Fused path against the CPU oracle END TO END (its own sigmoid, decode and NMS -- nothing shared with
    the device): kept anchors and their order exactly equal, box / corner / argmax columns bit-exact,
    confidence columns within the 1e-5 relative bar.  The input is built so that no threshold, order or
    IoU decision is within 1e-4 of flipping (round 1's version fell back to ">= 98 % overlap")."""
    conf, iou = 0.3, 0.4567
    levels = _separated_levels(B, H, W, seed=B * 1000 + H, conf=conf)
    head = lp_oracle.detect_decode([{k: v.numpy() for k, v in lv.items()} for lv in levels], (8, 16, 32))
    want, widx = lp_oracle.non_max_suppression(head, conf, iou, return_index=True)
    from yolo_lp_b200.head import PostprocessPlan
    plan = PostprocessPlan([{k: v.to(DEV) for k, v in lv.items()} for lv in levels], (8, 16, 32), 300, want_anchor=True)
    out, counts = plan.run(conf, iou)
    ks = counts.cpu().tolist()
    assert min(len(i) for i in widx) >= 10, "test input keeps too few detections to mean anything"
    for b, k in enumerate(ks):
        got_idx = plan.kept_anchor[b, :k].cpu().numpy()
        assert k == len(widx[b]) and np.array_equal(got_idx, widx[b]), f"[{b}] kept anchors differ from the CPU oracle"
        got = out[b, :k].cpu().numpy()
        assert np.array_equal(got[:, :12].view(np.uint32), want[b][:, :12].view(np.uint32)), f"[{b}] box / corner columns"
        np.testing.assert_allclose(got[:, 12:20], want[b][:, 12:20], rtol=1e-5, atol=0)
        assert np.array_equal(got[:, 20:], want[b][:, 20:]), f"[{b}] argmax columns"


def test_device_sigmoid_is_monotone_and_accurate():
    """max_j sigmoid(x_j) == sigmoid(max_j x_j) in the fused path rests on monotonicity: checked over
    every finite fp32 bit pattern, plus the 1e-5 relative accuracy bar on a dense sample."""
    from yolo_lp_b200 import _abi
    stream = torch.cuda.current_stream().cuda_stream
    CH = 1 << 26
    for sign in (0, 1):
        prev = None
        for lo in range(0, 0x7f800000, CH):
            hi = min(lo + CH, 0x7f800000)
            bits = torch.arange(lo, hi, dtype=torch.int64, device=DEV) + (0x80000000 if sign else 0)
            x = bits.to(torch.uint32).view(torch.float32)
            y = torch.empty_like(x)
            _abi.call("lp_debug_sigmoid_f32", x.data_ptr(), x.numel(), y.data_ptr(), stream)
            d = y[1:] - y[:-1]
            assert not bool(((d < 0) if not sign else (d > 0)).any()), f"sigmoid not monotone near bits {lo:#x}"
            if prev is not None:
                assert bool(y[0] >= prev) if not sign else bool(y[0] <= prev)
            prev = y[-1].clone()
            xs = x[::8192].double()
            ref = 1.0 / (1.0 + torch.exp(-xs))
            ok = ref > 1e-37
            rel = ((y[::8192].double() - ref).abs() / ref)[ok]
            assert rel.numel() == 0 or float(rel.max()) < 1e-5


def test_fused_pipeline_matches_serial_fused_path():
    from yolo_lp_b200.head import PostprocessPlan, PostprocessPipeline
    levels_a = synth.synth_levels(4, 320, 320, torch.device(DEV), seed=1)
    levels_b = synth.synth_levels(4, 320, 320, torch.device(DEV), seed=2)
    ref = []
    for lv in (levels_a, levels_b):
        plan = PostprocessPlan(lv, (8, 16, 32), 300)
        out, counts = plan.run(0.25, 0.45)
        ref.append((out.clone(), counts.clone()))
    pipe = PostprocessPipeline([PostprocessPlan(levels_a, (8, 16, 32), 300), PostprocessPlan(levels_b, (8, 16, 32), 300)])
    pipe.start()
    got = []
    for i in range(6):
        slot, out, counts = pipe.submit(0.25, 0.45)
        pipe.done[slot].synchronize()
        got.append((slot, out.clone(), counts.clone()))
    pipe.finish()
    torch.cuda.synchronize()
    for slot, out, counts in got:
        assert torch.equal(counts, ref[slot][1])
        for b, k in enumerate(counts.cpu().tolist()):
            assert torch.equal(out[b, :k], ref[slot][0][b, :k])
    assert int(ref[0][1].sum()) > 0


# ------------------------------------------------------------------ ordering fall-backs and fuzz
@pytest.mark.parametrize("A", [16000, 20000])
def test_degenerate_scores_take_the_full_sort_fallbacks(A):
    """All scores equal (the true random-init model): the score histogram has one
    bin, segmentation is refused and the full bitonic sort runs -- in shared memory for 16000
    candidates, in the global workspace for 20000.  Order is then purely by anchor index."""
    pred = synth.synth_head(1, A, 1280, 48, 0, seed=12)
    pred[0, :, 13:] = 0.5
    g = torch.Generator().manual_seed(4)
    pred[0, :, 0:2] = torch.rand((A, 2), generator=g) * 1200.0
    _check_against_oracle(pred, 0.25, 0.45, 300, f"degenerate A={A}")


@pytest.mark.parametrize("A,levels", [(600, 1), (1000, 1), (3000, 1), (9000, 1), (700, 3), (5000, 7), (8400, 40)])
def test_tied_scores_take_the_network_sorts_inside_a_segment(A, levels):
    """`levels` distinct scores only: the histogram bins hold hundreds to thousands of keys, far above
    the counting sort's 128-key bin limit, so the segments are ordered by the bitonic networks
    (block_sort up to 1024 keys, the shared-memory bitonic sort beyond) and ties fall back on the
    anchor index.  One level = a single bin holding everything."""
    pred = synth.synth_head(1, A, 640, 24, 0, seed=13)
    g = torch.Generator().manual_seed(5)
    lv = torch.randint(levels, (A, 1), generator=g).float() / max(levels, 2) * 0.4 + 0.5
    pred[0, :, 13:] = lv
    pred[0, :, 0:2] = torch.rand((A, 2), generator=g) * 600.0
    _check_against_oracle(pred, 0.25, 0.45, 300, f"tied scores A={A} levels={levels}")


@pytest.mark.parametrize("iou", IOU_BAND_THRESHOLDS)
def test_iou_within_ulps_of_the_threshold(iou):
    """K2 decides IoU tests without the division when the quotient is clearly off the threshold and
    takes the exact IEEE quotient inside a +-1e-6 band round it.  Here every image holds one pair of
    boxes whose IoU sits within a few ulps of the threshold (both sides, many magnitudes): the kept
    sets must still be the reference's bit for bit -- checked against the oracle AND against the kept
    counts the reference itself (torchvision CPU nms) produced for these pairs
    (tests/golden/iou_band.npz).  Includes thresholds whose float is the degenerate input of that
    shortcut (0, 1, a near-denormal one)."""
    pred = iou_band_pairs(iou)
    got = _check_against_oracle(pred, 0.25, iou, 300, f"iou threshold band thr={iou}")
    kept = np.array([len(g) for g in got])
    want = golden("iou_band")["kept_%d" % IOU_BAND_THRESHOLDS.index(iou)]
    assert np.array_equal(kept, want), "kept counts differ from the reference run"
    if 1e-30 < iou < 1.0:
        assert 0.1 < (kept == 1).mean() < 0.9, "the pairs must straddle the threshold"


# ------------------------------------------------------------------ fp16 head tensors
def _half_case(B, A, n_pos, seed, quant=None):
    pred = synth.synth_head(B, A, 640, 24, n_pos, seed=seed, quant=quant)
    return pred.half()


@pytest.mark.parametrize("B,A,n_pos,conf,iou,max_det", [
    (4, 8400, 300, 0.25, 0.45, 300),
    (2, 8400, 300, 0.001, 0.65, 300),     # every anchor a candidate
    (3, 315, 40, 0.05, 0.45, 300),        # odd A: rows only 4-byte aligned, ragged 64-row tiles
    (5, 33, 10, 0.05, 0.5, 7),            # A < 64: several images per tile
    (1, 33600, 4096, 0.25, 0.45, 300),
    (2, 2100, 2100, 0.0, 0.3, 1500),      # more kept rows than one gather batch
])
def test_half_head_tensor_equals_fp32_path_on_the_upcast_tensor(B, A, n_pos, conf, iou, max_det):
    """fp16 storage, exact upcast on load, fp32 arithmetic: kept anchors and all 28 columns must be
    bit for bit those of the oracle (and of the fp32 kernels) on ``pred.float()`` -- through the serial
    entry, the two-stream pipeline and the host-buffer path."""
    from yolo_lp_b200.nms import NmsPipeline, non_max_suppression_with_index
    ph = _half_case(B, A, n_pos, seed=70 + B)
    up = ph.float()
    want, widx = lp_oracle.non_max_suppression(up.numpy(), conf, iou, max_det=max_det, return_index=True)
    rows, idx = non_max_suppression_with_index(ph.to(DEV), conf, iou, max_det)
    rows32, idx32 = non_max_suppression_with_index(up.to(DEV), conf, iou, max_det)
    for b in range(B):
        assert np.array_equal(idx[b].cpu().numpy(), widx[b]), f"half[{b}]: kept anchors differ from the oracle"
        assert_rows_equal(rows[b].cpu().numpy(), want[b], f"half[{b}]")
        assert torch.equal(rows[b], rows32[b]) and torch.equal(idx[b], idx32[b])
    # pipelined entry (lp_nms_pipelined_f16), three submissions so the re-armed workspace is exercised
    pipe = NmsPipeline(B, A, max_det, torch.device(DEV))
    dev_h = ph.to(DEV)
    pipe.start()
    for _ in range(3):
        slot, out, counts = pipe.submit(dev_h, conf, iou)
    pipe.finish()
    torch.cuda.synchronize()
    for b, k in enumerate(counts.cpu().tolist()):
        assert_rows_equal(out[b, :k].cpu().numpy(), want[b], f"half pipelined[{b}]")
    # the reference-signature entry returns rows in the prediction's dtype (the reference's torch.cat,
    # nms.py:94-96): the fp32 rows rounded once to half -- device tensor and host-buffer path (the
    # halves travel as halves)
    host = lp.non_max_suppression(ph, conf, iou, max_det=max_det)
    devr = lp.non_max_suppression(dev_h, conf, iou, max_det=max_det)
    for b in range(B):
        w16 = torch.from_numpy(want[b]).half()
        assert host[b].dtype == torch.float16 and devr[b].dtype == torch.float16 and devr[b].is_cuda
        assert torch.equal(host[b].view(torch.int16), w16.view(torch.int16)), f"half host path[{b}]"
        assert torch.equal(devr[b].cpu().view(torch.int16), w16.view(torch.int16)), f"half device path[{b}]"


@pytest.mark.parametrize("B,H,W,conf", [(32, 640, 640, 0.25), (9, 640, 640, 0.001), (4, 1280, 1280, 0.25),
                                         (5, 416, 640, 0.05), (3, 320, 320, 0.1)])
def test_fused_path_on_half_level_tensors_equals_f32_path_on_the_upcast_tensors(B, H, W, conf):
    """fp16 level tensors (model.half()) through lp_detect_postprocess_f16 / lp_detect_pipelined_f16:
    exact upcast on load, so the detections must be bit for bit those of the f32 entries on the upcast
    tensors.  320x320 has a 10x10 level (h*w % 8 != 0): the shim upcasts and takes the f32 entry."""
    from yolo_lp_b200.head import PostprocessPlan, PostprocessPipeline
    levels = synth.synth_levels(B, H, W, DEV, seed=B + W)
    half = [{k: v.half() for k, v in lv.items()} for lv in levels]
    up = [{k: v.float() for k, v in lv.items()} for lv in half]
    ref = PostprocessPlan(up, (8, 16, 32), 300)
    ref_out, ref_counts = ref.run(conf, 0.45)
    plan = PostprocessPlan(half, (8, 16, 32), 300)
    assert plan.half == all((lv["reg"].shape[2] * lv["reg"].shape[3]) % 8 == 0 for lv in levels)
    plan.workspace.fill_(0xFF)
    out, counts = plan.run(conf, 0.45)
    torch.cuda.synchronize()
    assert torch.equal(counts, ref_counts) and int(counts.sum()) > 0
    for b, k in enumerate(counts.cpu().tolist()):
        assert torch.equal(out[b, :k], ref_out[b, :k]), f"image {b}"
    pipe = PostprocessPipeline([plan, PostprocessPlan(half, (8, 16, 32), 300)])
    pipe.start()
    for _ in range(5):
        slot, out, counts = pipe.submit(conf, 0.45)
    pipe.finish()
    torch.cuda.synchronize()
    assert torch.equal(counts, ref_counts)
    for b, k in enumerate(counts.cpu().tolist()):
        assert torch.equal(out[b, :k], ref_out[b, :k]), f"pipelined image {b}"


@pytest.mark.parametrize("half", [False, True])
def test_fused_path_with_foreign_kernels_sharing_the_sms(half):
    """Small foreign kernels on another stream move onto the SMs beside the persistent KF CTAs and
    perturb the relative speed of its warps.  That exposed an mbarrier phase-parity aliasing in an
    earlier producer / finisher assignment (a waiter two phases ahead passes try_wait.parity): wrong
    slots were overwritten and K2 then read garbage keys -- a launch failure after a few hundred runs.
    Every run's result is compared on the device with the first one."""
    from yolo_lp_b200.head import PostprocessPlan
    levels = synth.synth_levels(32, 640, 640, DEV, seed=1)
    if half:
        levels = [{k: v.half() for k, v in lv.items()} for lv in levels]
    plan = PostprocessPlan(levels, (8, 16, 32), 300)
    out, counts = plan.run(0.25, 0.45)
    ref_out, ref_counts = out.clone(), counts.clone()
    side = torch.cuda.Stream(DEV)
    rc = torch.randint(0, 300, (32,), device=DEV, dtype=torch.int32)
    live = torch.arange(300, device=DEV)[None, :, None] < ref_counts[:, None, None]
    bad = torch.zeros((), dtype=torch.int64, device=DEV)
    torch.cuda.synchronize()
    for _ in range(1500):
        out, counts = plan.run(0.25, 0.45)
        bad.add_(((out != ref_out) & live).any().long() + (counts != ref_counts).any().long())
        with torch.cuda.stream(side):
            _ = torch.arange(300, device=DEV)[None, :, None] < rc[:, None, None]
    torch.cuda.synchronize()
    assert int(bad) == 0


def test_decode_with_foreign_kernels_sharing_the_sms():
    """Same perturbation for the warp-specialised decode kernel (rings, named barriers, release warp)."""
    levels = synth.synth_levels(16, 640, 640, DEV, seed=2)
    plan = lp.DecodePlan(levels, (8, 16, 32))
    ref = plan.run().clone()
    side = torch.cuda.Stream(DEV)
    rc = torch.randint(0, 300, (32,), device=DEV, dtype=torch.int32)
    bad = torch.zeros((), dtype=torch.int64, device=DEV)
    torch.cuda.synchronize()
    for _ in range(600):
        out = plan.run()
        bad.add_((out.view(torch.int32) != ref.view(torch.int32)).any().long())
        with torch.cuda.stream(side):
            _ = torch.arange(300, device=DEV)[None, :, None] < rc[:, None, None]
    torch.cuda.synchronize()
    assert int(bad) == 0


def test_heavy_suppression_walks_many_segments():
    """Few tight clusters, every anchor a candidate: far fewer than max_det boxes survive, so the
    greedy walk has to consume every score segment (and every window) of the 8400 candidates."""
    A = 8400
    pred = synth.synth_head(2, A, 640, 24, 300, seed=13)
    g = torch.Generator().manual_seed(5)
    centres = torch.rand((10, 2), generator=g) * 600.0
    which = torch.randint(10, (2, A), generator=g)
    pred[:, :, 0:2] = centres[which] + torch.rand((2, A, 2), generator=g) * 4.0
    pred[:, :, 2:4] = 60.0
    got = _check_against_oracle(pred, 0.0, 0.45, 300, "heavy suppression")
    assert all(r.shape[0] < 100 for r in got)


@pytest.mark.parametrize("seed", range(24))
def test_nms_fuzz_against_oracle(seed):
    rng = np.random.default_rng(1000 + seed)
    B = int(rng.integers(1, 4))
    A = int(rng.choice([7, 64, 100, 333, 777, 1500, 2600, 4100]))
    n_pos = int(rng.integers(0, max(2, A // 3)))
    conf = float(rng.choice([0.0, 0.01, 0.05, 0.25, 0.6]))
    iou = float(rng.choice([0.0, 0.2, 0.45, 0.65, 1.0]))
    max_det = int(rng.choice([1, 7, 64, 300, 1000]))
    quant = rng.choice([0, 0, 4, 16])
    plates = int(rng.choice([1, 3, 12, 40]))
    pred = synth.synth_head(B, A, 640, plates, n_pos, seed=2000 + seed, quant=int(quant) or None)
    if seed % 3 == 0:   # objectness != 1 exercises the cls * obj product (nms.py:76)
        g = torch.Generator().manual_seed(seed)
        pred[..., 4] = torch.rand(pred.shape[:2], generator=g)
    _check_against_oracle(pred, conf, iou, max_det, f"fuzz{seed} B{B} A{A} conf{conf} iou{iou} md{max_det} q{quant}")


# ------------------------------------------------------------------ caller-side records
def test_txt_records_kernel_matches_reference_records():
    from yolo_lp_b200.inferer import txt_records, txt_lines
    g = golden("txt_records")
    n = int(g["n"])
    k = g["det0"].shape[0]
    det = torch.zeros((n, k + 5, 28))
    for i in range(n):
        det[i, :k] = torch.from_numpy(g[f"det{i}"])
    counts = torch.full((n,), k, dtype=torch.int32)
    rec = txt_records(det.to(DEV), counts.to(DEV), [g[f"src{i}"].tolist() for i in range(n)]).cpu()
    for i in range(n):
        want = g[f"rec{i}"]
        got = rec[i, :k].numpy()
        assert np.array_equal(got[:, :20].view(np.uint32), want[:, :20].view(np.uint32)), i
        np.testing.assert_allclose(got[:, 20], want[:, 20], rtol=1e-6)
        assert txt_lines(rec[i, :k]) == "".join(l + "\n" for l in g[f"lines{i}"].tolist())


def test_inferer_flow_nms_rescale_records_end_to_end():
    """Inferer.infer's post-model flow (inferer.py:82,100,103-120) in three launches: NMS with the
    fused rescale + round epilogue, then the record kernel; checked against the oracle chain."""
    from yolo_lp_b200.inferer import rescale_table, txt_records, txt_lines
    pred = synth.synth_head(2, 8400, 640, 24, 200, seed=71)
    ori, src = [(640, 640), (640, 640)], [(1160, 720, 3), (480, 854, 3)]
    plan = NmsPlan(2, 8400, 1000, torch.device(DEV))
    out, counts = plan.run(pred.to(DEV), 0.4, 0.45, rescale=rescale_table(ori, src, DEV), do_round=True)
    rec = txt_records(out, counts, src).cpu()
    want = lp_oracle.non_max_suppression(pred.numpy(), 0.4, 0.45, max_det=1000)
    for b, k in enumerate(counts.cpu().tolist()):
        rows = want[b].copy()
        assert k == rows.shape[0] > 0
        rows[:, :12] = lp_oracle.rescale(ori[b], rows[:, :12], src[b], do_round=True)
        wrec = lp_oracle.txt_records(rows, src[b])
        assert np.array_equal(rec[b, :k, :20].numpy().view(np.uint32), wrec[:, :20].view(np.uint32))
        assert txt_lines(rec[b, :k]) == "".join(lp_oracle.txt_line(r) + "\n" for r in wrec)


# ------------------------------------------------------------------ LP eval metric
def test_eval_metric_matches_reference_and_oracle():
    from yolo_lp_b200.evaler import lp_eval, eval_counts
    g = golden("eval_metric")
    n, split = int(g["n"]), int(g["split"])
    preds = [torch.from_numpy(g[f"pred{i}"]).to(DEV) for i in range(n)]
    targets = [torch.from_numpy(g[f"tgt{i}"]).to(DEV) for i in range(n)]
    res = lp_eval([preds[:split], preds[split:]], [targets[:split], targets[split:]])   # nested like Evaler.predict
    assert np.array_equal(np.array(res[:5]), g["scalars"])
    assert np.array_equal(np.array(res[5]), g["mAP_list"]) and np.array_equal(np.array(res[6]), g["recall_list"])
    counters, _ = eval_counts(preds, targets)
    per = []
    for i in range(n):
        ti, _m, ic, il = lp_oracle.eval_match(g[f"pred{i}"], g[f"tgt{i}"])
        per.append((g[f"pred{i}"].shape[0], ti, ic, il))
    want = lp_oracle.eval_accumulate(per)
    assert counters[2:12].tolist() == want["pred_cnts"] and counters[32:42].tolist() == want["right"]
    assert counters[12:22].tolist() == want["cor_right"] and counters[22:32].tolist() == want["cls_right"]


# ------------------------------------------------------------------ repeatability / multi-device driver
@pytest.mark.parametrize("cid,B", [(4, 8), (5, 4), (2, 16)])
def test_repeated_runs_are_bitwise_identical(cid, B):
    """Race hunt (compute-sanitizer is closed on this pool): the dense configs exercise the
    histogram segments, several windows and the replica / chunk barriers of K2; 40 back-to-back
    runs, alternating between two workspaces, must reproduce the first result bit for bit."""
    cfg = synth.CONFIGS[cid]
    pred = synth.synth_head(B, cfg["A"], cfg["img"], cfg["n_plates"], cfg["n_pos"], cfg["seed"]).to(DEV)
    plans = [NmsPlan(B, cfg["A"], cfg["max_det"], torch.device(DEV), want_anchor=True) for _ in range(2)]
    out0, cnt0 = plans[0].run(pred, cfg["conf"], cfg["iou"])
    out0, cnt0, idx0 = out0.clone(), cnt0.clone(), plans[0].kept_anchor.clone()
    ks = cnt0.cpu().tolist()
    for it in range(40):
        plan = plans[it & 1]
        out, cnt = plan.run(pred, cfg["conf"], cfg["iou"])
        assert torch.equal(cnt, cnt0), f"run {it}: counts differ"
        for b, k in enumerate(ks):
            assert torch.equal(plan.kept_anchor[b, :k], idx0[b, :k]), f"run {it}: kept anchors of image {b} differ"
            assert torch.equal(out[b, :k], out0[b, :k]), f"run {it}: rows of image {b} differ"


def test_sharded_driver_concatenates_in_image_order():
    from yolo_lp_b200.shard import ShardedNms
    n_dev = torch.cuda.device_count()
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]      # one GPU: two shards on the same device
    B = 7
    pred = synth.synth_head(B, 2100, 320, 8, 100, seed=81)
    want = lp_oracle.non_max_suppression(pred.numpy(), 0.2, 0.45)
    drv = ShardedNms(B, 2100, 300, devices=devices)
    shards = [pred[lo:hi].to(f"cuda:{d}") for d, (lo, hi) in zip(devices, drv.ranges)]
    got = drv.run(shards, 0.2, 0.45)
    assert len(got) == B
    for b in range(B):
        assert_rows_equal(got[b].numpy(), want[b], f"shard[{b}]")


def test_prepare_targets_matches_reference():
    from yolo_lp_b200.evaler import prepare_targets
    g = golden("eval_targets")
    out = prepare_targets(torch.from_numpy(g["targets"]).to(DEV), int(g["w"]), int(g["h"]), int(g["bs"]))
    assert len(out) == int(g["bs"])
    for i, o in enumerate(out):
        assert np.array_equal(o.cpu().numpy().view(np.uint32), g[f"out{i}"].view(np.uint32)), i


@pytest.mark.parametrize("n,shift", [(400, 4.0), (1500, 4.0), (300, 2.5)])
def test_long_suppression_chain_takes_the_sequential_pass(n, shift):
    """A row of equal boxes in which each overlaps its successors a little less: with shift 4 (IoU 0.43 at
    distance one... 10-wide boxes) the greedy result alternates down the whole list -- every decision
    depends on the previous one, the worst case for K2's parallel fixed-point rounds (they do not settle
    within FIXPOINT_ROUNDS and the sequential pass must take over); shift 2.5 makes every box kill its
    next two.  Kept anchors and rows bit-exact vs the oracle; several windows for n = 1500."""
    x = torch.zeros((1, n, 290))
    k = torch.arange(n, dtype=torch.float32)
    x[0, :, 0] = 100.0 + k * shift          # cx
    x[0, :, 1] = 50.0
    x[0, :, 2] = 10.0                        # w: IoU(d) = (10 - d) / (10 + d) for centre distance d
    x[0, :, 3] = 20.0
    x[0, :, 4] = 1.0
    x[0, :, 13:] = (0.9 - 0.0005 * k)[:, None]     # strictly decreasing scores: list order = anchor order
    want, widx = lp_oracle.non_max_suppression(x.numpy(), 0.1, 0.4, max_det=n, return_index=True)
    rows, idx = non_max_suppression_with_index(x.to(DEV), 0.1, 0.4, n)
    assert len(widx[0]) > n // 4
    assert np.array_equal(idx[0].cpu().numpy(), widx[0])
    assert_rows_equal(rows[0].cpu().numpy(), want[0], "chain")


@pytest.mark.parametrize("half", [False, True])
def test_fused_pipeline_to_host_entry(half):
    """lp_detect_pipelined_to_host_f32 / _f16: KF + K2 + the D2H copy of the step's detections in one native
    call; what lands in the pinned host buffers equals the serial entry's result, step after step (three
    slots' worth, so every workspace is re-armed)."""
    from yolo_lp_b200.head import PostprocessPlan, PostprocessPipeline
    B = 6
    levels = synth.synth_levels(B, 640, 640, DEV, seed=21)
    if half:
        levels = [{k: v.half() for k, v in lv.items()} for lv in levels]
    ref = PostprocessPlan(levels, (8, 16, 32), 300)
    ref_out, ref_counts = (t.cpu() for t in ref.run(0.25, 0.45))
    plans = [PostprocessPlan(levels, (8, 16, 32), 300) for _ in range(2)]
    assert plans[0].half == half
    pipe = PostprocessPipeline(plans)
    copy_stream = torch.cuda.Stream(DEV)
    outs = [torch.full((B, 300, 28), float("nan")).pin_memory() for _ in plans]
    cnts = [torch.full((B,), -1, dtype=torch.int32).pin_memory() for _ in plans]
    evs = [torch.cuda.Event() for _ in plans]
    pipe.start()
    for step in range(6):
        slot = pipe.n % 2
        if step >= 2:
            evs[slot].synchronize()          # the host takes the slot's previous batch before it is reused
        outs[slot].fill_(float("nan"))
        got = pipe.submit_to_host(0.25, 0.45, outs[slot], cnts[slot], copy_stream, evs[slot])
        assert got == slot
    for slot in range(2):
        evs[slot].synchronize()
        assert torch.equal(cnts[slot], ref_counts) and int(ref_counts.sum()) > 0
        for b, k in enumerate(ref_counts.tolist()):
            assert torch.equal(outs[slot][b, :k], ref_out[b, :k]), f"slot {slot} image {b}"
    pipe.finish()
    with pytest.raises(ValueError):
        pipe.submit_to_host(0.25, 0.45, torch.empty((B, 300, 28)), cnts[0], copy_stream, evs[0])   # not pinned
