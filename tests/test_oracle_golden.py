"""Pin the oracle (numpy restatement and torch port) to outputs of the REFERENCE
itself (tests/golden/*.npz, made by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import lp_oracle, torch_port
from _util import (golden, golden_names, split_rows, seeded_inputs, assert_rows_equal, iou_band_pairs,
                   IOU_BAND_THRESHOLDS)

SEEDED = [n for n in golden_names("nms_cfg") + golden_names("nms_eval")]
EDGES = golden_names("nms_edge_")
DECODE = [n for n in golden_names("decode_") if "half" not in n]


def _knobs(g):
    return float(g["conf"]), float(g["iou"]), int(g["max_det"])


@pytest.mark.parametrize("name", SEEDED)
def test_numpy_oracle_seeded(name):
    g = golden(name)
    conf, iou, max_det = _knobs(g)
    pred = seeded_inputs(g).numpy()
    want = split_rows(g["counts"], g["rows"])
    got = lp_oracle.non_max_suppression(pred, conf, iou, max_det=max_det)
    for b, (a, w) in enumerate(zip(got, want)):
        assert_rows_equal(a, w, f"{name}[{b}]")


@pytest.mark.parametrize("name", EDGES)
def test_numpy_oracle_edges(name):
    g = golden(name)
    conf, iou, max_det = _knobs(g)
    pred = g["pred"]
    pred = pred[None] if pred.ndim == 2 else pred
    want = split_rows(g["counts"], g["rows"])
    got, idx = lp_oracle.non_max_suppression(pred, conf, iou, max_det=max_det, return_index=True)
    for b, (a, w) in enumerate(zip(got, want)):
        assert_rows_equal(a, w, f"{name}[{b}]")
        assert len(idx[b]) == len(a)


@pytest.mark.parametrize("name", SEEDED[:3] + EDGES)
def test_torch_port(name):
    g = golden(name)
    conf, iou, max_det = _knobs(g)
    pred = torch.from_numpy(g["pred"]) if "pred" in g else seeded_inputs(g)
    pred = pred[None] if pred.ndim == 2 else pred
    want = split_rows(g["counts"], g["rows"])
    got = torch_port.non_max_suppression(pred.clone(), conf, iou, max_det=max_det)
    for b, (a, w) in enumerate(zip(got, want)):
        assert_rows_equal(a.numpy(), w, f"{name}[{b}]")


def test_more_than_max_nms_candidates_golden():
    """nms.py:115-116 at the default max_nms = 30000: numpy oracle and torch port against the reference's
    output on 33600 passing anchors with distinct scores (tests/_util.maxnms_input)."""
    import hashlib
    from _util import maxnms_input
    from yolo_lp_b200 import synth
    g = golden("nms_maxnms_33600")
    x = maxnms_input()
    assert synth.sha256_of(x) == str(g["sha256"]), "numpy generator is not bit-reproducible on this host"
    for tag in ("a", "b"):
        iou, max_det = float(g[f"iou_{tag}"]), int(g[f"max_det_{tag}"])
        got = lp_oracle.non_max_suppression(x[None].numpy(), 0.0, iou, max_det=max_det)
        assert_rows_equal(got[0], g[f"rows_{tag}"], f"numpy oracle {tag}")
        tp = torch_port.non_max_suppression(x[None].clone(), 0.0, iou, max_det=max_det)
        assert_rows_equal(tp[0].numpy(), g[f"rows_{tag}"], f"torch port {tag}")
    rows = torch_port.non_max_suppression(x[None].clone(), 0.0, 0.999, max_det=33600)[0].numpy()
    assert rows.shape[0] == int(g["count_c"]) == 30000 and np.array_equal(rows[-16:], g["tail_c"])
    assert hashlib.sha256(rows.tobytes()).hexdigest() == str(g["sha256_c"])


def _levels(g):
    return [{k: g[f"l{l}_{k}"] for k in ("pro", "alp", "ad0", "ad1", "ad2", "ad3", "ad4", "ad5", "reg", "cor")}
            for l in range(3)]


@pytest.mark.parametrize("name", DECODE)
def test_decode_oracle(name):
    g = golden(name)
    want = g["out"]
    got = lp_oracle.detect_decode(_levels(g), (8, 16, 32))
    assert got.shape == want.shape
    # geometry columns and obj are exactly restatable; sigmoid is held to 1e-5 rel
    assert np.array_equal(got[..., :13].view(np.uint32), want[..., :13].view(np.uint32))
    np.testing.assert_allclose(got[..., 13:], want[..., 13:], rtol=1e-5, atol=0)
    tp = torch_port.detect_decode([{k: torch.from_numpy(v) for k, v in lv.items()} for lv in _levels(g)], (8, 16, 32))
    np.testing.assert_allclose(tp.numpy(), want, rtol=1e-6, atol=0)
    # NMS on the reference's own decoded tensor
    rows = lp_oracle.non_max_suppression(want, *_knobs(g)[:2], max_det=_knobs(g)[2])
    for b, (a, w) in enumerate(zip(rows, split_rows(g["counts"], g["rows"]))):
        assert_rows_equal(a, w, f"{name}[{b}]")


def test_decode_half_mode_oracle():
    """model.half() forward of the reference's Detect (run in half on the CPU for the golden): the numpy
    oracle with half_scores reproduces its fp32 head tensor exactly -- geometry bit for bit (computed in
    fp32 from the exactly upcast half conv outputs, because the reference's anchors are fp32) and every
    class score (sigmoid rounded to half)."""
    g = golden("decode_half_96x160")
    levels = _levels(g)
    assert all(v.dtype == np.float16 for lv in levels for v in lv.values())
    got = lp_oracle.detect_decode(levels, (8, 16, 32), half_scores=True)
    want = g["out"]
    assert want.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got[..., :13].view(np.uint32), want[..., :13].view(np.uint32))
    assert (got[..., 13:] == want[..., 13:]).mean() >= 0.999
    np.testing.assert_allclose(got[..., 13:], want[..., 13:], rtol=2.0 ** -10, atol=0)
    rows = lp_oracle.non_max_suppression(want, *_knobs(g)[:2], max_det=_knobs(g)[2])
    for b, (a, w) in enumerate(zip(rows, split_rows(g["counts"], g["rows"]))):
        assert_rows_equal(a, w, f"half[{b}]")


def test_geometry_oracle():
    g = golden("geometry")
    ap, st = lp_oracle.generate_anchors_eval([tuple(x) for x in g["hw"].tolist()], (8, 16, 32))
    assert np.array_equal(ap, g["anchor_points"]) and np.array_equal(st, g["stride_tensor"])
    assert np.array_equal(lp_oracle.dist2bbox(g["dist"], ap, "xyxy"), g["bbox_xyxy"])
    assert np.array_equal(lp_oracle.dist2bbox(g["dist"], ap, "xywh"), g["bbox_xywh"])
    assert np.array_equal(lp_oracle.dist2cor(g["cdist"], ap), g["corners"])


def test_rescale_oracle():
    g = golden("rescale")
    for n in range(int(g["n"])):
        hi, wi, h0, w0 = g[f"shape{n}"].tolist()
        got = lp_oracle.rescale((hi, wi), g[f"in{n}"], (h0, w0, 3))
        assert np.array_equal(got.view(np.uint32), g[f"out{n}"].view(np.uint32)), n
        got = lp_oracle.rescale((hi, wi), g[f"in{n}"], (h0, w0, 3), do_round=True)
        assert np.array_equal(got, g[f"round{n}"]), n
        tp = torch_port.rescale((hi, wi), torch.from_numpy(g[f"in{n}"].copy()), (h0, w0, 3))
        assert np.array_equal(tp.numpy().view(np.uint32), g[f"out{n}"].view(np.uint32)), n


def test_txt_records_oracle_and_native_formatter():
    """--save-txt records (inferer.py:92-93,103-120): numeric record bit-exact (conf: torch.mean vs a
    sequential sum, 1 ulp), text lines identical -- both for the oracle's '%g' and for the
    library's native host formatter (lp_txt_lines_host, no GPU involved)."""
    from yolo_lp_b200 import build
    from yolo_lp_b200.inferer import txt_lines
    build.build()
    g = golden("txt_records")
    for n in range(int(g["n"])):
        rec = lp_oracle.txt_records(g[f"det{n}"], g[f"src{n}"].tolist())
        want = g[f"rec{n}"]
        assert np.array_equal(rec[:, :20].view(np.uint32), want[:, :20].view(np.uint32))
        np.testing.assert_allclose(rec[:, 20], want[:, 20], rtol=1e-6)
        lines = g[f"lines{n}"].tolist()
        assert [lp_oracle.txt_line(r) for r in rec] == lines
        assert txt_lines(torch.from_numpy(want)) == "".join(l + "\n" for l in lines)
    assert txt_lines(torch.zeros((0, 21))) == ""


def _eval_case():
    g = golden("eval_metric")
    n = int(g["n"])
    return g, [g[f"pred{i}"] for i in range(n)], [g[f"tgt{i}"] for i in range(n)]


def test_eval_metric_oracle_matches_reference():
    """Evaler.eval (evaler.py:153-283) run by the reference on 48 images (empty predictions, empty
    targets, wrong characters, loose corners, an IoU of exactly 1.0 -> stale bin index)."""
    g, preds, targets = _eval_case()
    per = []
    for p, t in zip(preds, targets):
        ti, _m, ic, il = lp_oracle.eval_match(p, t)
        per.append((p.shape[0], ti, ic, il))
    res = lp_oracle.eval_summary(lp_oracle.eval_accumulate(per))
    assert np.array_equal(np.array(res[:5]), g["scalars"])
    assert np.array_equal(np.array(res[5], float), g["mAP_list"])
    assert np.array_equal(np.array(res[6]), g["recall_list"])


def test_eval_accumulate_host_matches_oracle_counters():
    """The native host accumulator (lp_eval_accumulate_host, no GPU) on oracle matches."""
    import ctypes
    from yolo_lp_b200 import _abi, build
    build.build()
    g, preds, targets = _eval_case()
    per, rows, timg = [], [], []
    for b, (p, t) in enumerate(zip(preds, targets)):
        ti, m, ic, il = lp_oracle.eval_match(p, t)
        per.append((p.shape[0], ti, ic, il))
        for k in range(len(ti)):
            rows.append([ti[k], m[k], float(ic[k]), float(il[k])])
            timg.append(b)
    want = lp_oracle.eval_accumulate(per)
    match = np.array(rows, np.float32).reshape(-1, 4)
    timg = np.array(timg, np.int32)
    counts = np.array([p.shape[0] for p in preds], np.int32)
    counters = np.zeros(42, np.int64)
    summary = np.zeros(25, np.float64)
    _abi.call("lp_eval_accumulate_host", match.ctypes.data, timg.ctypes.data, counts.ctypes.data, len(preds), len(timg),
              counters.ctypes.data, summary.ctypes.data)
    assert counters[0] == want["true_cnt"] and counters[1] == want["pred_cnt"]
    assert counters[2:12].tolist() == want["pred_cnts"] and counters[12:22].tolist() == want["cor_right"]
    assert counters[22:32].tolist() == want["cls_right"] and counters[32:42].tolist() == want["right"]
    assert np.array_equal(summary[:5], g["scalars"])
    assert np.array_equal(summary[5:15], g["mAP_list"]) and np.array_equal(summary[15:25], g["recall_list"])


def test_prepare_targets_oracle():
    g = golden("eval_targets")
    out = lp_oracle.prepare_targets(g["targets"], int(g["w"]), int(g["h"]), int(g["bs"]))
    for i, o in enumerate(out):
        assert np.array_equal(o.view(np.uint32), g[f"out{i}"].view(np.uint32)), i


@pytest.mark.parametrize("i", range(len(IOU_BAND_THRESHOLDS)))
def test_oracle_iou_within_ulps_of_the_threshold_matches_reference(i):
    """Pairs of boxes whose IoU sits within a few ulps of the threshold: the oracle's float-IoU vs
    double-threshold compare must reproduce the kept counts of the reference run (torchvision CPU)."""
    iou = IOU_BAND_THRESHOLDS[i]
    pred = iou_band_pairs(iou, 6000)
    out = lp_oracle.non_max_suppression(pred.numpy(), 0.25, iou)
    kept = np.array([len(o) for o in out], np.int8)
    assert np.array_equal(kept, golden("iou_band")["kept_%d" % i])
