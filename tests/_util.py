"""Shared helpers for the parity tests: golden loading and comparisons."""
import glob
import os

import numpy as np
import torch

from yolo_lp_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def split_rows(counts, flat):
    out, o = [], 0
    for c in counts.tolist():
        out.append(flat[o:o + c])
        o += c
    return out


def seeded_inputs(g):
    """Regenerate the (unstored) inputs of a seeded golden and check their SHA-256 pin."""
    cfg = synth.CONFIGS[int(g["config_id"])]
    quant = int(g["quant"]) or None
    xs = []
    for i, sha in zip(g["indices"].tolist(), g["sha256"].tolist()):
        x = synth.synth_image(cfg["A"], cfg["img"], cfg["n_plates"], cfg["n_pos"], cfg["seed"], i, quant=quant)
        assert synth.sha256_of(x) == sha, "synthetic generator is not bit-reproducible on this host"
        xs.append(x)
    return torch.stack(xs)


def assert_rows_equal(got, want, what=""):
    """Bit-exact comparison of [k,28] detection rows (counts, order, every value)."""
    got = np.asarray(got, np.float32)
    want = np.asarray(want, np.float32)
    assert got.shape == want.shape, f"{what}: kept count {got.shape[0]} != {want.shape[0]}"
    if not np.array_equal(got.view(np.uint32), want.view(np.uint32)):
        bad = np.argwhere(got.view(np.uint32) != want.view(np.uint32))
        r, c = bad[0]
        raise AssertionError(f"{what}: {len(bad)} differing values, first at row {r} col {c}: {got[r, c]!r} vs {want[r, c]!r}")


IOU_BAND_THRESHOLDS = (0.45, 0.65, 0.5, 0.3, 1e-3, 0.999, 1.0, 0.0, 1e-36)


def iou_band_pairs(iou, n=6000):
    """``[n, 2, 290]`` head rows, one pair of equal-sized boxes per image whose IoU sits within a few
    ulps of ``iou`` (both sides, extents over four decades): the first box has the higher score, the
    second is suppressed iff ``(double)ovr > iou``.  numpy only, so the bits are the same everywhere."""
    rng = np.random.default_rng(int(iou * 1e6) % 9973 + 3)
    t = np.float32(iou)
    if float(t) > iou:
        t = np.nextafter(t, np.float32(-np.inf))
    w = (10.0 ** rng.uniform(-1.0, 3.0, n)).astype(np.float32)
    h = (10.0 ** rng.uniform(-1.0, 3.0, n)).astype(np.float32)
    # two w x h boxes shifted in x by (w - dx): IoU = dx / (2w - dx); aim at the threshold, then nudge
    # the shift by -24..24 ulps
    tt = min(max(float(t), 1e-7), 1.0)
    dx = (2.0 * w.astype(np.float64) * tt / (1.0 + tt)).astype(np.float32)
    shift = (w - dx).astype(np.float32)
    k = rng.integers(-24, 25, n)
    for _ in range(24):
        shift = np.where(k > 0, np.nextafter(shift, np.float32(np.inf)),
                         np.where(k < 0, np.nextafter(shift, np.float32(-np.inf)), shift))
        k = k - np.sign(k)
    x0 = rng.uniform(0, 500, n).astype(np.float32)
    y0 = rng.uniform(0, 500, n).astype(np.float32)
    pred = np.zeros((n, 2, 290), np.float32)
    cx0, cy0 = x0 + w / 2, y0 + h / 2
    pred[:, 0, 0], pred[:, 0, 1] = cx0, cy0
    pred[:, 1, 0], pred[:, 1, 1] = (cx0 + shift).astype(np.float32), cy0
    pred[:, :, 2], pred[:, :, 3] = w[:, None], h[:, None]
    pred[:, :, 4] = 1.0
    pred[:, 0, 13:] = 0.9      # the first box wins the order
    pred[:, 1, 13:] = 0.8
    return torch.from_numpy(pred)


def maxnms_input(A=33600, img=1280, seed=20261019):
    """One ``[A, 290]`` image in which EVERY anchor passes ``conf_thres = 0`` and all ``A`` scores are
    distinct -- so the reference's more-than-30000-candidates cut (nms.py:115-116, an unstable argsort)
    has exactly one answer.  Each group carries one class at ``v = k / 65536`` (k distinct per anchor)
    and zeros elsewhere: sums of up to eight equal multiples of 2^-16 are exact, so both 8-term means
    equal ``v`` exactly.  numpy's PCG64 only, so the bits are the same everywhere (pinned by SHA-256)."""
    rng = np.random.default_rng(seed)
    x = np.zeros((A, 290), np.float32)
    centres = rng.uniform(0, img, (192, 2))
    which = rng.integers(0, 192, A)
    cxy = centres[which] + rng.normal(0, 14.0, (A, 2))
    wh = rng.uniform(20, 90, (A, 2))
    x[:, 0:2], x[:, 2:4], x[:, 4] = cxy.astype(np.float32), wh.astype(np.float32), 1.0
    x[:, 5:13] = (np.tile(cxy, 4) + (rng.uniform(size=(A, 8)) - 0.5) * np.tile(wh, 4)).astype(np.float32)
    k = rng.permutation(np.arange(6000, 6000 + A))          # distinct, 8 * k < 2^24
    v = (k / 65536.0).astype(np.float32)
    widths = (31, 24, 37, 37, 37, 37, 37, 37)
    col = 13
    for w in widths:
        j = rng.integers(0, w, A)
        x[np.arange(A), col + j] = v
        col += w
    return torch.from_numpy(x)
