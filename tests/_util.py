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
