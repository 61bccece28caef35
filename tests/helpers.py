"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

import checkers
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

INT_STAGES = ["dcan_raw", "dcan", "support", "tri1", "tri2"]
FLOAT_STAGES = ["planes1", "planes2", "D1_raw", "D2_raw", "D1_lr", "D2_lr", "D1_seg", "D1_gap", "D1", "D2"]


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    p = checkers.Params.from_buffer_copy(g.pop("params").tobytes())
    if "I1" in g:
        L, R = g.pop("I1"), g.pop("I2")
    else:
        assert name.startswith("synth_320x120_d63")
        L, R, _ = synth.synthetic_pair(320, 120, 63, seed=3)
    return L, R, p, g


def full_golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "full", "*.npz")))


def load_full_golden(name):
    """Full-size real pairs (libelas/img/urban1..4, 1344x391): inputs, parameter block and the reference's
    outputs; integer-valued maps are stored as int16 and widened back to float32 here (exact)."""
    g = dict(np.load(os.path.join(GOLDEN, "full", name + ".npz")))
    p = checkers.Params.from_buffer_copy(g.pop("params").tobytes())
    L, R = g.pop("I1"), g.pop("I2")
    for k in ("D1_raw", "D2_raw", "D2"):
        g[k] = g[k].astype(np.float32)
    return L, R, p, g


def bits_equal(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype == np.float32:
        return np.array_equal(a.view(np.uint32), b.view(np.uint32))
    return np.array_equal(a, b)


def disparity_report(got, want, tol=1.0):
    """Mismatch statistics between two disparity maps: validity mask and |delta| on both-valid."""
    got, want = np.asarray(got).ravel(), np.asarray(want).ravel()
    vg, vw = got >= 0, want >= 0
    both = vg & vw
    delta = np.abs(got[both] - want[both])
    return {
        "mask_mismatch": int((vg != vw).sum()),
        "over_tol": int((delta > tol).sum()),
        "max_delta": float(delta.max()) if delta.size else 0.0,
        "valid": int(vw.sum()),
    }
