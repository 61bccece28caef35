"""SURVEY 8(f) rank 1 -- the consumers of D1 inside StereoThread: HSV colour map
(stereothread.cpp:116-147) and back-projection with the intensity border gain (:180-255).

CPU part: the oracle's plain-C restatement is pinned, bit for bit, against the reference's own
statements (cut out of stereothread.cpp at build time and compiled by oracle/Makefile into
oracle/_ref/libview_ref.so) and against committed golden vectors generated from them.
GPU part (test_gpu_view.py): the CUDA kernels against the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import checkers  # noqa: E402
from view_cases import view_case, CASES  # noqa: E402


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.skipif(not checkers.have_view_ref(), reason="oracle/_ref/libview_ref.so not built")
@pytest.mark.parametrize("case", CASES)
def test_oracle_equals_reference_statements(case):
    I1, D1, view, H = view_case(case)
    ref, ora = checkers.ViewChecker("ref"), checkers.ViewChecker("oracle")
    assert same_bits(ref.colormap(D1), ora.colormap(D1))
    for name, a, b in zip("IDXYZ", ref.reproject(I1, D1, view, H), ora.reproject(I1, D1, view, H)):
        assert same_bits(a, b), f"{case}: {name}"


def test_oracle_matches_golden_vectors():
    """Golden vectors written by tests/golden/view/make_view_golden.py from the reference statements."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "view", "view_small.npz"))
    ora = checkers.ViewChecker("oracle")
    assert same_bits(ora.colormap(g["D1"]), g["color"])
    for name, a in zip("IDXYZ", ora.reproject(g["I1"], g["D1"], g["view"], g["H"])):
        assert same_bits(a, g[name]), name


def test_colormap_properties():
    """Size-independent properties: invalid and zero disparities are black, saturation at d >= 200,
    every coloured pixel has one channel at 1 and one at 0."""
    ora = checkers.ViewChecker("oracle")
    D1 = np.array([[-10, -1, 0, 1e-3, 50, 100, 199.5, 200, 255, 1000]], np.float32)
    c = ora.colormap(D1)[0]
    assert (c[:3] == 0).all()
    assert (c[7:] == c[7]).all() and tuple(c[7]) == (1.0, 0.0, 0.0)
    lit = c[3:]
    assert (lit.max(axis=1) == 1).all() and (lit.min(axis=1) == 0).all()


def _same_fusion(a, b, tag):
    cur_a, pd_a, pp_a, pc_a = a
    cur_b, pd_b, pp_b, pc_b = b
    valid = cur_a[1] > 0
    assert same_bits(cur_a[0], cur_b[0]) and same_bits(cur_a[1], cur_b[1]), f"{tag}: I / D"
    for k, name in ((2, "X"), (3, "Y"), (4, "Z")):      # X/Y/Z are defined where the fused map is valid
        assert np.array_equal(cur_a[k][valid].view(np.uint32), cur_b[k][valid].view(np.uint32)), f"{tag}: {name}"
    assert (pd_a is None) == (pd_b is None) and (pd_a is None or same_bits(pd_a, pd_b)), f"{tag}: previous D"
    assert same_bits(pp_a, pp_b), f"{tag}: points_prev ({len(pp_a)} vs {len(pp_b)})"
    assert same_bits(pc_a, pc_b), f"{tag}: points_curr ({len(pc_a)} vs {len(pc_b)})"


@pytest.mark.skipif(not checkers.have_view_ref(), reason="oracle/_ref/libview_ref.so not built")
@pytest.mark.parametrize("name", ["street", "tiny", "backwards"])
def test_fusion_oracle_equals_reference_statements(name):
    """StereoThread::addDisparityMapToReconstruction (stereothread.cpp:290-437) over a short
    sequence, the previous map of a frame being the fused map of the frame before."""
    from view_cases import fusion_sequence
    ref, ora = checkers.ViewChecker("ref"), checkers.ViewChecker("oracle")
    prev_r = prev_o = None
    merged = created = 0
    for k, (I1, D1, view, H) in enumerate(fusion_sequence(name)):
        r = ref.fuse(I1, D1, view, H, prev_r)
        o = ora.fuse(I1, D1, view, H, prev_o)
        _same_fusion(r, o, f"{name} frame {k}")
        if k:
            merged += int(((prev_o[1] > 0) & ~(o[1] > 0)).sum())          # previous points taken over by the fusion
            created += int((o[0][1] == 1).sum())
        prev_r, prev_o = r[0], o[0]
    if name == "street":
        assert merged > 100 and created > 10      # the sequence exercises the average and the create branches
