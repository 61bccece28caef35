"""CPU tests (-m "not gpu"): pin the plain-C oracle restatement.

1. against the golden vectors generated from the unmodified reference (tests/golden/make_golden.py)
   -- these run everywhere, also where /root/reference is absent;
2. against the compiled reference itself (oracle/_ref), stage by stage and bit for bit, on seeded
   synthetic pairs, with the parameter sets the reference's two call sites use
   (stereothread.cpp:76-80, main.cpp:61-62), subsampling, the MIDDLEBURY preset and edge cases.
"""
import numpy as np
import pytest

import checkers
import synth
import os

from helpers import (INT_STAGES, FLOAT_STAGES, bits_equal, full_golden_cases, golden_cases, load_full_golden,
                     load_golden)


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_golden(oracle, name):
    L, R, p, g = load_golden(name)
    rc, D1, D2, st = oracle.run_stages(L, R, p)
    assert rc == 0
    for k in INT_STAGES + FLOAT_STAGES:
        assert bits_equal(st[k], g[k]), f"{name}: stage {k} differs from the reference's golden vector"
    assert bits_equal(D1.ravel(), g["D1"]) and bits_equal(D2.ravel(), g["D2"])


@pytest.mark.parametrize("name", full_golden_cases())
def test_oracle_matches_full_size_real_pairs(oracle, name):
    """libelas/img/urban1..4 at full size (1344x391, the reference's KITTI-size real images): the
    restatement against the committed outputs of the reference."""
    L, R, p, g = load_full_golden(name)
    rc, D1, D2, st = oracle.run_stages(L, R, p)
    assert rc == 0
    for k in ("dcan", "support", "tri1", "tri2", "D1_raw", "D2_raw", "D1", "D2"):
        assert bits_equal(st[k], g[k].ravel()), f"{name}: stage {k} differs from the reference's golden vector"


REF_IMG = "/root/reference/libelas/img"
DEMO_PAIRS = ["cones", "aloe", "raindeer", "urban1", "urban2", "urban3", "urban4"]      # main.cpp:105-113


@pytest.mark.skipif(not os.path.isdir(REF_IMG), reason="the reference's demo images are only present in the build container")
@pytest.mark.parametrize("name", DEMO_PAIRS)
def test_oracle_equals_reference_on_the_demo_pairs(ref, oracle, name):
    """`./elas demo` (libelas/src/main.cpp:105-113) runs these seven pairs; every stage of the restatement
    is bit-identical to the compiled reference on each, with the demo's parameter set (main.cpp:61-62)
    and, for the street scenes, stereomapper's (stereothread.cpp:76-80)."""
    L = synth.read_pgm(f"{REF_IMG}/{name}_left.pgm")
    R = synth.read_pgm(f"{REF_IMG}/{name}_right.pgm")
    sets = [checkers.demo(255)] + ([checkers.stereomapper(255)] if name.startswith("urban") else [])
    for p in sets:
        rc, st = _all_stages_equal(ref, oracle, L, R, p)
        assert rc == 0 and len(st["support"]) // 3 > 500


def test_bandwidth_config_4096x2160(ref, oracle):
    """BASELINE config 4 geometry (4096x2160, d_max 256): final maps bit-identical to the reference."""
    L, R, _ = synth.synthetic_pair(4096, 2160, 256, 1)
    p = checkers.stereomapper(256)
    _, R1, R2 = ref.process(L, R, p)
    _, O1, O2 = oracle.process(L, R, p)
    assert bits_equal(R1, O1) and bits_equal(R2, O2)


def test_oracle_process_equals_run_stages(oracle):
    L, R, p, g = load_golden("synth_320x120_d63")
    rc, D1, D2 = oracle.process(L, R, p)
    assert rc == 0 and bits_equal(D1.ravel(), g["D1"]) and bits_equal(D2.ravel(), g["D2"])


def _all_stages_equal(ref, oracle, L, R, p):
    rc_r, D1r, D2r, sr = ref.run_stages(L, R, p)
    rc_o, D1o, D2o, so = oracle.run_stages(L, R, p)
    assert rc_r == rc_o
    for k, a in sr.items():
        assert k in so, k
        assert bits_equal(a, so[k]), f"stage {k}: {int((a != so[k]).sum()) if a.shape == so[k].shape else 'shape'}"
    return rc_r, sr


CASES = [
    ("stereomapper", 416, 200, 95, 0, lambda d: checkers.stereomapper(d)),
    ("demo", 416, 200, 95, 1, lambda d: checkers.demo(d)),
    ("ragged-width", 333, 131, 63, 2, lambda d: checkers.stereomapper(d)),
    ("subsampling", 416, 200, 95, 3, lambda d: checkers.stereomapper(d).copy(subsampling=1)),
    ("middlebury", 320, 160, 63, 4, lambda d: checkers.middlebury().copy(disp_max=d)),
    ("median", 320, 160, 63, 5, lambda d: checkers.demo(d).copy(filter_median=1)),
    ("dmin", 320, 160, 63, 6, lambda d: checkers.stereomapper(d).copy(disp_min=3)),
    # parameter corners of the fused kernels: no gap filling / no mean, both maps post-processed with a
    # narrower gap, plane radius 3 with a texture threshold, another grid cell size and speckle size, and a
    # lattice stride at which two support points can collapse onto one right-image point (duplicate
    # vertices: Triangle's randomised sort decides which one survives)
    ("gap0-nomean", 320, 160, 63, 7, lambda d: checkers.stereomapper(d).copy(ipol_gap_width=0, filter_adaptive_mean=0)),
    ("gap2-both", 320, 160, 63, 8, lambda d: checkers.demo(d).copy(ipol_gap_width=2)),
    ("radius3", 320, 160, 63, 9, lambda d: checkers.stereomapper(d).copy(sradius=3.0, match_texture=40)),
    ("grid16", 320, 160, 63, 10, lambda d: checkers.stereomapper(d).copy(grid_size=16, speckle_size=50)),
    ("step4", 320, 160, 63, 11, lambda d: checkers.stereomapper(d).copy(candidate_stepsize=4)),
]


@pytest.mark.parametrize("tag,W,H,dmax,seed,mk", CASES, ids=[c[0] for c in CASES])
def test_oracle_equals_reference_all_stages(ref, oracle, tag, W, H, dmax, seed, mk):
    L, R, _ = synth.synthetic_pair(W, H, dmax, seed)
    rc, st = _all_stages_equal(ref, oracle, L, R, mk(dmax))
    assert rc == 0 and len(st["support"]) >= 9


def test_reference_process_equals_its_stagewise_replay(ref):
    """The stage-dump harness is sound: replaying the private stages gives Elas::process's output."""
    L, R, _ = synth.synthetic_pair(416, 200, 95, 7)
    p = checkers.stereomapper(95)
    rc, D1, D2 = ref.process(L, R, p)
    rc2, E1, E2, _ = ref.run_stages(L, R, p)
    assert bits_equal(D1, E1) and bits_equal(D2, E2)


def test_full_size_k_config(ref, oracle):
    """BASELINE config 1 (1242x375, d_max 255): every stage bit-identical to the reference."""
    L, R, _ = synth.synthetic_pair(1242, 375, 255, 0)
    rc, st = _all_stages_equal(ref, oracle, L, R, checkers.stereomapper(255))
    assert len(st["support"]) // 3 == 535 and len(st["tri1"]) // 3 == 1003


def test_row_stride_wider_than_width(ref, oracle):
    """dims[2] != width (stereomapper passes OpenCV's widthStep, stereothread.cpp:111)."""
    L, R, _ = synth.synthetic_pair(322, 120, 63, 8)
    Lp = np.zeros((120, 324), np.uint8); Lp[:, :322] = L
    Rp = np.zeros((120, 324), np.uint8); Rp[:, :322] = R
    p = checkers.stereomapper(63)
    _, A1, A2 = oracle.process(Lp[:, :322], Rp[:, :322], p)   # strided views: dims[2] = 324
    _, B1, B2 = oracle.process(L, R, p)
    _, C1, C2 = ref.process(Lp[:, :322], Rp[:, :322], p)
    assert bits_equal(A1, B1) and bits_equal(A2, B2) and bits_equal(A1, C1) and bits_equal(A2, C2)


def test_too_few_support_points(ref, oracle):
    """Flat images: < 3 support points, the reference returns without touching D (elas.cpp:69-75)."""
    flat = np.full((100, 160), 90, np.uint8)
    p = checkers.stereomapper(63)
    rc, D1, D2 = oracle.process(flat, flat, p)
    assert rc == 1 and (D1 == -77).all() and (D2 == -77).all()
    _, R1, _ = ref.process(flat, flat, p)
    assert (R1 == -77).all()


def test_delaunay_matches_triangle_on_degenerate_sets(ref, oracle):
    """Triangle-compatible tie-breaking: co-circular lattices, duplicates, collinear points."""
    rng = np.random.default_rng(11)
    for it in range(120):
        n = int(rng.integers(3, 300))
        mode = it % 4
        if mode == 0:      # stride-5 lattice, masses of co-circular quads
            pts = np.stack([rng.integers(1, 60, n) * 5, rng.integers(1, 40, n) * 5, rng.integers(0, 60, n)], 1)
        elif mode == 1:    # general position
            pts = np.stack([rng.integers(0, 2000, n), rng.integers(0, 1000, n), rng.integers(0, 255, n)], 1)
        elif mode == 2:    # many exact duplicates
            pts = np.stack([rng.integers(1, 8, n) * 5, rng.integers(1, 8, n) * 5, rng.integers(0, 3, n) * 5], 1)
        else:              # all collinear
            pts = np.stack([rng.integers(1, 100, n) * 5, np.full(n, 50), rng.integers(0, 10, n)], 1)
        for right in (0, 1):
            a, b = ref.delaunay(pts, right), oracle.delaunay(pts, right)
            assert a.shape == b.shape and np.array_equal(a, b), (it, mode, n, right)
