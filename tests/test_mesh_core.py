"""CPU tests of the data-parallel mesh stage (stereo-vision_b200/csrc/mesh_core.h): the phases the k_mesh
kernels run between CTA barriers, executed here thread by thread in scrambled order
(tests/native/mesh_emulate.cpp), must reproduce the sequential host stage -- lattice filters incl. their
scan-order dependence, the support list, Triangle's triangles in Triangle's order, the raster units.
The host stage itself is pinned to the oracle / reference in tests/test_abi.py and tests/test_oracle.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import checkers
import elas_b200
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("mesh") / "libmesh_emu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", so,
                           os.path.join(ROOT, "tests", "native", "mesh_emulate.cpp")])
    lib = C.CDLL(so)
    lib.mesh_emulate_lattice.argtypes = [C.c_int] * 6 + [C.c_void_p] * 3 + [C.c_int, C.c_uint, C.POINTER(C.c_int)]
    lib.mesh_emulate_delaunay.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int),
                                          C.c_int, C.c_uint]
    return lib


def run_lattice(emu, d, step, p, nthr, seed):
    Hc, Wc = d.shape
    dcan = np.ascontiguousarray(d, np.int16).copy()
    incon = np.empty_like(dcan)
    sup = np.zeros((Wc * Hc, 3), np.int32)
    rounds = C.c_int(0)
    n = emu.mesh_emulate_lattice(Wc, Hc, step, p.incon_window_size, p.incon_threshold, p.incon_min_support,
                                 dcan.ctypes.data, incon.ctypes.data, sup.ctypes.data, nthr, seed, C.byref(rounds))
    return dcan, sup[:n].copy(), rounds.value


def run_delaunay(emu, sup, right, W, H, nthr, seed, unit_cap=1 << 20):
    sup = np.ascontiguousarray(sup, np.int32)
    n = len(sup)
    tri = np.zeros((2 * n + 8, 3), np.int32)
    units = np.zeros((1 << 18, 2), np.int32)
    ovf = np.zeros(2 * n + 8, np.int32)
    nu, no = C.c_int(0), C.c_int(0)
    nt = emu.mesh_emulate_delaunay(sup.ctypes.data, n, right, W, H, 32, unit_cap, tri.ctypes.data, units.ctypes.data,
                                   C.byref(nu), ovf.ctypes.data, C.byref(no), nthr, seed)
    return nt, tri[:max(nt, 0)].copy(), units[:nu.value].copy(), ovf[:no.value].copy()


def expected_units(sup, tri, right, W, H, band=32):
    x = sup[:, 0] - sup[:, 2] if right else sup[:, 0]
    y = sup[:, 1]
    out = []
    for t, (a, b, c) in enumerate(tri):
        u_lo, u_hi = max(min(x[a], x[b], x[c]), 0), min(max(x[a], x[b], x[c]), W)
        v_lo, v_hi = max(min(y[a], y[b], y[c]) - 1, 0), min(max(y[a], y[b], y[c]) + 1, H)
        chunks, bands = (u_hi - u_lo + 31) // 32, (v_hi - v_lo + band - 1) // band
        if chunks <= 0 or bands <= 0:
            continue
        out += [(t | (right << 30), ch | (bd << 16)) for ch in range(chunks) for bd in range(bands)]
    return np.array(out, np.int32).reshape(-1, 2)


def test_random_lattices_match_the_sequential_host_stage(emu):
    rng = np.random.default_rng(5)
    p = elas_b200.stereomapper(63)
    max_rounds = 0
    for it in range(40):
        Wc, Hc = int(rng.integers(12, 90)), int(rng.integers(10, 50))
        base = rng.integers(0, 60, (Hc, Wc))
        smooth = (np.add.outer(np.arange(Hc), np.arange(Wc)) // 3) % 50
        d = np.where(rng.random((Hc, Wc)) < 0.5, smooth, base)
        d = np.minimum(d, np.maximum(5 * np.arange(Wc) - 5, 0)[None, :])     # K2 only returns d <= u - 5 (elas.cpp:384-387)
        d = np.where(rng.random((Hc, Wc)) < rng.choice([0.1, 0.3, 0.6]), -1, d).astype(np.int16)
        d[0, :] = 0; d[:, 0] = 0                      # calloc'ed row/column
        W, H = Wc * 5 - 2, Hc * 5 - 1
        want = elas_b200.host_stage(p, W, H, d)
        for nthr, seed in ((1, 0), (37, it), (256, 1000 + it)):
            dcan, sup, rounds = run_lattice(emu, d, 5, p, nthr, seed)
            max_rounds = max(max_rounds, rounds)
            assert np.array_equal(dcan, want["dcan"]), (it, nthr)
            assert np.array_equal(sup, want["support"]), (it, nthr)
        if len(want["support"]) >= 3:
            for right, key in ((0, "tri1"), (1, "tri2")):
                for nthr, seed in ((1, 0), (64, it)):
                    nt, tri, units, ovf = run_delaunay(emu, want["support"], right, W, H, nthr, seed)
                    if nt < 0:      # duplicate right-image points: the device path is not used for such parameters
                        assert right == 1
                        continue
                    assert np.array_equal(tri, want[key]), (it, key, nthr)
                    assert len(ovf) == 0 and np.array_equal(units, expected_units(want["support"], tri, right, W, H)), (it, key)
    assert max_rounds >= 3          # the inconsistency filter did cascade in some case


def test_full_size_lattice_of_the_metric_configuration(emu, oracle):
    """1242x375 d_max 255: the lattice K2 produces (taken from the oracle's stage dump)."""
    L, R, _ = synth.synthetic_pair(1242, 375, 255, 0)
    p = checkers.stereomapper(255)
    _, _, _, st = oracle.run_stages(L, R, p, names=["dcan_raw", "dcan", "support", "tri1", "tri2", "lattice_dims"])
    Wc, Hc = (int(v) for v in st["lattice_dims"])
    raw = st["dcan_raw"].reshape(Hc, Wc)
    dcan, sup, rounds = run_lattice(emu, raw, 5, elas_b200.stereomapper(255), 1024, 7)
    assert np.array_equal(dcan.ravel(), st["dcan"]) and np.array_equal(sup.ravel(), st["support"])
    for right, key in ((0, "tri1"), (1, "tri2")):
        nt, tri, units, ovf = run_delaunay(emu, sup, right, 1242, 375, 1024, 3)
        assert np.array_equal(tri.ravel(), st[key]), key
    # a small unit capacity: the triangles that do not fit are listed as overflow, the rest is unchanged
    nt, tri, units_all, _ = run_delaunay(emu, sup, 0, 1242, 375, 64, 1)
    nt, tri, units, ovf = run_delaunay(emu, sup, 0, 1242, 375, 64, 1, unit_cap=len(units_all) // 2)
    assert len(ovf) > 0 and np.array_equal(units, units_all[:len(units)])
    assert set(ovf.tolist()) == set((units_all[len(units):, 0] & 0x3FFFFFFF).tolist())


def test_degenerate_point_sets(emu, oracle):
    """Co-circular lattices, general position and collinear sets against the oracle's Triangle restatement."""
    rng = np.random.default_rng(12)
    for it in range(90):
        n = int(rng.integers(3, 400))
        mode = it % 3
        if mode == 0:      # stride-5 lattice, masses of co-circular quads
            pts = np.stack([rng.integers(1, 60, n) * 5, rng.integers(1, 40, n) * 5, np.zeros(n, np.int64)], 1)
        elif mode == 1:    # general position
            pts = np.stack([rng.integers(0, 2000, n), rng.integers(0, 1000, n), np.zeros(n, np.int64)], 1)
        else:              # all collinear
            pts = np.stack([rng.integers(1, 100, n) * 5, np.full(n, 50), np.zeros(n, np.int64)], 1)
        pts = np.unique(pts, axis=0).astype(np.int32)        # this path never sees duplicates
        if len(pts) < 3:
            continue
        nt, tri, _, _ = run_delaunay(emu, pts, 0, 4096, 4096, 128, it)
        assert nt >= 0 and np.array_equal(tri, oracle.delaunay(pts, 0)), (it, mode, len(pts))
