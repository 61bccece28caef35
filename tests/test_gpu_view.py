"""GPU parity of the colour-map and back-projection kernels (k_view.cu) against the oracle, through the
C ABI (elas_b200_colormap / elas_b200_reproject); bit-exact."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200"))
import checkers  # noqa: E402
import elas_b200  # noqa: E402
import synth  # noqa: E402
from view_cases import view_case, CASES  # noqa: E402

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("case", CASES)
def test_view_kernels_match_oracle_on_given_maps(case):
    I1, D1, view, H = view_case(case)
    h, w = D1.shape
    ora = checkers.ViewChecker("oracle")
    e = elas_b200.ElasB200(elas_b200.stereomapper(63), w, h, n_slots=1)
    try:
        color = e.colormap(D1)
        outs = e.reproject(view, H, I1=I1, D1=D1)
    finally:
        e.close()
    assert same_bits(color, ora.colormap(D1))
    for name, a, b in zip("IDXYZ", outs, ora.reproject(I1, D1, view, H)):
        assert same_bits(a, b), f"{case}: {name}"


def test_view_of_the_frame_left_on_the_device():
    """process() then colormap()/reproject() with no D1/I1: the maps never leave HBM in between."""
    L, R, _ = synth.synthetic_pair(640, 240, 127, 7)
    view = np.array([721.5377, 320.0, 120.0, 0.5371657, 30.0, 1.2], np.float32)
    H = np.hstack([np.eye(3), [[0.1], [0.2], [0.3]]])
    ora = checkers.ViewChecker("oracle")
    e = elas_b200.ElasB200(elas_b200.stereomapper(127), 640, 240, n_slots=1)
    try:
        rc, D1, _ = e.process(L, R)
        color = e.colormap()
        outs = e.reproject(view, H)
    finally:
        e.close()
    assert rc == 0 and (D1 >= 0).mean() > 0.3
    assert same_bits(color, ora.colormap(D1))
    for name, a, b in zip("IDXYZ", outs, ora.reproject(L, D1, view, H)):
        assert same_bits(a, b), name
    assert (outs[4][outs[1] > 0] > 0).all()          # reconstructed points lie in front of the camera


def test_colormap_with_subsampling():
    L, R, _ = synth.synthetic_pair(640, 240, 127, 8)
    p = elas_b200.stereomapper(127).copy(subsampling=1)
    e = elas_b200.ElasB200(p, 640, 240, n_slots=1)
    try:
        rc, D1, _ = e.process(L, R)
        color = e.colormap()
    finally:
        e.close()
    assert D1.shape == (120, 320) and same_bits(color, checkers.ViewChecker("oracle").colormap(D1))


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


@pytest.mark.parametrize("name", ["street", "tiny", "backwards"])
def test_fusion_matches_oracle_over_a_sequence(name):
    """elas_b200_fuse against the oracle's addDisparityMapToReconstruction
    (stereothread.cpp:290-437) frame after frame, each side feeding on its own previous fused map."""
    from view_cases import fusion_sequence
    ora = checkers.ViewChecker("oracle")
    seq = fusion_sequence(name)
    h, w = seq[0][1].shape
    e = elas_b200.ElasB200(elas_b200.stereomapper(63), w, h, n_slots=1)
    try:
        prev_o = prev_g = None
        for k, (I1, D1, view, H) in enumerate(seq):
            o = ora.fuse(I1, D1, view, H, prev_o)
            cur = e.reproject(view, H, I1=I1, D1=D1)
            g = e.fuse(view, H, cur, prev_g)
            _same_fusion(o, g, f"{name} frame {k}")
            prev_o, prev_g = o[0], g[0]
    finally:
        e.close()


def test_fusion_many_previous_points_on_one_pixel():
    """Degenerate previous map: every point is the same 3-d point, so all of them project onto one current pixel
    and the running average is a chain as long as the image (the long-list path of k_fuse_apply)."""
    rng = np.random.default_rng(3)
    w, h = 160, 96
    view = np.array([300.0, 80.0, 48.0, 0.54, 30.0, 1.2], np.float32)
    H = np.hstack([np.eye(3), [[0.0], [0.0], [0.0]]])
    I1 = rng.integers(0, 256, (h, w), dtype=np.uint8)
    D1 = np.full((h, w), 20.0, np.float32)
    ora = checkers.ViewChecker("oracle")
    base = list(ora.reproject(I1, D1, view, H))
    prev = [a.copy() for a in base]
    prev[0] = rng.random((h, w)).astype(np.float32)
    prev[2][:] = base[2][40, 70]; prev[3][:] = base[3][40, 70]; prev[4][:] = base[4][40, 70]
    prev[1][rng.random((h, w)) < 0.1] = -1
    e = elas_b200.ElasB200(elas_b200.stereomapper(63), w, h, n_slots=1)
    try:
        g = e.fuse(view, H, base, prev)
    finally:
        e.close()
    lib = ora.lib
    import ctypes as C
    cur = [a.copy() for a in base]
    pv = [a.copy() for a in prev]
    n = w * h
    pp, pc = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32)
    n_p, n_c = C.c_int32(0), C.c_int32(0)
    lib.oracle_fuse.restype = None
    lib.oracle_fuse.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p] + [C.c_void_p] * 10 + \
                               [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int32)]
    Hd = np.ascontiguousarray(H, np.float64).reshape(12)
    lib.oracle_fuse(w, h, view.ctypes.data, Hd.ctypes.data, *[a.ctypes.data for a in pv], *[a.ctypes.data for a in cur],
                    pp.ctypes.data, C.byref(n_p), pc.ctypes.data, C.byref(n_c))
    o = (cur, pv[1], pp[:n_p.value], pc[:n_c.value])
    assert int((pv[1] == -1).sum()) > n // 2          # the chain was long
    _same_fusion(o, g, "one pixel")
