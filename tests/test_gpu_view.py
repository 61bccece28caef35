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
