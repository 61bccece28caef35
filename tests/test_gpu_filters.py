"""GPU parity of the Matcher feature filters (k_viso.cu) against the oracle and the committed golden vectors,
through the C ABI (elas_b200_matcher_filters); bit-exact."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200"))
import checkers  # noqa: E402
import elas_b200  # noqa: E402
from filter_cases import CASES, filter_case, comparable  # noqa: E402

pytestmark = pytest.mark.gpu
NAMES = ("du", "dv", "f1", "f2")


@pytest.mark.parametrize("case", CASES)
def test_filters_match_oracle(case):
    I = filter_case(case)
    ora = checkers.MatcherFilterChecker("oracle")
    for name, a, b in zip(NAMES, elas_b200.matcher_filters(I), ora(I)):
        assert np.array_equal(a, b), f"{case}: {name} ({int((a != b).sum())} differ, first at {np.argwhere(a != b)[:3].tolist()})"


def test_filters_match_reference_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "viso", "filters_small.npz"))
    for case in ("saturating", "narrow"):
        for name, a in zip(NAMES, elas_b200.matcher_filters(g[f"{case}_I"])):
            assert np.array_equal(comparable(name, a), comparable(name, g[f"{case}_{name}"])), f"{case}: {name}"


def test_filters_on_device_buffers_in_place():
    """Device pointers are read and written in place (no staging copies); same bits as the host-buffer call."""
    import ctypes as C
    import torch
    I = filter_case("kitti_like")
    h, w = I.shape
    d_I = torch.from_numpy(I).cuda()
    d_u8 = [torch.zeros((h, w), dtype=torch.uint8, device="cuda") for _ in range(2)]
    d_i16 = [torch.zeros((h, w), dtype=torch.int16, device="cuda") for _ in range(2)]
    lib = elas_b200.load_library()
    ms = C.c_float(0)
    rc = lib.elas_b200_matcher_filters(0, d_I.data_ptr(), w, h, d_u8[0].data_ptr(), d_u8[1].data_ptr(),
                                       d_i16[0].data_ptr(), d_i16[1].data_ptr(), 5, C.byref(ms))
    torch.cuda.synchronize()
    assert rc == 0 and ms.value > 0
    for name, a, b in zip(NAMES, [t.cpu().numpy() for t in d_u8 + d_i16], elas_b200.matcher_filters(I)):
        assert np.array_equal(a, b), name
