"""SURVEY 8(f) rank 4 -- the feature filters of libviso2's Matcher (libviso2/src/filter.cpp:474-530 as called at
matcher.cpp:799-801).  CPU part: the oracle's plain-C restatement against libviso2's own filter.cpp compiled by
oracle/Makefile, and against committed golden vectors generated from it.  GPU part: test_gpu_filters.py."""
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

NAMES = ("du", "dv", "f1", "f2")


@pytest.mark.skipif(not checkers.have_filter_ref(), reason="oracle/_ref/libvisofilter_ref.so not built")
@pytest.mark.parametrize("case", CASES)
def test_oracle_equals_reference_filters(case):
    I = filter_case(case)
    ref, ora = checkers.MatcherFilterChecker("ref"), checkers.MatcherFilterChecker("oracle")
    for name, a, b in zip(NAMES, ref(I), ora(I)):
        assert np.array_equal(comparable(name, a), comparable(name, b)), f"{case}: {name}"


def test_oracle_matches_golden_filter_vectors():
    g = np.load(os.path.join(ROOT, "tests", "golden", "viso", "filters_small.npz"))
    ora = checkers.MatcherFilterChecker("oracle")
    for case in ("saturating", "narrow"):
        for name, a in zip(NAMES, ora(g[f"{case}_I"])):
            assert np.array_equal(comparable(name, a), comparable(name, g[f"{case}_{name}"])), f"{case}: {name}"


def test_filter_properties():
    """Size-independent properties: a constant image has no gradient (du = dv = 128), no checkerboard response and
    a zero blob response (the mask sums to 0) away from the flat-index seams; on full-contrast step edges du/dv reach
    the extremes the 1/128 scaling allows, 128 +- 96 (the u8 saturation of the reference's packus never triggers)."""
    ora = checkers.MatcherFilterChecker("oracle")
    du, dv, f1, f2 = ora(np.full((40, 64), 77, np.uint8))
    assert (du[4:-4, 4:-4] == 128).all() and (dv[4:-4, 4:-4] == 128).all()
    assert (f2[4:-4, 4:-4] == 0).all() and (f1[3:-3, 3:-2] == 0).all()
    du, dv, _, _ = ora(filter_case("saturating"))
    for m in (du, dv):
        body = m.ravel()[2:]
        assert body.min() == 32 and body.max() == 223, (body.min(), body.max())


def test_filter_entry_refuses_bad_arguments_and_missing_device():
    lib = elas_b200.load_library()
    I = np.zeros((8, 32), np.uint8)
    out = [np.zeros((8, 32), np.uint8), np.zeros((8, 32), np.uint8), np.zeros((8, 32), np.int16), np.zeros((8, 32), np.int16)]
    call = lambda w, h: lib.elas_b200_matcher_filters(0, I.ctypes.data, w, h, *[a.ctypes.data for a in out], 1, None)
    assert call(24, 8) == -3 and call(32, 5) == -3                  # width not a multiple of 16 (filter.cpp:294), too few rows
    if lib.elas_b200_device_count() == 0:
        assert call(32, 8) == -1                                    # ELAS_B200_E_NO_DEVICE: no CPU fallback
        with pytest.raises(RuntimeError):
            elas_b200.matcher_filters(I)
