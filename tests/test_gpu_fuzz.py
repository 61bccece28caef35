"""Randomised differential parity (tools/fuzz_parity.py) with fixed seeds: random sizes, disparity ranges and parameter
blocks, every stage of a single-frame run and the final maps of batches over frame groups against the CPU oracle.
The MIDDLEBURY-family run is the regression test for plane windows around d_plane = -3, -2 (plane radius 3), which a
too narrow strip padding in k_matching once dropped."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


def test_fuzz_all_parameter_families():
    import fuzz_parity
    assert fuzz_parity.run_cases(24, seed=7, max_w=520, max_h=200) == []


def test_fuzz_plane_radius_3():
    import fuzz_parity
    assert fuzz_parity.run_cases(16, seed=11, kinds=("middlebury",), max_w=700, max_h=200) == []


def test_fuzz_view_fusion_and_filters():
    import fuzz_view
    assert fuzz_view.run_cases(12, seed=3) == []
