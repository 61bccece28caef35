"""CPU tests (-m "not gpu") of the C-ABI library and the host middle stage (no compute kernels run)."""
import os
import re

import numpy as np
import pytest

import checkers
import elas_b200
import synth
from helpers import bits_equal, golden_cases, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    elas_b200.build_library()
    return elas_b200.load_library()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "elas_b200.h")).read()
    declared = set(re.findall(r"\b(elas_b200_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/elas_b200.h but not exported"
    assert declared == set(elas_b200.EXPORTS)


def test_presets_match_reference_parameters(lib):
    assert bytes(elas_b200.robotics()) == bytes(checkers.robotics())        # elas.h:93-118
    assert bytes(elas_b200.middlebury()) == bytes(checkers.middlebury())    # elas.h:121-146
    assert bytes(elas_b200.stereomapper()) == bytes(checkers.stereomapper())  # stereothread.cpp:76-80


def test_presets_match_compiled_reference(lib, ref):
    for which, mine in ((0, elas_b200.robotics()), (1, elas_b200.middlebury())):
        p = checkers.Params()
        ref.lib.ref_default_params(__import__("ctypes").byref(p), which)
        assert bytes(p) == bytes(mine)


def test_no_device_fails_loudly(lib):
    """This container has no GPU: every compute entry point must refuse, never fall back to a CPU path."""
    if lib.elas_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    L, R, _ = synth.synthetic_pair(160, 100, 31, 0)
    with pytest.raises(RuntimeError):
        elas_b200.process(L, R, elas_b200.stereomapper(31))
    with pytest.raises(RuntimeError):
        elas_b200.ElasB200(elas_b200.stereomapper(31), 160, 100)


@pytest.mark.parametrize("name", golden_cases())
def test_host_stage_matches_reference_golden(lib, name):
    """Lattice filters + support list + Delaunay + planes (host_stage.cc) on the reference's own lattice."""
    L, R, p, g = load_golden(name)
    Wc, Hc = g["lattice_dims"]
    out = elas_b200.host_stage(elas_b200.Params.from_buffer_copy(bytes(p)), L.shape[1], L.shape[0],
                               g["dcan_raw"].reshape(Hc, Wc))
    assert out["rc"] == 0
    for k in ["dcan", "support", "tri1", "tri2", "planes1", "planes2"]:
        assert bits_equal(out[k].ravel(), g[k]), k


def test_host_stage_matches_oracle_full_size(lib, oracle):
    L, R, _ = synth.synthetic_pair(1242, 375, 255, 0)
    p = checkers.stereomapper(255)
    _, _, _, st = oracle.run_stages(L, R, p, names=["lattice_dims", "dcan_raw", "dcan", "support", "tri1", "tri2", "planes1", "planes2"])
    Wc, Hc = st["lattice_dims"]
    out = elas_b200.host_stage(elas_b200.stereomapper(255), 1242, 375, st["dcan_raw"].reshape(Hc, Wc))
    for k in ["dcan", "support", "tri1", "tri2", "planes1", "planes2"]:
        assert bits_equal(out[k].ravel(), st[k]), k
    assert len(out["support"]) == 535 and len(out["tri1"]) == 1003


def test_host_stage_random_lattices_match_oracle(lib, oracle):
    """Scan-order-dependent filters + Triangle tie-breaks on synthetic lattices (no images needed)."""
    rng = np.random.default_rng(3)
    p = checkers.stereomapper(63)
    for it in range(25):
        Wc, Hc = int(rng.integers(12, 70)), int(rng.integers(10, 40))
        base = rng.integers(0, 60, (Hc, Wc))
        smooth = (np.add.outer(np.arange(Hc), np.arange(Wc)) // 3) % 50
        d = np.where(rng.random((Hc, Wc)) < 0.5, smooth, base)
        d = np.where(rng.random((Hc, Wc)) < 0.3, -1, d).astype(np.int16)
        d[0, :] = 0; d[:, 0] = 0                      # calloc'ed row/column
        W, H = Wc * 5 - 2, Hc * 5 - 1
        mine = elas_b200.host_stage(elas_b200.stereomapper(63), W, H, d)
        want = d.copy()
        oracle.lib.oracle_lattice_filters(__import__("ctypes").byref(p), want.ctypes.data, Wc, Hc)
        assert np.array_equal(mine["dcan"], want), it
        sup = [(u * 5, v * 5, int(want[v, u])) for u in range(1, Wc) for v in range(1, Hc) if want[v, u] >= 0]
        sup = np.array(sup, np.int32).reshape(-1, 3)
        assert np.array_equal(mine["support"], sup), it
        if len(sup) >= 3:
            for right, key in ((0, "tri1"), (1, "tri2")):
                assert np.array_equal(mine[key], oracle.delaunay(sup, right)), (it, key)
