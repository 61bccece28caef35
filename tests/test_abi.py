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


def test_drop_in_class_compiles_and_refuses_without_device(lib, tmp_path):
    """The C++ drop-in Elas class builds against its own header with the reference call-site code, and
    without a GPU it reports the failure and marks the maps invalid instead of computing on the CPU."""
    import subprocess
    src = os.path.join(ROOT, "stereo-vision_b200", "dropin", "libelas", "src")
    exe = str(tmp_path / "dropin_demo")
    subprocess.check_call(["g++", "-O2", "-std=c++11", "-Wall", "-I" + src, os.path.join(ROOT, "tests", "dropin_demo.cpp"),
                           os.path.join(src, "elas.cpp"), os.path.join(src, "descriptor.cpp"), "-ldl", "-o", exe])
    if lib.elas_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    L, R, _ = synth.synthetic_pair(160, 100, 31, 0)
    L.tofile(str(tmp_path / "l.raw")); R.tofile(str(tmp_path / "r.raw"))
    env = dict(os.environ, ELAS_B200_LIB=elas_b200.LIB_PATH)
    r = subprocess.run([exe, "stereomapper", "160", "100", "31", str(tmp_path / "l.raw"), str(tmp_path / "r.raw"),
                        str(tmp_path / "d1.out"), str(tmp_path / "d2.out")], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "elas_b200_process failed" in r.stderr
    assert (np.fromfile(str(tmp_path / "d1.out"), np.float32) == -10).all()


def test_drop_in_header_matches_reference_header(tmp_path):
    """Field order, types and preset values of Elas::parameters are those of the reference header."""
    import subprocess
    ref_hdr = "/root/reference/libelas/src"
    if not os.path.isdir(ref_hdr):
        pytest.skip("/root/reference not present")
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "elas.h"
#define F(x) printf("%s %zu %zu ", #x, offsetof(Elas::parameters, x), sizeof(((Elas::parameters*)0)->x))
int main() {
  F(disp_min); F(disp_max); F(support_threshold); F(support_texture); F(candidate_stepsize); F(incon_window_size);
  F(incon_threshold); F(incon_min_support); F(add_corners); F(grid_size); F(beta); F(gamma); F(sigma); F(sradius);
  F(match_texture); F(lr_threshold); F(speckle_sim_threshold); F(speckle_size); F(ipol_gap_width); F(filter_median);
  F(filter_adaptive_mean); F(postprocess_only_left); F(subsampling);
  printf("size %zu\n", sizeof(Elas::parameters));
  for (int s = 0; s < 2; s++) { Elas::parameters p(s ? Elas::MIDDLEBURY : Elas::ROBOTICS);
    printf("%d %d %.6f %d %d %d %d %d %d %d %.6f %.6f %.6f %.6f %d %d %.6f %d %d %d %d %d %d\n", p.disp_min, p.disp_max,
      p.support_threshold, p.support_texture, p.candidate_stepsize, p.incon_window_size, p.incon_threshold,
      p.incon_min_support, (int)p.add_corners, p.grid_size, p.beta, p.gamma, p.sigma, p.sradius, p.match_texture,
      p.lr_threshold, p.speckle_sim_threshold, p.speckle_size, p.ipol_gap_width, (int)p.filter_median,
      (int)p.filter_adaptive_mean, (int)p.postprocess_only_left, (int)p.subsampling); }
  return 0; }
'''
    (tmp_path / "layout.cpp").write_text(prog)
    mine = os.path.join(ROOT, "stereo-vision_b200", "dropin", "libelas", "src")
    outs = []
    for inc, extra in ((ref_hdr, []), (mine, [os.path.join(mine, "elas.cpp"), "-ldl"])):
        exe = str(tmp_path / ("layout_" + ("ref" if inc == ref_hdr else "mine")))
        subprocess.check_call(["g++", "-std=c++11", "-msse3", "-w", "-I" + inc, str(tmp_path / "layout.cpp")] + extra + ["-o", exe])
        outs.append(subprocess.check_output([exe], text=True))
    assert outs[0] == outs[1]
