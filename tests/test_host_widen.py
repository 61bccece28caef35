"""CPU test of the host half of the narrowed right map (stereo-vision_b200/csrc/host_widen.h): int16 -> float and
u8 + validity bits -> float, for every destination alignment, widths that are no multiple of 16 or 32 and rows whose
validity words straddle 32-bit boundaries.  The GPU half (k_lr_rows) is covered by tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def widen(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("widen") / "libwiden_test.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", so,
                           os.path.join(ROOT, "tests", "native", "widen_test.cpp")])
    lib = C.CDLL(so)
    lib.test_widen_i16.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.test_widen_u8_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    return lib


def _dst(n, offset_floats):
    """n floats starting `offset_floats` floats behind a 16-byte boundary, with guard values around them."""
    raw = np.full(n + 16, 123.0, np.float32)
    start = (-raw.ctypes.data // 4) % 4 + offset_floats
    guard = np.ones(n + 16, bool)
    guard[start:start + n] = False
    return (raw, guard), raw[start:start + n]


@pytest.mark.parametrize("offset", [0, 1, 2, 3])
def test_widen_int16(widen, offset):
    rng = np.random.default_rng(offset)
    for n in (1, 7, 8, 9, 1242 * 3 + 5):
        src = rng.integers(-10, 4096, n).astype(np.int16)
        raw, dst = _dst(n, offset)
        widen.test_widen_i16(src.ctypes.data, dst.ctypes.data, n)
        assert np.array_equal(dst, src.astype(np.float32))
        assert (raw[0][raw[1]] == 123.0).all()     # nothing written outside


@pytest.mark.parametrize("offset", [0, 1, 2, 3])
@pytest.mark.parametrize("Dw,Dh", [(16, 3), (31, 4), (33, 5), (512, 6), (629, 7), (1242, 5), (47, 9)])
def test_widen_u8_with_validity_bits(widen, offset, Dw, Dh):
    rng = np.random.default_rng(Dw * 7 + offset)
    wpr = (Dw + 31) // 32
    vals = rng.integers(0, 256, (Dh, Dw), dtype=np.uint8)
    valid = rng.random((Dh, Dw)) < 0.7
    valid[0, :] = True; valid[-1, :] = False
    mask = np.zeros((Dh, wpr + 1), np.uint32)                     # one spare word: the landing buffer has slack as well
    for u in range(Dw):
        mask[:, u // 32] |= (valid[:, u].astype(np.uint32) << np.uint32(u % 32))
    mask[:, wpr] = 0xFFFFFFFF                                     # garbage beyond the row must not matter
    mask_rows = np.ascontiguousarray(mask[:, :wpr])
    raw, dst = _dst(Dw * Dh, offset)
    widen.test_widen_u8_mask(vals.ctypes.data, mask_rows.ctypes.data, wpr, dst.ctypes.data, Dw, Dh)
    want = np.where(valid, vals.astype(np.float32), np.float32(-10)).ravel()
    assert np.array_equal(dst, want)
    assert (raw[0][raw[1]] == 123.0).all()
