"""Seeded inputs for the Matcher feature-filter tests (shared by the CPU and GPU tests)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["kitti_like", "minimal", "narrow", "saturating", "urban_crop", "big"]


def filter_case(name):
    """uint8 image [h][bytes_per_line], bytes_per_line a multiple of 16 (filter.cpp:294 asserts it)."""
    rng = np.random.default_rng(CASES.index(name) + 21)
    if name == "kitti_like":
        return rng.integers(0, 256, (375, 1248), dtype=np.uint8)            # 1242 wide, padded to 1248 (matcher dims[2])
    if name == "minimal":
        return rng.integers(0, 256, (6, 16), dtype=np.uint8)                # the smallest image blob5x5 accepts
    if name == "narrow":
        return rng.integers(0, 256, (203, 16), dtype=np.uint8)              # every blob column but 11 straddles a row end
    if name == "saturating":
        I = np.zeros((64, 96), np.uint8)                                    # step edges: du/dv reach both clamps
        I[:, 48:] = 255; I[32:, :] = 255 - I[32:, :]
        I[5:9, 7:11] = rng.integers(0, 256, (4, 4), dtype=np.uint8)
        return I
    if name == "urban_crop":
        g = np.load(os.path.join(ROOT, "tests", "golden", "urban1_crop.npz"))
        I = g["I1"]
        h, w = I.shape
        out = np.zeros((h, (w + 15) // 16 * 16), np.uint8)
        out[:, :w] = I
        return out
    return rng.integers(0, 256, (1080, 1920), dtype=np.uint8)


def comparable(name, a):
    """The part of a feature map that the reference defines: du/dv[n-2], [n-1] come from reads beyond its buffers."""
    a = a.ravel()
    return a[:-2] if name in ("du", "dv") else a
