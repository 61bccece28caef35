"""Seeded inputs for the colour-map / back-projection tests (shared by the CPU and GPU tests)."""
import numpy as np

CASES = ["kitti_like", "odd_small", "no_gain", "far_and_near"]


def view_case(name):
    rng = np.random.default_rng(CASES.index(name) + 1)
    if name == "kitti_like":
        w, h, step = 1242, 375, 1244
    elif name == "odd_small":
        w, h, step = 97, 53, 100
    elif name == "no_gain":
        w, h, step = 640, 480, 640
    else:
        w, h, step = 416, 200, 416
    I1 = rng.integers(0, 256, (h, step), dtype=np.uint8)[:, :]
    # disparity map in the shapes Elas::process produces: -10 invalid, integers, halves, means
    D1 = rng.integers(0, 256, (h, w)).astype(np.float32)
    D1 += rng.choice(np.array([0, 0, 0.5, 1 / 3, 0.1], np.float32), (h, w))
    D1[rng.random((h, w)) < 0.3] = -10
    if name == "far_and_near":
        D1[:, : w // 4] = np.float32(0.25)        # z beyond max_dist
        D1[:, w // 4 : w // 2] = np.float32(4000)  # z below 0.1
        D1[0, 0] = 0
    view = np.array([721.5377, 609.5593, 172.854, 0.5371657, 20.0, 0.0 if name == "no_gain" else 1.35], np.float32)
    ang = 0.1
    H = np.array([[np.cos(ang), 0.02, np.sin(ang), 0.3],
                  [-0.01, 0.999, 0.03, -0.12],
                  [-np.sin(ang), 0.01, np.cos(ang), 1.7]], np.float64)
    view_I1 = np.ascontiguousarray(I1)
    # the image handed to the checkers keeps its row pitch (step) in strides[0]
    return np.lib.stride_tricks.as_strided(view_I1, (h, w), (step, 1)), D1, view, H


def fusion_sequence(name="street", frames=3):
    """A short sequence for the map-fusion tests: a slanted ground plane plus boxes seen from a camera that moves
    forward and yaws a little; disparities carry noise and holes so that every branch of the fusion loop is taken
    (average, create, keep apart, outside the image, out of range).  Yields (I1, D1, view, H) per frame."""
    rng = np.random.default_rng({"street": 7, "tiny": 11, "backwards": 13}[name])
    w, h = (97, 53) if name == "tiny" else (416, 200)
    f, cu, cv, base = 300.0, w / 2 - 3.5, h / 2 + 1.25, 0.54
    view = np.array([f, cu, cv, base, 12.0 if name == "tiny" else 30.0, 1.2], np.float32)
    step = -1.5 if name == "backwards" else 0.9       # moving backwards piles several previous points on one pixel
    out = []
    for k in range(frames):
        I1 = rng.integers(0, 256, (h, w), dtype=np.uint8)
        v = np.arange(h, dtype=np.float32)[:, None] + np.zeros((1, w), np.float32)
        depth = np.where(v > cv + 2, f * 1.6 / np.maximum(v - cv, 1e-3), 25.0) - step * k      # the ground comes closer
        D1 = (f * base / np.maximum(depth, 0.3)).astype(np.float32)
        D1 = np.round(D1 * 2) / 2 + rng.choice(np.array([0, 0, 0.25, -0.5], np.float32), (h, w))
        D1[rng.random((h, w)) < 0.25] = -10
        D1[:, : w // 8] = np.float32(0.2)           # beyond max_dist
        ang = 0.02 * k
        H = np.array([[np.cos(ang), 0.0, np.sin(ang), 0.05 * k],
                      [0.0, 1.0, 0.0, -0.01 * k],
                      [-np.sin(ang), 0.0, np.cos(ang), step * k]], np.float64)
        out.append((I1, D1.astype(np.float32), view, H))
    return out
