"""Seeded synthetic stereo pairs (the workload BASELINE.json's configs name).

Random multi-scale texture warped by a ground-plane disparity ramp plus six fronto-parallel boxes.
Pure numpy, deterministic for (W, H, dmax, seed); used by tests/, bench.py and smoke().
"""
import numpy as np


def synthetic_pair(width, height, dmax, seed=0):
    """Returns (left u8 [H][W], right u8 [H][W], disparity i32 [H][W] left-referenced ground truth)."""
    W, H = int(width), int(height)
    rng = np.random.default_rng(seed)

    def noise(s):
        h = (H + s - 1) // s + 1
        w = (W + dmax + s - 1) // s + 2
        n = rng.integers(0, 256, (h, w)).astype(np.float32)
        return np.kron(n, np.ones((s, s), np.float32))[:H, :W + dmax + 1]

    tex = 0.5 * noise(1) + 0.3 * noise(3) + 0.2 * noise(9)

    v = np.arange(H, dtype=np.float64)[:, None]
    ramp = np.clip((v - 0.35 * H) / (0.65 * H) * 0.45 * dmax + 4, 4, 0.45 * dmax)
    d = np.repeat(ramp, W, axis=1)
    for _ in range(6):
        x0 = int(rng.integers(0, W - W // 6))
        y0 = int(rng.integers(0, H - H // 3))
        bw = int(rng.integers(W // 12, W // 6))
        bh = int(rng.integers(H // 6, H // 3))
        dk = float(rng.integers(int(0.2 * dmax), int(0.8 * dmax)))
        d[y0:y0 + bh, x0:x0 + bw] = np.maximum(d[y0:y0 + bh, x0:x0 + bw], dk)
    d = np.round(d).astype(np.int32)

    uu = np.arange(W, dtype=np.int64)[None, :]
    rows = np.arange(H, dtype=np.int64)[:, None]
    right = tex[:, dmax:dmax + W]
    left = tex[rows, uu - d + dmax]
    to_u8 = lambda a: np.clip(a, 0, 255).astype(np.uint8)
    return np.ascontiguousarray(to_u8(left)), np.ascontiguousarray(to_u8(right)), d


def read_pgm(path):
    """Binary PGM (P5, maxval 255) -> u8 [H][W]."""
    with open(path, "rb") as f:
        data = f.read()
    tokens, pos = [], 0
    while len(tokens) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            while data[pos:pos + 1] != b"\n":
                pos += 1
            continue
        start = pos
        while not data[pos:pos + 1].isspace():
            pos += 1
        tokens.append(data[start:pos])
    pos += 1
    assert tokens[0] == b"P5" and int(tokens[3]) == 255, "binary 8-bit PGM expected"
    w, h = int(tokens[1]), int(tokens[2])
    return np.frombuffer(data, np.uint8, w * h, pos).reshape(h, w).copy()
