"""Randomised differential test of the D1 consumers (colour map, back-projection, map fusion over short sequences) and of
the Matcher feature filters against the CPU oracle.  Usage: python tools/fuzz_view.py [cases] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import checkers, elas_b200


def bits(a, b):
    return a.shape == b.shape and np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def run_cases(cases, seed=1):
    rng = np.random.default_rng(seed)
    view_o, filt_o = checkers.ViewChecker("oracle"), checkers.MatcherFilterChecker("oracle")
    failures = []
    t0 = time.time()
    for case in range(cases):
        # ---- feature filters
        h, w = int(rng.integers(6, 300)), 16 * int(rng.integers(1, 90))
        kind = rng.integers(0, 4)
        I = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if kind == 1: I[:] = rng.integers(0, 256)
        if kind == 2: I = (np.add.outer(np.arange(h), np.arange(w)) * int(rng.integers(1, 9)) % 256).astype(np.uint8)
        if kind == 3: I[rng.random((h, w)) < 0.5] = 255
        for name, a, b in zip(("du", "dv", "f1", "f2"), elas_b200.matcher_filters(I), filt_o(I)):
            if not bits(a, b): failures.append((f"case {case}: filters {w}x{h} kind {kind}", name))
        # ---- colour map, back-projection, fusion over a short sequence
        W, H = int(rng.integers(96, 500)), int(rng.integers(48, 260))
        f = float(rng.uniform(150, 800)); base = float(rng.uniform(0.1, 0.8))
        view = np.array([f, W / 2 + rng.uniform(-5, 5), H / 2 + rng.uniform(-5, 5), base, float(rng.choice([8.0, 20.0, 50.0])),
                         float(rng.choice([0.0, 1.2, 2.0]))], np.float32)
        e = elas_b200.ElasB200(elas_b200.stereomapper(31), W, H, n_slots=1)
        try:
            prev_o = prev_g = None
            motion = rng.uniform(-1.5, 1.5)
            for k in range(int(rng.integers(1, 4))):
                pitch = W + int(rng.integers(0, 9))
                I1 = rng.integers(0, 256, (H, pitch), dtype=np.uint8)[:, :W]
                D1 = rng.uniform(0.5, 60, (H, W)).astype(np.float32)
                if rng.random() < 0.5: D1 = np.round(D1)
                D1 += rng.choice(np.array([0, 0, 0.5, 0.25], np.float32), (H, W))
                D1[rng.random((H, W)) < rng.uniform(0, 0.6)] = -10
                ang = rng.uniform(-0.05, 0.05) * k
                Hm = np.array([[np.cos(ang), rng.uniform(-0.01, 0.01), np.sin(ang), rng.uniform(-0.2, 0.2) * k],
                               [rng.uniform(-0.01, 0.01), 1.0, rng.uniform(-0.01, 0.01), rng.uniform(-0.05, 0.05) * k],
                               [-np.sin(ang), rng.uniform(-0.01, 0.01), np.cos(ang), motion * k]], np.float64)
                if not bits(e.colormap(D1), view_o.colormap(D1)): failures.append((f"case {case}: colormap {W}x{H}", "colour"))
                cur = e.reproject(view, Hm, I1=I1, D1=D1)
                for name, a, b in zip("IDXYZ", cur, view_o.reproject(I1, D1, view, Hm)):
                    if not bits(a, b): failures.append((f"case {case}: reproject {W}x{H} frame {k}", name))
                o = view_o.fuse(I1, D1, view, Hm, prev_o)
                g = e.fuse(view, Hm, cur, prev_g)
                valid = o[0][1] > 0
                ok = bits(o[0][0], g[0][0]) and bits(o[0][1], g[0][1]) and all(bits(o[0][j][valid], g[0][j][valid]) for j in (2, 3, 4))
                ok = ok and (o[1] is None) == (g[1] is None) and (o[1] is None or bits(o[1], g[1])) and bits(o[2], g[2]) and bits(o[3], g[3])
                if not ok: failures.append((f"case {case}: fuse {W}x{H} frame {k} motion {motion:.2f}", "fusion"))
                prev_o, prev_g = o[0], g[0]
        finally:
            e.close()
    for f_ in failures[:20]:
        print("DIFF", f_, flush=True)
    print(f"fuzz_view: {cases} cases, {len(failures)} differences, {time.time() - t0:.0f} s", flush=True)
    return failures


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    run_cases(a[0] if a else 30, a[1] if len(a) > 1 else 1)
