"""Per-stage device times of one frame (CUDA events on the slot stream) + host-stage wall time.
Usage: python tools/stage_times.py [W H DMAX] [reps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200"))
import elas_b200  # noqa: E402
import synth  # noqa: E402


def main():
    W, H, DMAX = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (1242, 375, 255)
    reps = int(sys.argv[4]) if len(sys.argv) >= 5 else 20
    L, R, _ = synth.synthetic_pair(W, H, DMAX, 0)
    p = elas_b200.stereomapper(DMAX)
    e = elas_b200.ElasB200(p, W, H, n_slots=1)
    for _ in range(3):
        e.process(L, R)
    e.set_timing(True)
    acc = {}
    wall = []
    for _ in range(reps):
        t0 = time.perf_counter()
        e.process(L, R)
        wall.append(time.perf_counter() - t0)
        for name, ms in e.stage_times():
            acc.setdefault(name, []).append(ms)
    print(f"{W}x{H} d_max {DMAX}: single slot, wall per frame best {min(wall)*1e3:.3f} ms, median {np.median(wall)*1e3:.3f} ms")
    total = 0.0
    for name, v in acc.items():
        print(f"  {name:14s} {np.median(v)*1e3:9.1f} us   (min {min(v)*1e3:8.1f})")
        total += np.median(v)
    print(f"  {'sum':14s} {total*1e3:9.1f} us")
    for flush in (True, False):
        print(f"  k_matching isolated ({'L2 flushed' if flush else 'warm L2'}): {e.time_matching(50, flush)*1e3:.2f} us/launch")
    # host stage alone
    e.set_timing(False)
    e.process(L, R, capture=True)
    dcan = e.stage("dcan_raw").reshape(-1)
    lat = e.stage("lattice_dims")
    t = []
    for _ in range(10):
        t0 = time.perf_counter()
        elas_b200.host_stage(p, W, H, dcan.reshape(lat[1], lat[0]))
        t.append(time.perf_counter() - t0)
    print(f"  host stage via ABI (incl. ctypes/alloc overhead): best {min(t)*1e3:.3f} ms")
    e.close()


if __name__ == "__main__":
    main()
