"""Stress of the frame-group scheduler: the scenario of tests/test_gpu_parity.py::test_frame_groups_equal_single_frames
repeated with varying worker / group counts; reports every frame whose maps or status differ from the single-frame
results.  Usage: python tools/stress_groups.py [iterations]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200"))
import numpy as np
import elas_b200, synth
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
p = elas_b200.stereomapper(95)
pairs = [synth.synthetic_pair(416, 200, 95, s)[:2] for s in (41, 42, 43, 44, 45)]
blank = (np.full((200, 416), 90, np.uint8), np.full((200, 416), 90, np.uint8))
order = [0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 1, 0, 5, 2, 3, 4, 0, 1]
frames = [pairs[i] if i < 5 else blank for i in order]
bad = 0
want = None
for it in range(iters):
    slots, workers, fpg = [(3, 2, 4), (2, 1, 4), (4, 3, 2), (6, 4, 1), (3, 3, 8)][it % 5]
    e = elas_b200.ElasB200(p, 416, 200, n_slots=slots, n_workers=workers, frames_per_group=fpg)
    if want is None:
        want = [e.process(L, R)[1:] for L, R in pairs]
    for rep in range(3):
        status, D1, D2 = e.process_batch([a for a, _ in frames], [b for _, b in frames])
        for i, k in enumerate(order):
            if k == 5:
                ok = status[i] == elas_b200.E_FEW_SUPPORT and (D1[i] == -10).all() and (D2[i] == -10).all()
            else:
                ok = status[i] == 0 and np.array_equal(D1[i].view(np.uint32), want[k][0].view(np.uint32)) and \
                     np.array_equal(D2[i].view(np.uint32), want[k][1].view(np.uint32))
            if not ok:
                bad += 1
                n1 = int((D1[i].view(np.uint32) != (want[k][0] if k < 5 else np.full_like(D1[i], -10)).view(np.uint32)).sum())
                n2 = int((D2[i].view(np.uint32) != (want[k][1] if k < 5 else np.full_like(D2[i], -10)).view(np.uint32)).sum())
                print(f"iter {it} rep {rep} cfg {(slots, workers, fpg)} frame {i} (kind {k}): status {status[i]}, D1 differs at {n1}, D2 at {n2}", flush=True)
    e.close()
print("stress done:", iters, "iterations,", bad, "bad frames", flush=True)
