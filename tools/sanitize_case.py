"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200"))
import numpy as np
import elas_b200, synth
for name, p, (W, H, D) in (("stereomapper", elas_b200.stereomapper(63), (320, 120, 63)),
                           ("middlebury", elas_b200.middlebury().copy(disp_max=63), (320, 120, 63)),
                           ("subsampling", elas_b200.stereomapper(95).copy(subsampling=1), (416, 200, 95)),
                           ("odd width", elas_b200.demo(63), (333, 131, 63))):
    L, R, _ = synth.synthetic_pair(W, H, D, 1)
    e = elas_b200.ElasB200(p, W, H, n_slots=3, n_workers=2)
    rc, D1, D2 = e.process(L, R)
    st, B1, B2 = e.process_batch([L] * 5, [R] * 5)
    if not p.subsampling:
        e.colormap(); e.reproject((700.0, W / 2, H / 2, 0.5, 30.0, 1.2), np.hstack([np.eye(3), np.zeros((3, 1))]))
    e.close()
    same = all(np.array_equal(b.view(np.uint32), D1.view(np.uint32)) for b in B1)
    print(name, "rc", rc, "valid", int((D1 >= 0).sum()), "batch==single", same, flush=True)
