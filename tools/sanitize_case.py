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

# map fusion (short and long lists) and the Matcher feature filters
sys.path.insert(0, os.path.join(ROOT, "tests"))
from view_cases import fusion_sequence
for name in ("tiny", "backwards"):
    seq = fusion_sequence(name, frames=2)
    h, w = seq[0][1].shape
    e = elas_b200.ElasB200(elas_b200.stereomapper(63), w, h, n_slots=1)
    prev = None
    for I1, D1, view, H in seq:
        cur = e.reproject(view, H, I1=I1, D1=D1)
        fused = e.fuse(view, H, cur, prev)
        prev = fused[0]
    e.close()
    print("fusion", name, "points", len(fused[2]), len(fused[3]), flush=True)
w, h = 96, 64
e = elas_b200.ElasB200(elas_b200.stereomapper(63), w, h, n_slots=1)
view = np.array([300.0, 48.0, 32.0, 0.54, 30.0, 1.2], np.float32)
Hm = np.hstack([np.eye(3), np.zeros((3, 1))])
base = list(e.reproject(view, Hm, I1=np.full((h, w), 9, np.uint8), D1=np.full((h, w), 20.0, np.float32)))
prev = [a.copy() for a in base]
for k in (2, 3, 4):
    prev[k][:] = base[k][30, 40]           # every previous point on one pixel: the long-list path
fused = e.fuse(view, Hm, base, prev)
e.close()
print("fusion one-pixel", len(fused[2]), len(fused[3]), flush=True)
rng = np.random.default_rng(0)
for shape in ((6, 16), (53, 112), (375, 1248)):
    out = elas_b200.matcher_filters(rng.integers(0, 256, shape, dtype=np.uint8))
    print("filters", shape, int(out[2].astype(np.int64).sum()), flush=True)
