"""One K-size sequence through the D1 consumers (colour map, back-projection, fusion) and the Matcher filters, for an
ncu launch list:  ncu --metrics gpu__time_duration.sum --csv --log-file out.csv python tools/view_case.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200"))
import numpy as np
import elas_b200, synth
W, H, D = 1242, 375, 255
L, R, _ = synth.synthetic_pair(W, H, D, 0)
e = elas_b200.ElasB200(elas_b200.stereomapper(D), W, H, n_slots=1)
view = np.array([721.5377, 609.5593, 172.854, 0.5371657, 30.0, 1.2], np.float32)
prev = None
for k in range(3):
    rc, D1, _ = e.process(L, R)
    e.colormap()
    Hm = np.hstack([np.eye(3), [[0.02 * k], [0.0], [0.35 * k]]])
    cur = e.reproject(view, Hm)
    fused = e.fuse(view, Hm, cur, prev)
    prev = fused[0]
    print("frame", k, "points kept / current:", len(fused[2]), len(fused[3]), flush=True)
e.close()
I = np.zeros((H, 1248), np.uint8); I[:, :W] = L
for _ in range(3):
    elas_b200.matcher_filters(I)
