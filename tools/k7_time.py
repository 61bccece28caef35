"""K7 timing at the three BASELINE geometries (CUDA events, L2 flushed): python tools/k7_time.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200"))
import elas_b200, synth
def bmatch(w, h, dmax, gs=20):
    return 72 * w * h + 8 * (-(-w // gs)) * (-(-h // gs)) * (dmax + 2)
for (W, H, D) in ((1242, 375, 255), (1920, 1080, 128), (4096, 2160, 256)):
    L, R, _ = synth.synthetic_pair(W, H, D, 0)
    e = elas_b200.ElasB200(elas_b200.stereomapper(D), W, H, n_slots=1)
    e.process(L, R)
    ms = e.time_matching(30, True); warm = e.time_matching(30, False)
    b = bmatch(W, H, D)
    print(f"{W}x{H} d{D}: {ms*1e3:.2f} us flushed ({b/ms/1e6:.0f} GB/s, frac {b/ms/1e6/6552.3:.3f}), {warm*1e3:.2f} us warm", flush=True)
    e.close()
