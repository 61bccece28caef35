"""K7 timing as the bench measures it (one launch per frame group, CUDA events, L2 flushed) at K (8 frames per launch)
and 4096x2160 (1 frame), plus a checksum of the final maps so that kernel variants can be compared for identical
results.  Usage: [ELAS_B200_LIB=path/libelas_b200.so] python tools/k7_group_time.py [K|HD|4K|both]"""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200"))
import numpy as np
import elas_b200, synth
def bmatch(w, h, dmax, gs=20):
    return 72 * w * h + 8 * (-(-w // gs)) * (-(-h // gs)) * (dmax + 2)
which = sys.argv[1] if len(sys.argv) > 1 else "both"
for tag, (W, H, D) in (("K", (1242, 375, 255)), ("HD", (1920, 1080, 128)), ("4K", (4096, 2160, 256))):
    if which not in (tag, "both"):
        continue
    e = elas_b200.ElasB200(elas_b200.stereomapper(D), W, H, n_slots=1, frames_per_group=0)
    nf = e.frames_per_group
    pairs = [synth.synthetic_pair(W, H, D, seed=s)[:2] for s in range(min(nf, 4))]
    Ls = [pairs[i % len(pairs)][0] for i in range(nf)]; Rs = [pairs[i % len(pairs)][1] for i in range(nf)]
    status, D1s, D2s = e.process_batch(Ls, Rs)
    crc = 0
    for a in D1s + D2s:
        crc = zlib.crc32(a.tobytes(), crc)
    best = min(e.time_matching(iters=30, flush_l2=True, per_frame=False)[0] for _ in range(3))
    b = bmatch(W, H, D) * nf
    print(f"{tag}: {nf} frames/launch {best*1e3:.2f} us  {b/best/1e6:.0f} GB/s  frac {b/best/1e6/6552.3:.4f}  crc {crc:08x}", flush=True)
    e.close()
