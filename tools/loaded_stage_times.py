"""Per-stage elapsed times of launch chains WHILE the pipeline is loaded (CUDA events on the groups' streams):
shows where chains queue.  Usage: python tools/loaded_stage_times.py [workers groups frames_per_group]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200"))
import numpy as np, torch
import elas_b200, synth
W, H, D, B = 1242, 375, 255, 512
workers, groups, fpg = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (4, 12, 8)
bpl = W + 15 - (W - 1) % 16
pairs = [synth.synthetic_pair(W, H, D, seed=i)[:2] for i in range(8)]
h_I = torch.zeros((B, 2, H, bpl), dtype=torch.uint8)
for i in range(B):
    h_I[i, 0, :, :W] = torch.from_numpy(pairs[i % 8][0]); h_I[i, 1, :, :W] = torch.from_numpy(pairs[i % 8][1])
d_I = h_I.cuda(); d_D = torch.empty((B, 2, H, W), dtype=torch.float32, device="cuda")
ptr = lambda t, k: [t[i, k].data_ptr() for i in range(B)]
dev = (ptr(d_I, 0), ptr(d_I, 1), ptr(d_D, 0), ptr(d_D, 1))
e = elas_b200.ElasB200(elas_b200.stereomapper(D), W, H, n_slots=groups, n_workers=workers, frames_per_group=fpg)
for _ in range(2):
    e.process_batch_ptrs(*dev, bpl, device=True)
e.set_timing(True)
acc = {}
for _ in range(3):
    e.process_batch_ptrs(*dev, bpl, device=True)
    for g in range(groups):
        for name, ms in e.stage_times(slot=g):
            acc.setdefault(name, []).append(ms)
tot = 0
for name, v in acc.items():
    print(f"  {name:16s} median {np.median(v)*1e3:8.1f} us   max {max(v)*1e3:8.1f}")
    tot += np.median(v)
print(f"  chain median sum {tot*1e3:.1f} us for {fpg} frames")
e.close()
