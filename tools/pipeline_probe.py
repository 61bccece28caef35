"""Throughput of the device-resident batch path for a few (workers, groups, frames per group) settings, with the
host-side time split per frame.  Usage: python tools/pipeline_probe.py [W H DMAX] [B]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200"))
import torch
import elas_b200, synth

W, H, D = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (1242, 375, 255)
B = int(sys.argv[4]) if len(sys.argv) >= 5 else 512
bpl = W + 15 - (W - 1) % 16
pairs = [synth.synthetic_pair(W, H, D, seed=i)[:2] for i in range(8)]
h_I = torch.zeros((B, 2, H, bpl), dtype=torch.uint8).pin_memory()
for i in range(B):
    h_I[i, 0, :, :W] = torch.from_numpy(pairs[i % 8][0]); h_I[i, 1, :, :W] = torch.from_numpy(pairs[i % 8][1])
d_I = h_I.cuda(); d_D = torch.empty((B, 2, H, W), dtype=torch.float32, device="cuda")
h_D = torch.empty((B, 2, H, W), dtype=torch.float32).pin_memory()
ptr = lambda t, k: [t[i, k].data_ptr() for i in range(B)]
dev = (ptr(d_I, 0), ptr(d_I, 1), ptr(d_D, 0), ptr(d_D, 1)); host = (ptr(h_I, 0), ptr(h_I, 1), ptr(h_D, 0), ptr(h_D, 1))
settings = [(1, 3, 8), (2, 6, 8), (4, 12, 8), (6, 18, 8), (4, 12, 4), (4, 16, 2), (8, 24, 1), (2, 4, 8), (3, 6, 8)]
if os.environ.get("PROBE"):
    settings = [tuple(int(x) for x in s.split(",")) for s in os.environ["PROBE"].split(";")]
for workers, groups, fpg in settings:
    e = elas_b200.ElasB200(elas_b200.stereomapper(D), W, H, n_slots=groups, n_workers=workers, frames_per_group=fpg)
    for path, bufs, isdev in (("device", dev, True), ("host", host, False)):
        for _ in range(2):
            e.process_batch_ptrs(*bufs, bpl, device=isdev)
        e.host_times()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            e.process_batch_ptrs(*bufs, bpl, device=isdev)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        ht, n = e.host_times()
        print(f"workers {workers} groups {groups} x{fpg} {path:6s}: {B / dt:8.0f} pairs/s   per frame us: " +
              " ".join(f"{k}={v * 1e3:.1f}" for k, v in ht.items() if k != "unused"), flush=True)
    e.close()
