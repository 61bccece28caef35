"""Where does a pipelined batch spend its time?  Host-phase wall times per frame + PCIe bandwidth."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200"))
import numpy as np, torch
import elas_b200, synth

W, H, D = 1242, 375, 255
B = int(os.environ.get("PROBE_B", "64"))
slots = int(sys.argv[1]) if len(sys.argv) > 1 else 16
bpl = W + 15 - (W - 1) % 16
pairs = [synth.synthetic_pair(W, H, D, seed=i)[:2] for i in range(8)]
h_I = torch.zeros((B, 2, H, bpl), dtype=torch.uint8).pin_memory()
for i in range(B):
    h_I[i, 0, :, :W] = torch.from_numpy(pairs[i % 8][0]); h_I[i, 1, :, :W] = torch.from_numpy(pairs[i % 8][1])
h_D = torch.empty((B, 2, H, W), dtype=torch.float32).pin_memory()
d_I = h_I.cuda(); d_D = torch.empty((B, 2, H, W), dtype=torch.float32, device="cuda")
# PCIe
big_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); big_d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: big_d.copy_(big_h, non_blocking=True)), ("D2H", lambda: big_h.copy_(big_d, non_blocking=True))):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(4): fn()
    torch.cuda.synchronize(); print(f"PCIe {name}: {4 * 0.268435456 / (time.perf_counter() - t):.1f} GB/s")
e = elas_b200.ElasB200(elas_b200.stereomapper(D), W, H, n_slots=slots, n_workers=int(os.environ.get("PROBE_WORKERS", "0")))
P = lambda t, k: [t[i, k].data_ptr() for i in range(B)]
for dev, I, Dm in ((True, d_I, d_D), (False, h_I, h_D)):
    for _ in range(3): e.process_batch_ptrs(P(I, 0), P(I, 1), P(Dm, 0), P(Dm, 1), bpl, device=dev)
    e.host_times()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): e.process_batch_ptrs(P(I, 0), P(I, 1), P(Dm, 0), P(Dm, 1), bpl, device=dev)
    dt = time.perf_counter() - t
    ht, n = e.host_times()
    print(f"{'device' if dev else 'host  '} buffers, {slots} slots: {10 * B / dt:8.1f} pairs/s; per-frame host wall (ms):", {k: round(v, 3) for k, v in ht.items()}, "sum", round(sum(ht.values()), 3))
e.close()
