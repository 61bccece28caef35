"""Which direction costs the end-to-end path its throughput?  host/device x in/out combinations through process_batch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200"))
import torch, elas_b200, synth
W, H, D = 1242, 375, 255
B = 256; slots = int(sys.argv[1]) if len(sys.argv) > 1 else 16
bpl = W + 15 - (W - 1) % 16
pairs = [synth.synthetic_pair(W, H, D, seed=i)[:2] for i in range(8)]
h_I = torch.zeros((B, 2, H, bpl), dtype=torch.uint8).pin_memory()
for i in range(B):
    h_I[i, 0, :, :W] = torch.from_numpy(pairs[i % 8][0]); h_I[i, 1, :, :W] = torch.from_numpy(pairs[i % 8][1])
h_D = torch.empty((B, 2, H, W), dtype=torch.float32).pin_memory()
d_I = h_I.cuda(); d_D = torch.empty((B, 2, H, W), dtype=torch.float32, device="cuda")
e = elas_b200.ElasB200(elas_b200.stereomapper(D), W, H, n_slots=slots, n_workers=int(os.environ.get("PROBE_WORKERS", "0")))
P = lambda t, k: [t[i, k].data_ptr() for i in range(B)]
for name, I, Dm in (("dev in,  dev out ", d_I, d_D), ("host in, dev out ", h_I, d_D), ("dev in,  host out", d_I, h_D), ("host in, host out", h_I, h_D)):
    for _ in range(2): e.process_batch_ptrs(P(I, 0), P(I, 1), P(Dm, 0), P(Dm, 1), bpl, device=False)
    e.host_times()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(6): e.process_batch_ptrs(P(I, 0), P(I, 1), P(Dm, 0), P(Dm, 1), bpl, device=False)
    dt = time.perf_counter() - t
    ht, n = e.host_times()
    print(f"{name} (copy path), {slots} slots: {6 * B / dt:8.1f} pairs/s;", {k: round(v, 3) for k, v in ht.items()})
e.close()
