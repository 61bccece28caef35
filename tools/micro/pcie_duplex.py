"""D2H throughput of frame-sized copies with and without concurrent H2D image copies and a busy GPU."""
import time, torch
dev = torch.device("cuda")
out = [torch.empty(1863000, dtype=torch.uint8).pin_memory() for _ in range(64)]
src = [torch.empty(1863000, dtype=torch.uint8, device=dev) for _ in range(64)]
img_h = [torch.empty(468000, dtype=torch.uint8).pin_memory() for _ in range(64)]
img_d = [torch.empty(468000, dtype=torch.uint8, device=dev) for _ in range(64)]
a = torch.randn(4096, 4096, device=dev); 
streams = [torch.cuda.Stream() for _ in range(16)]
up = [torch.cuda.Stream() for _ in range(16)]
def run(h2d, busy, n=2048):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        with torch.cuda.stream(streams[i % 16]):
            out[i % 64].copy_(src[i % 64], non_blocking=True)
        if h2d and i % 2 == 0:
            with torch.cuda.stream(up[i % 16]):
                img_d[i % 64].copy_(img_h[i % 64], non_blocking=True)
                img_d[(i + 1) % 64].copy_(img_h[(i + 1) % 64], non_blocking=True)
        if busy and i % 8 == 0:
            (a @ a)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"h2d={h2d} busy={busy}: D2H {n * 1.863e6 / dt / 1e9:5.1f} GB/s  ({n / 2 / dt:7.0f} frames/s worth of maps)")
for h2d in (0, 1):
    for busy in (0, 1):
        run(h2d, busy)
