"""D2H throughput of frame-sized copies while the SMs are busy (memory-bound and compute-bound kernels on other streams)."""
import time, torch
dev = torch.device("cuda")
out = [torch.empty(1863000, dtype=torch.uint8).pin_memory() for _ in range(64)]
src = [torch.empty(1863000, dtype=torch.uint8, device=dev) for _ in range(64)]
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
streams = [torch.cuda.Stream() for _ in range(16)]
work = torch.cuda.Stream()
def run(kind, n=2048):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        with torch.cuda.stream(streams[i % 16]):
            out[i % 64].copy_(src[i % 64], non_blocking=True)
        if i % 16 == 0:
            with torch.cuda.stream(work):
                if kind == "matmul": a @ a
                if kind == "memset": big.fill_(i & 255)
    for s in streams: s.synchronize()
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"{kind:8s}: D2H {n * 1.863e6 / dt / 1e9:5.1f} GB/s")
for kind in ("idle", "matmul", "memset", "idle"):
    run(kind)
