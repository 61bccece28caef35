import time, torch
dev = torch.device("cuda")
for size in (1863000, 3726000, 16 << 20, 64 << 20):
    n = max(8, (256 << 20) // size)
    h = [torch.empty(size, dtype=torch.uint8).pin_memory() for _ in range(min(n, 16))]
    d = [torch.empty(size, dtype=torch.uint8, device=dev) for _ in range(min(n, 16))]
    for nstreams in (1, 4, 16):
        streams = [torch.cuda.Stream() for _ in range(nstreams)]
        for direction in ("D2H", "H2D"):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for i in range(n):
                with torch.cuda.stream(streams[i % nstreams]):
                    if direction == "D2H": h[i % len(h)].copy_(d[i % len(d)], non_blocking=True)
                    else: d[i % len(d)].copy_(h[i % len(h)], non_blocking=True)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            print(f"size {size/1e6:7.2f} MB  streams {nstreams:2d}  {direction}: {n*size/dt/1e9:6.1f} GB/s")
