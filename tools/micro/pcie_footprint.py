"""D2H throughput of 1.86 MB copies as a function of the pinned destination footprint (IOMMU/TLB reach?)."""
import time, torch
dev = torch.device("cuda")
src = [torch.empty(1863000, dtype=torch.uint8, device=dev) for _ in range(16)]
streams = [torch.cuda.Stream() for _ in range(16)]
for nbuf in (64, 256, 1024):
    big = torch.empty((nbuf, 1863000), dtype=torch.uint8).pin_memory()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 2048
        for i in range(n):
            with torch.cuda.stream(streams[i % 16]):
                big[i % nbuf].copy_(src[i % 16], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{nbuf:5d} pinned buffers ({nbuf * 1.863e-3:6.2f} GB): D2H {n * 1.863e6 / dt / 1e9:5.1f} GB/s")
    del big
