// SM-driven device -> pinned-host writes (zero-copy) vs copy-engine D2H, alone and with a memory-bound kernel running.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_copy_out(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void k_copy_out4(const float* __restrict__ src, float* __restrict__ dst, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void k_busy(uint4* p, size_t n, int reps)
{
    for (int r = 0; r < reps; r++)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            uint4 v = p[i]; v.x += r; p[i] = v;
        }
}
int main()
{
    const size_t bytes = 1863000 / 16 * 16, n = bytes / 16;
    const int NB = 64, COPIES = 1024;
    uint4 *d_src, *d_busy; cudaMalloc(&d_src, bytes); cudaMalloc(&d_busy, (size_t)1 << 30);
    uint4* h[NB]; for (int i = 0; i < NB; i++) cudaHostAlloc(&h[i], bytes, cudaHostAllocMapped);
    cudaStream_t st[16], busy; for (auto& s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&busy, cudaStreamNonBlocking);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int load = 0; load < 2; load++)
        for (int mode = 0; mode < 4; mode++) {
            cudaDeviceSynchronize();
            if (load) k_busy<<<148 * 4, 256, 0, busy>>>(d_busy, ((size_t)1 << 30) / 16, 150);
            cudaEventRecord(e0, st[0]);
            for (int i = 0; i < COPIES; i++) {
                cudaStream_t s = st[i % 16];
                if (mode == 0) cudaMemcpyAsync(h[i % NB], d_src, bytes, cudaMemcpyDeviceToHost, s);
                if (mode == 1) k_copy_out<<<32, 256, 0, s>>>(d_src, h[i % NB], n);
                if (mode == 2) k_copy_out<<<148, 256, 0, s>>>(d_src, h[i % NB], n);
                if (mode == 3) k_copy_out4<<<148, 256, 0, s>>>((const float*)d_src, (float*)h[i % NB], n * 4);
            }
            for (int i = 1; i < 16; i++) { cudaEventRecord(e1, st[i]); cudaStreamWaitEvent(st[0], e1, 0); }
            cudaEventRecord(e1, st[0]); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const char* names[4] = {"copy engine", "SM stores, 32 CTAs", "SM stores, 148 CTAs", "SM 4-byte stores"};
            printf("%s, %-20s: %5.1f GB/s\n", load ? "busy GPU" : "idle GPU", names[mode], COPIES * (double)bytes / ms / 1e6);
            cudaDeviceSynchronize();
        }
}
