// Issue-rate probe: warp instructions per clock per SM for VABSDIFF4, HADD2 (plain / |x| operand), IADD3,
// and a 50/50 VABSDIFF4 + HADD2 mix.  Each thread runs 8 independent chains.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned* out, int iters, unsigned seed)
{
    unsigned a[8], b = seed + threadIdx.x;
    __half2 h[8], hb = __floats2half2_rn((float)(threadIdx.x & 7), 3.f);
#pragma unroll
    for (int j = 0; j < 8; j++) { a[j] = seed * (j + 1) + threadIdx.x; h[j] = __floats2half2_rn((float)j, (float)(j + 1)); }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (MODE == 0) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[j]) : "r"(b), "r"(a[(j + 1) & 7]));
            if (MODE == 1) h[j] = __hadd2(h[j], hb);
            if (MODE == 2) h[j] = __hadd2(__habs2(h[j]), hb);
            if (MODE == 3) a[j] = a[j] + b + (unsigned)i;
            if (MODE == 4) {
                if (j & 1) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[j]) : "r"(b), "r"(a[(j + 1) & 7]));
                else { h[j] = __hadd2(__habs2(h[j]), hb); h[j + 1] = __hsub2(h[j + 1], hb); }
            }
            if (MODE == 5) { unsigned t; asm volatile("prmt.b32 %0, %1, %2, 0x7140;" : "=r"(t) : "r"(a[j]), "r"(b)); a[j] = t; }
        }
    }
    unsigned r = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) r += a[j] + (unsigned)__half2float(__low2half(h[j])) + (unsigned)__half2float(__high2half(h[j]));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, int per_iter)
{
    unsigned* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 16, 1);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, iters, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double warp_inst = 148.0 * 8 * 8 * iters * per_iter;     // 8 warps per CTA
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %.3f ms  %.2f warp-inst/clk/SM  (%.1f lanes/clk/SM)\n", name, ms, warp_inst / cycles / 148, 32 * warp_inst / cycles / 148);
    cudaFree(d);
}
int main()
{
    run<0>("VABSDIFF4.ACC", 8);
    run<1>("HADD2", 8);
    run<2>("HADD2 |a|", 8);
    run<3>("IADD3", 8);
    run<4>("4 VABSDIFF4 + 8 HADD2 (mix)", 12);
    run<5>("PRMT", 8);
}
