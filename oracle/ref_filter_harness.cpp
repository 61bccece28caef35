// TEST INFRASTRUCTURE -- never loaded by the product path.
//
// The feature filters of libviso2's Matcher: Matcher::computeFeatures calls
//   filter::sobel5x5(I, I_du, I_dv, bpl, h); filter::blob5x5(I, I_f1, bpl, h); filter::checkerboard5x5(I, I_f2, bpl, h)
// (libviso2/src/matcher.cpp:799-801).  oracle/Makefile compiles libviso2/src/filter.cpp where it lies under
// /root/reference (unmodified, -O3 -msse3) and links this wrapper to it.
//
// The reference's row filters store 2 bytes past the end of du/dv and read 4 int16 past the end of their
// temporaries (filter.cpp:136, :188: `for (; i4 < end_input; ...)` in steps of 16): the outputs here live in
// padded scratch buffers, and du/dv[w*h-2 .. w*h) depend on that out-of-bounds read (the parity tests leave them
// out).  Elements the reference never writes (it allocates the maps with _mm_malloc, matcher.cpp:795-798) come
// back as 0.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include "filter.h"              // libviso2/src/filter.h through -I$(REFVISO)

extern "C" void ref_matcher_filters(const uint8_t* I, int32_t w, int32_t h, uint8_t* du, uint8_t* dv, int16_t* f1, int16_t* f2)
{
    const size_t n = (size_t)w * h;
    uint8_t* in = (uint8_t*)aligned_alloc(16, n + 64);
    uint8_t* pu = (uint8_t*)aligned_alloc(16, n + 64);
    uint8_t* pv = (uint8_t*)aligned_alloc(16, n + 64);
    int16_t* p1 = (int16_t*)aligned_alloc(16, 2 * n + 64);
    int16_t* p2 = (int16_t*)aligned_alloc(16, 2 * n + 64);
    memcpy(in, I, n); memset(in + n, 0, 64);
    memset(pu, 0, n + 64); memset(pv, 0, n + 64); memset(p1, 0, 2 * n + 64); memset(p2, 0, 2 * n + 64);
    filter::sobel5x5(in, pu, pv, w, h);
    filter::blob5x5(in, p1, w, h);
    filter::checkerboard5x5(in, p2, w, h);
    memcpy(du, pu, n); memcpy(dv, pv, n); memcpy(f1, p1, 2 * n); memcpy(f2, p2, 2 * n);
    free(in); free(pu); free(pv); free(p1); free(p2);
}
