// Host side of the narrowed right map (NarrowD2, k_postproc.cu): the checked right map crosses PCIe as int16, or as
// one byte per pixel plus one validity bit per pixel, and is widened into the caller's float map here (SSE2).
// Plain host code: included by elas_b200.cu and by tests/native/widen_test.cpp (CPU test, tests/test_host_widen.py).
#pragma once
#include <cstddef>
#include <cstdint>
#include <emmintrin.h>

namespace elasb {

constexpr int kWidenInvalid = -10;      // = kInvalid (common.cuh), elas.cpp:977-980

// int16 -> float, exact (|x| < 2^15); streaming stores: the destination is not read again by this core
void widen_i16_to_f32(const int16_t* src, float* dst, size_t n)
{
    size_t i = 0;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        for (; i + 8 <= n; i += 8) {
            const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
            const __m128i lo = _mm_srai_epi32(_mm_unpacklo_epi16(v, v), 16), hi = _mm_srai_epi32(_mm_unpackhi_epi16(v, v), 16);
            _mm_stream_ps(dst + i, _mm_cvtepi32_ps(lo));
            _mm_stream_ps(dst + i + 4, _mm_cvtepi32_ps(hi));
        }
        _mm_sfence();
    }
    for (; i < n; i++) dst[i] = (float)src[i];
}

// u8 values + one validity bit per pixel (rows of 32-bit words) -> float, invalid = -10 (NarrowD2 mode 2)
void widen_u8_mask_to_f32(const uint8_t* vals, const uint32_t* mask, int words_per_row, float* dst, int Dw, int Dh)
{
    alignas(16) static const uint32_t lane_mask[16][4] = {
        {0, 0, 0, 0}, {~0u, 0, 0, 0}, {0, ~0u, 0, 0}, {~0u, ~0u, 0, 0}, {0, 0, ~0u, 0}, {~0u, 0, ~0u, 0}, {0, ~0u, ~0u, 0}, {~0u, ~0u, ~0u, 0},
        {0, 0, 0, ~0u}, {~0u, 0, 0, ~0u}, {0, ~0u, 0, ~0u}, {~0u, ~0u, 0, ~0u}, {0, 0, ~0u, ~0u}, {~0u, 0, ~0u, ~0u}, {0, ~0u, ~0u, ~0u}, {~0u, ~0u, ~0u, ~0u}};
    const __m128 invalid = _mm_set1_ps((float)kWidenInvalid);
    const __m128i zero = _mm_setzero_si128();
    for (int v = 0; v < Dh; v++) {
        const uint8_t* src = vals + (size_t)v * Dw;
        const uint32_t* m = mask + (size_t)v * words_per_row;
        float* d = dst + (size_t)v * Dw;
        auto bit = [&](int u) { return (m[u >> 5] >> (u & 31)) & 1u; };
        int u = 0;
        for (; u < Dw && (reinterpret_cast<uintptr_t>(d + u) & 15); u++) d[u] = bit(u) ? (float)src[u] : (float)kWidenInvalid;
        for (; u + 16 <= Dw; u += 16) {
            // 16 validity bits starting at bit u of the row (they may straddle two words)
            const int w = u >> 5, sh = u & 31;
            uint32_t bits = m[w] >> sh;
            if (sh > 16) bits |= m[w + 1] << (32 - sh);          // w + 1 < words_per_row: u + 16 <= Dw lies beyond word w
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + u));
            const __m128i lo = _mm_unpacklo_epi8(b, zero), hi = _mm_unpackhi_epi8(b, zero);
            const __m128i q[4] = {_mm_unpacklo_epi16(lo, zero), _mm_unpackhi_epi16(lo, zero), _mm_unpacklo_epi16(hi, zero), _mm_unpackhi_epi16(hi, zero)};
            for (int k = 0; k < 4; k++) {
                const __m128 keep = _mm_load_ps(reinterpret_cast<const float*>(lane_mask[(bits >> (4 * k)) & 15]));
                _mm_stream_ps(d + u + 4 * k, _mm_or_ps(_mm_and_ps(keep, _mm_cvtepi32_ps(q[k])), _mm_andnot_ps(keep, invalid)));
            }
        }
        for (; u < Dw; u++) d[u] = bit(u) ? (float)src[u] : (float)kWidenInvalid;
    }
    _mm_sfence();
}

}  // namespace elasb
