// the feature filters of libviso2's Matcher -- filter::sobel5x5, filter::blob5x5 and
// filter::checkerboard5x5 (libviso2/src/filter.cpp:474-530) as Matcher::computeFeatures calls them on every image
// (matcher.cpp:799-801) -- as one fused kernel: the image is read once and the four feature maps (du, dv u8; blob,
// checkerboard int16) are written once; the reference's four int16 temporaries and its int32 integral image never exist.
//
// The reference walks the image as ONE flat array of n = w*h elements (w = bytes per line, a multiple of 16):
//   column pass, rows 2..h-3, zero elsewhere (filter.cpp:291-349, :353-391)
//     tv = i(-2) + 4 i(-1) + 6 i(0) + 4 i(+1) + i(+2)      th = i(-2) + 2 i(-1) - 2 i(+1) - i(+2)      tc = i(-2) + i(-1) - i(+1) - i(+2)
//   row pass over FLAT positions (a pixel next to a row end takes taps from the neighbouring row; :124-173, :180-222, :395-417)
//     du[o] = sat_u8(((tv[o-2] + 2 tv[o-1] - 2 tv[o+1] - tv[o+2]) >> 7) + 128)             2 <= o < n
//     dv[o] = sat_u8(((th[o-2] + 4 th[o-1] + 6 th[o] + 4 th[o+1] + th[o+2]) >> 7) + 128)   2 <= o < n
//     f2[o] = tc[o-2] + tc[o-1] - tc[o+1] - tc[o+2]                                         2 <= o < n - 6
//   blob (:507-530) from a 2-d integral image, again over flat positions p:
//     f1[p+3+3w] = (int16)(-(I[p+5+5w] - I[p+5] - I[p+5w] + I[p]) + 2 (I[p+4+4w] - I[p+4+w] - I[p+1+4w] + I[p+1+w]) + 7 in[p+3+3w])
//   which is the 5x5 mask (-1 ring, +1 ring, 8 centre) wherever the 5x5 box does not straddle a row end: columns
//   3..w-3.  For the other five columns per row the flat walk combines integral-image entries of two different rows:
//   those outputs are large row-sum differences truncated to int16.  They are reproduced from row sums by
//   extra warps of the same launch (blob_wrapped_row), in the reference's wrapping int32 arithmetic.
// Elements the reference leaves unwritten (it allocates the maps uninitialised, matcher.cpp:795-798) are 0 here.  The
// reference reads tv/th[n .. n+3] -- beyond its buffers -- for du/dv[n-2], [n-1]; here those temporaries count as 0.
#include "common.cuh"

namespace elasb {
namespace {

constexpr int kFiltThreads = 256, kFiltPer = 4, kFiltSpan = kFiltThreads * kFiltPer;      // flat outputs per CTA
constexpr int kFiltRowWords = kFiltSpan / 4 + 2;                                          // [p0 - 4, p0 + span + 4)

__device__ __forceinline__ int byte_of(const uint32_t (&w)[3], int i) { return (w[i >> 2] >> (8 * (i & 3))) & 255; }
__device__ __forceinline__ int sat_u8(int x) { return min(max(x, 0), 255); }

__device__ __forceinline__ void blob_wrapped_row(const uint8_t* __restrict__ in, int w, int h, int row, int lane, int16_t* __restrict__ f1);

// blocks [0, main_blocks): kFiltSpan flat outputs each; blocks behind them: the row-straddling blob positions, one warp per row
__global__ void __launch_bounds__(kFiltThreads)
k_matcher_filters(const uint8_t* __restrict__ in, int w, int h, int main_blocks, uint8_t* __restrict__ du, uint8_t* __restrict__ dv,
                  int16_t* __restrict__ f1, int16_t* __restrict__ f2)
{
    __shared__ uint32_t rows[5][kFiltRowWords];
    if ((int)blockIdx.x >= main_blocks) {
        const int row = ((int)blockIdx.x - main_blocks) * (kFiltThreads / 32) + (threadIdx.x >> 5);
        if (row < h) blob_wrapped_row(in, w, h, row, threadIdx.x & 31, f1);
        return;
    }
    const long long n = (long long)w * h;
    const long long p0 = (long long)blockIdx.x * kFiltSpan;
    // image rows -2..+2 around the CTA's flat range, as aligned words (w and p0 are multiples of 4); 0 outside the image
    for (int i = threadIdx.x; i < 5 * kFiltRowWords; i += kFiltThreads) {
        const int k = i / kFiltRowWords, j = i - k * kFiltRowWords;
        const long long q = p0 - 4 + 4ll * j + (long long)(k - 2) * w;
        rows[k][j] = (q >= 0 && q + 4 <= n) ? __ldg(reinterpret_cast<const uint32_t*>(in + q)) : 0u;
    }
    __syncthreads();
    const long long p = p0 + kFiltPer * threadIdx.x;          // this thread's outputs: flat p .. p+3
    if (p >= n) return;
    // bytes p-4 .. p+7 of the five rows
    uint32_t r[5][3];
#pragma unroll
    for (int k = 0; k < 5; k++)
#pragma unroll
        for (int j = 0; j < 3; j++) r[k][j] = rows[k][threadIdx.x + j];
    // column pass at flat q = p-2 .. p+5 (byte index 2..9); temporaries are 0 outside image rows [2, h-2) and beyond n
    const long long qa = p - 2;
    const int row_a = qa >= 0 ? (int)(qa / w) : -1;
    const long long next_row_at = (long long)(row_a + 1) * w;         // first flat position of the next row
    int tv[8], th[8], tc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const long long q = qa + i;
        const int row = q >= next_row_at ? row_a + 1 : row_a;
        const bool live = row >= 2 && row < h - 2;
        const int a = byte_of(r[0], i + 2), b = byte_of(r[1], i + 2), c = byte_of(r[2], i + 2), d = byte_of(r[3], i + 2), e = byte_of(r[4], i + 2);
        tv[i] = live ? a + 4 * b + 6 * c + 4 * d + e : 0;
        th[i] = live ? a + 2 * b - 2 * d - e : 0;
        tc[i] = live ? a + b - d - e : 0;
    }
    uint32_t pu = 0, pv = 0;
    int16_t c2[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const long long o = p + k;
        const int su = sat_u8(((tv[k] + 2 * tv[k + 1] - 2 * tv[k + 3] - tv[k + 4]) >> 7) + 128);
        const int sv = sat_u8(((th[k] + 4 * th[k + 1] + 6 * th[k + 2] + 4 * th[k + 3] + th[k + 4]) >> 7) + 128);
        const bool on = o >= 2;
        pu |= (uint32_t)(on ? su : 0) << (8 * k);
        pv |= (uint32_t)(on ? sv : 0) << (8 * k);
        c2[k] = (on && o < n - 6) ? (int16_t)(tc[k] + tc[k + 1] - tc[k + 3] - tc[k + 4]) : (int16_t)0;
    }
    *reinterpret_cast<uint32_t*>(du + p) = pu;
    *reinterpret_cast<uint32_t*>(dv + p) = pv;
    *reinterpret_cast<uint2*>(f2 + p) = make_uint2((uint16_t)c2[0] | ((uint32_t)(uint16_t)c2[1] << 16), (uint16_t)c2[2] | ((uint32_t)(uint16_t)c2[3] << 16));
    // blob: 5x5 and 3x3 box sums around o for the columns whose box stays inside the row
    const int row_p = (int)(p / w), col_p = (int)(p - (long long)row_p * w);      // p..p+3 lie in one row (w % 4 == 0)
    int s5[4] = {0, 0, 0, 0}, s3[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 5; k++) {
        int b[8];
#pragma unroll
        for (int i = 0; i < 8; i++) b[i] = byte_of(r[k], i + 2);          // columns o-2 .. o+5 of this row
#pragma unroll
        for (int o = 0; o < 4; o++) {
            s5[o] += b[o] + b[o + 1] + b[o + 2] + b[o + 3] + b[o + 4];
            if (k >= 1 && k <= 3) s3[o] += b[o + 1] + b[o + 2] + b[o + 3];
        }
    }
    const bool rows_ok = row_p >= 3 && row_p <= h - 3;
    uint32_t packed[2] = {0u, 0u};
    bool wrapped = false;
#pragma unroll
    for (int o = 0; o < 4; o++) {
        const int c = col_p + o;
        const long long at = p + o;
        const bool in_range = at >= 3 + 3ll * w && at < n - 2 - 2ll * w;      // the positions the reference writes at all
        const bool clean = rows_ok && c >= 3 && c <= w - 3;
        wrapped |= in_range && !clean;
        const int v = clean ? -s5[o] + 2 * s3[o] + 7 * byte_of(r[2], o + 4) : 0;
        packed[o >> 1] |= (uint32_t)(uint16_t)(int16_t)v << (16 * (o & 1));
    }
    if (!wrapped) {
        *reinterpret_cast<uint2*>(f1 + p) = make_uint2(packed[0], packed[1]);
    } else {
        // a group that holds row-straddling positions: those belong to blob_wrapped_row
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const int c = col_p + o;
            const long long at = p + o;
            const bool in_range = at >= 3 + 3ll * w && at < n - 2 - 2ll * w;
            const bool clean = rows_ok && c >= 3 && c <= w - 3;
            if (clean || !in_range) f1[at] = (int16_t)(packed[o >> 1] >> (16 * (o & 1)));
        }
    }
}

// R(x, y): sum of row y from column 0 to x inclusive, for x within 8 of either row end (uint32, wrapping);
// total = the sum of the whole row
__device__ __forceinline__ uint32_t row_prefix(const uint8_t* __restrict__ row, int w, int x, uint32_t total)
{
    uint32_t s;
    if (x < 8) { s = 0; for (int i = 0; i <= x; i++) s += row[i]; }
    else { s = total; for (int i = x + 1; i < w; i++) s -= row[i]; }
    return s;
}

// The five row-straddling blob positions of image row `row`: columns w-2, w-1, 0, 1, 2 (filter.cpp:511-527 on flat
// indices).  One warp: all lanes sum the seven rows row-3 .. row+3 (the integral-image entries of a straddling
// position are prefixes that reach almost to the end of those rows), lanes 0..4 then evaluate one position each.
__device__ __forceinline__ void blob_wrapped_row(const uint8_t* __restrict__ in, int w, int h, int row, int lane, int16_t* __restrict__ f1)
{
    uint32_t total[7];
#pragma unroll
    for (int k = 0; k < 7; k++) {
        const int y = row - 3 + k;
        uint32_t s = 0;
        if (y >= 0 && y < h) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (size_t)y * w);
            for (int j = lane; j < w / 4; j += 32) s += __dp4a(__ldg(src + j), 0x01010101u, 0u);
        }
        total[k] = __reduce_add_sync(0xffffffffu, s);
    }
    if (lane >= 5) return;
    const int c = lane < 2 ? w - 2 + lane : lane - 2;
    const long long n = (long long)w * h, o = (long long)row * w + c;
    if (o < 3 + 3ll * w || o >= n - 2 - 2ll * w) return;
    const long long p = o - 3 - 3ll * w;
    // I[q + hi*w] - I[q + lo*w] for a flat position q: both entries share q's column, so the difference is the sum of
    // that column's row prefixes over the rows (row(q) + lo, row(q) + hi]
    auto integral_diff = [&](long long q, int lo, int hi) {
        const int y = (int)(q / w), x = (int)(q - (long long)y * w);
        uint32_t s = 0;
        for (int r = y + lo + 1; r <= y + hi; r++) {
            uint32_t t = 0;
#pragma unroll
            for (int k = 0; k < 7; k++) t = (r - (row - 3) == k) ? total[k] : t;        // registers, not a local array
            s += row_prefix(in + (size_t)r * w, w, x, t);
        }
        return s;
    };
    uint32_t res = 0u - (integral_diff(p + 5, 0, 5) - integral_diff(p, 0, 5));
    res += 2u * (integral_diff(p + 4, 1, 4) - integral_diff(p + 1, 1, 4));
    res += 7u * in[o];
    f1[o] = (int16_t)(uint16_t)res;
}

}  // namespace

// in: w*h bytes (w % 16 == 0, h >= 6), all pointers 16-byte aligned device memory
void launch_matcher_filters(const uint8_t* in, int w, int h, uint8_t* du, uint8_t* dv, int16_t* f1, int16_t* f2, cudaStream_t s)
{
    const long long n = (long long)w * h;
    const int main_blocks = (int)((n + kFiltSpan - 1) / kFiltSpan), wrap_blocks = (h + kFiltThreads / 32 - 1) / (kFiltThreads / 32);
    k_matcher_filters<<<main_blocks + wrap_blocks, kFiltThreads, 0, s>>>(in, w, h, main_blocks, du, dv, f1, f2);
    count_launch();
}

}  // namespace elasb
