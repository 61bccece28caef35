// K7: dense matching -- the kernel BASELINE.json's roofline metric is quoted on.
//
// Reference: Elas::computeDisparity (elas.cpp:960-1118) scan-converts every triangle and calls
// Elas::findMatch (elas.cpp:814-955) per covered pixel, once for the left and once for the right
// image.  Here the scan conversion has already produced a triangle-id map per image (k_diffuse_raster),
// so the work is a flat per-pixel pass.
//
// Decomposition (v7): one CTA of 256 threads per (ROW PAIR v0, v0+1; column segment of 256 pixels; frame).
// Both matching directions of a row read the SAME two descriptor rows (left image: own = desc1,
// other = desc2; right image: the reverse), so the CTA stages, once, the desc1 strip
// [x0, x1 + disp_max) and the desc2 strip [x0 - disp_max, x1) of both rows plus the candidate lists of
// the ~14 grid cells under the segment (both images): six TMA bulk copies (cp.async.bulk, contiguous
// 16 B/pixel rows) completing on one mbarrier.  While they are in flight every thread fetches, for its
// own column, the four triangle-id entries (2 rows x 2 images) and the planes behind them and turns
// them into d_plane / validity in registers -- the only global-memory latency of the kernel, hidden
// behind the TMA latency.  Thread t then runs findMatch for the pixels (x0+t, v0) and (x0+t, v0+1) of
// the left image TOGETHER, then of the right image: the two pixels lie in the same grid cell
// (grid_size even), so they share the candidate list, the list loads and the address arithmetic, and
// the two SAD chains give the instruction-level parallelism.  Every candidate SAD is one LDS.128 + 4
// VABSDIFF4.U8.ACC from shared memory.
//
// Per pixel (findMatch): candidates = the grid cell's disparities OUTSIDE the plane window in
// ascending order (cost = SAD), then the plane window d_plane-r..d_plane+r ascending
// (cost = SAD + prior if the triangle is valid); strict '<' keeps the first minimum (elas.cpp:790,805).
//
// Candidates are ranked by the key (cost << 8 | evaluation order): the minimum key is the lowest cost
// and, among equal costs, the candidate the reference evaluates first.  The inner loop carries NO
// per-candidate window test: every list entry is evaluated with its plain SAD (order = list position)
// and every window tap with SAD + prior (order = 64 + tap).  The prior is never positive
// (-log(1 + e/gamma)/beta, elas.cpp:984-992; checked at context creation), so a list entry that lies
// inside the window can only come out as the overall minimum when its window tap has the SAME cost
// (prior 0: invalid triangle).  The overall minimum is therefore exact unless it is such an entry --
// one test per pixel after the loops -- and only then the list pass is redone with the reference's
// window exclusion (elas.cpp:893).  The range test "warped column inside [2, W-2)" (elas.cpp:896-899,
// :907-910) is likewise hoisted: a warp whose columns all satisfy it for every disparity runs loops
// without it.
#include <cstdlib>

#include "common.cuh"

namespace elasb {
namespace {

constexpr int kThreads = 256;        // = pixels per column segment: thread t owns column x0 + t
constexpr int kSegW = kThreads;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// TMA bulk copy global -> shared (contiguous bytes, multiple of 16), completes on the mbarrier
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}

struct MatchArgs {
    FrameGeom g;
    int disp_max, match_texture, grid_size, max_cells, map_pitch;
    int map_tag_bits, map_tag_mask;      // triangle-id map entries are tag_bits | index; other tags = stale = uncovered
    uint32_t grid_magic;                 // floor(u / grid_size) == (u * grid_magic) >> 32 for u, grid_size < 65536
    int prior4[4];                       // prior of the plane window offsets 0..3 (elas.cpp:984-992)
    MatchBuffers b;
};

constexpr int kInitKey = (10000 << 8) | 255;     // elas.cpp:878-879: min_val = 10000, nothing found
constexpr int kPad = 6;     // strip entries before/after the addressed range: the taps of every plane window that reaches [0, disp_max]
                            // (d_plane in [-R, disp_max + R]) must be addressable around d_plane itself: kPad >= 2 R for R = 2, 3

// ---- shared-memory access by 32-bit shared address (no generic-pointer conversion in the loops) ----
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
template <int OFF>
__device__ __forceinline__ uint4 lds128_off(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ int lds_u16(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return (int)v;
}
// SAD of two 16-byte descriptors on top of `init` (the prior of a window tap, 0 otherwise): one chain of
// four VABSDIFF4.U8.ACC; the two pixels of a thread give two independent chains
__device__ __forceinline__ int sad16_from(const uint4& a, const uint4& b, int init)
{
    unsigned s = (unsigned)init;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(s) : "r"(a.x), "r"(b.x));
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(s) : "r"(a.y), "r"(b.y));
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(s) : "r"(a.z), "r"(b.z));
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(s) : "r"(a.w), "r"(b.w));
    return (int)s;
}

// rare path: a cell holding more than kGridListCap candidates is scanned from its bitmask in global memory
__device__ __noinline__ int scan_cell_bitmask(const uint32_t* __restrict__ cell, int gwords, int dlo, int dhi,
                                              int hi_ok, uint4 own, uint32_t oth_addr, int step16)
{
    // evaluation order no longer fits the key's 8 bits: keep (cost, d) with the sequential rule
    int min_val = 10000, min_d = -1;
    for (int w = 0; w < gwords; w++) {
        uint32_t m = __ldg(cell + w);
        while (m) {
            const int d = 32 * w + __ffs(m) - 1;
            m &= m - 1;
            if ((d >= dlo && d <= dhi) || d > hi_ok) continue;
            const int val = sad16_from(own, lds128(oth_addr + step16 * d), 0);
            if (val < min_val) { min_val = val; min_d = d; }
        }
    }
    return min_d < 0 ? -1 : ((min_val << 16) | min_d);
}

// rare path: the list pass exactly as the reference runs it (entries inside the plane window and entries
// whose warped column leaves [2, W-2) are skipped, elas.cpp:893-899); used when the fast pass's minimum is a
// list entry inside the window (see the header)
__device__ __noinline__ int exact_list_pass(uint32_t list, int cnt, uint4 own, uint32_t oth, int step16,
                                            int dlo, int dhi, int hi_ok)
{
    int best = kInitKey;
    for (int k = 1; k <= cnt; k++) {
        const int d = lds_u16(list + 2u * k);
        if ((d >= dlo && d <= dhi) || d > hi_ok) continue;
        best = min(best, sad16_from(own, lds128(oth + step16 * d), 0) * 256 + k);
    }
    return best;
}

// Packed per-pixel triangle state (registers): bit 0 = covered by a triangle of THIS frame, bit 1 = valid
// (elas.cpp:1072), bits 2.. = d_plane + kPlaneBias with d_plane = (int32_t)(plane_a*u + plane_b*v + plane_c)
// (elas.cpp:861, evaluated left to right with separate roundings).  d_plane is clamped to
// [-kPlaneBias, 2*kPlaneBias]: beyond [-radius-1, disp_max+radius+1] every value behaves the same (empty window).
constexpr int kPlaneBias = 16384;

__device__ __forceinline__ int pack_plane(const float4& pl, bool covered, float fu, float fv)
{
    const float dp = __fadd_rn(__fadd_rn(__fmul_rn(pl.x, fu), __fmul_rn(pl.y, fv)), pl.z);
    const int d_plane = min(max(__float2int_rz(dp), -kPlaneBias), 2 * kPlaneBias);     // cvt.rzi saturates, NaN -> 0
    // TriRaster::valid holds 2 or 0: it is bit 1 of the packed state
    return covered ? (((d_plane + kPlaneBias) << 2) | __float_as_int(pl.w) | 1) : 0;
}

// plane window of one pixel (elas.cpp:904-913, :934-943): taps d_plane-R..d_plane+R at consecutive shared
// addresses from wbase, cost = SAD + prior (the accumulation starts from the prior), order 64 + tap.
// CHECK = false: every tap is known to lie in [0, hi_ok].
template <int IMG, int RADIUS, bool CHECK>
__device__ __forceinline__ int window_taps(const uint4& own, uint32_t wbase, int d_plane, int hi_ok,
                                           int pr0, int pr1, int pr2, int pr3)
{
    constexpr int step16 = IMG ? 16 : -16;
    int best = kInitKey;
#define ELASB_TAP(K, PRIOR)                                                                                     \
    {                                                                                                            \
        const int val = sad16_from(own, lds128_off<step16 * (K)>(wbase), (PRIOR));                               \
        int key = val * 256 + (64 + (K) + RADIUS);                                                               \
        if (CHECK) key = (unsigned)(d_plane + (K)) > (unsigned)hi_ok ? kInitKey : key;                           \
        best = min(best, key);                                                                                   \
    }
    if (RADIUS >= 3) ELASB_TAP(-3, pr3)
    ELASB_TAP(-2, pr2) ELASB_TAP(-1, pr1) ELASB_TAP(0, pr0) ELASB_TAP(1, pr1) ELASB_TAP(2, pr2)
    if (RADIUS >= 3) ELASB_TAP(3, pr3)
#undef ELASB_TAP
    return best;
}

// generic plane radius (RADIUS template argument 0): priors from the table, every tap tested
template <int IMG>
__device__ __noinline__ int window_generic(const int32_t* __restrict__ prior, uint4 own, uint32_t oth, int d_plane, int radius,
                                           int dlo, int dhi, int hi_ok, bool valid)
{
    constexpr int step16 = IMG ? 16 : -16;
    int best = kInitKey;
    for (int d = dlo; d <= dhi; d++) {
        const int val = sad16_from(own, lds128(oth + step16 * d), valid ? __ldg(prior + abs(d - d_plane)) : 0);
        const int key = d > hi_ok ? kInitKey : val * 256 + (64 + d - (d_plane - radius));
        best = min(best, key);
    }
    return best;
}

struct SegCtx {
    uint32_t strip[2];      // shared address of strip k, row 0, AT COLUMN 0: + 16*column (+ row_stride for row 1)
    uint32_t row_stride;    // bytes between the two rows of a strip
    uint32_t lists;         // [2][max_cells][kGridListStride] u16, AT CELL COLUMN 0 of the grid row
};

// findMatch (elas.cpp:814-955) for the pixels (u, vA) and (u, vA+1) of image IMG; pkA / pkB = pack_plane()
// (0 = nothing to match: uncovered, outside [2, W-2), or the row does not exist).
// NPX = 1: only the first pixel exists (odd grid sizes, subsampling).
// Every thread of the warp runs through this function to the end (one warp-wide vote); lanes without an
// active pixel run the same instructions on harmless operands (an empty list, disqualified taps).
template <int IMG, int RADIUS, int NPX>
__device__ __forceinline__ void match_pair(const MatchArgs& a, const SegCtx& r, int u, int pkA, int pkB,
                                           bool warp_range_safe, int p0, int p1, int p2, int p3,
                                           float& outA, float& outB)
{
    constexpr int step16 = IMG ? 16 : -16;                        // warped column = u + step * d, 16 bytes per column
    const int radius = RADIUS ? RADIUS : a.g.plane_radius;
    const uint32_t S = r.row_stride;
    const uint32_t own_addr = r.strip[IMG] + 16u * (uint32_t)u;
    const uint32_t oth = r.strip[1 - IMG] + 16u * (uint32_t)u;    // other descriptor at disparity d: oth + step16*d

    outA = outB = (float)kInvalid;                                             // elas.cpp:977-980
    if (!__any_sync(0xffffffffu, (pkA | pkB) & 1)) return;                     // warp-uniform: nothing covered here
    const uint4 mid = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
    const uint4 ownA = lds128(own_addr);
    const uint4 ownB = NPX == 2 ? lds128(own_addr + S) : ownA;
    const bool actA = (pkA & 1) && sad16_from(ownA, mid, 0) >= a.match_texture;             // elas.cpp:851-859
    const bool actB = NPX == 2 && (pkB & 1) && sad16_from(ownB, mid, 0) >= a.match_texture;

    // the warped column stays inside [2, W-2) and d inside [0, disp_max]  <=>  0 <= d <= hi_ok
    const int hi_ok = min(a.disp_max, IMG ? a.g.W - 3 - u : u - 2);
    const int dpA = (pkA >> 2) - kPlaneBias, dpB = (pkB >> 2) - kPlaneBias;

    // (i) the grid cell's candidate list (elas.cpp:890-903, :919-932), plain SAD, both pixels per entry
    const uint32_t list = r.lists + (__umulhi((uint32_t)u, a.grid_magic) + (uint32_t)(IMG * a.max_cells)) * (kGridListStride * 2);
    const int cnt_raw = lds_u16(list);
    const bool wide = cnt_raw == 0xFFFF;
    const int cnt = (actA || actB) && !wide ? cnt_raw : 0;
    int bestA = kInitKey, bestB = kInitKey;
    if (warp_range_safe) {
#pragma unroll 1
        for (int k = 1; k <= cnt; k++) {
            const uint32_t at = oth + step16 * lds_u16(list + 2u * k);
            bestA = min(bestA, sad16_from(ownA, lds128(at), 0) * 256 + k);
            if (NPX == 2) bestB = min(bestB, sad16_from(ownB, lds128(at + S), 0) * 256 + k);
        }
    } else {
#pragma unroll 1
        for (int k = 1; k <= cnt; k++) {
            const int d = lds_u16(list + 2u * k);
            const uint32_t at = oth + step16 * d;
            const int kk = d > hi_ok ? kInitKey : k;                         // disqualified: the key lands at or above kInitKey
            bestA = min(bestA, sad16_from(ownA, lds128(at), 0) * 256 + kk);
            if (NPX == 2) bestB = min(bestB, sad16_from(ownB, lds128(at + S), 0) * 256 + kk);
        }
    }

    // (ii) the plane windows with the prior (elas.cpp:904-913, :934-943)
    int winA = kInitKey, winB = kInitKey;
    if (RADIUS) {
        // all taps of the warp's active pixels inside [0, hi_ok]: no per-tap test
        const bool tapsA = !actA || (dpA >= RADIUS && dpA + RADIUS <= hi_ok);
        const bool tapsB = !actB || (dpB >= RADIUS && dpB + RADIUS <= hi_ok);
        const int mA = (pkA & 2) ? -1 : 0, mB = (pkB & 2) ? -1 : 0;           // triangle valid: prior applies (elas.cpp:1072)
        // taps are addressed around a centre clamped into the padded strip (an inactive pixel, or a d_plane
        // far outside [0, disp_max], computes harmless taps that are all disqualified)
        const int dcA = min(max(dpA, RADIUS - kPad), a.disp_max + kPad - RADIUS);
        const int dcB = min(max(dpB, RADIUS - kPad), a.disp_max + kPad - RADIUS);
        const uint32_t wA = oth + step16 * dcA, wB = oth + S + step16 * dcB;
        if (__all_sync(0xffffffffu, tapsA && tapsB)) {
            winA = window_taps<IMG, RADIUS, false>(ownA, wA, dpA, hi_ok, p0 & mA, p1 & mA, p2 & mA, p3 & mA);
            if (NPX == 2) winB = window_taps<IMG, RADIUS, false>(ownB, wB, dpB, hi_ok, p0 & mB, p1 & mB, p2 & mB, p3 & mB);
        } else {
            winA = window_taps<IMG, RADIUS, true>(ownA, wA, dcA == dpA ? dpA : -64, hi_ok, p0 & mA, p1 & mA, p2 & mA, p3 & mA);
            if (NPX == 2) winB = window_taps<IMG, RADIUS, true>(ownB, wB, dcB == dpB ? dpB : -64, hi_ok, p0 & mB, p1 & mB, p2 & mB, p3 & mB);
        }
    }

    // (iii) per pixel: overall minimum, the exactness test of the header, decode evaluation order -> disparity
#pragma unroll
    for (int px = 0; px < NPX; px++) {
        const bool act = px ? actB : actA;
        const int d_plane = px ? dpB : dpA;
        int bl = px ? bestB : bestA;
        int bw = px ? winB : winA;
        const int wlo = d_plane - radius;
        if (!RADIUS || wide) {
            // generic plane radius / a cell with more candidates than a list holds: rare, per pixel
            if (act) {
                const uint4 own = px ? ownB : ownA;
                const uint32_t othp = px ? oth + S : oth;
                const int dlo = max(wlo, 0), dhi = min(d_plane + radius, a.disp_max);
                if (!RADIUS) bw = window_generic<IMG>(a.b.prior, own, othp, d_plane, radius, dlo, dhi, hi_ok, ((px ? pkB : pkA) & 2) != 0);
                if (wide) {
                    const int cell = (int)__umulhi((uint32_t)u, a.grid_magic);
                    const int v = (int)(a.b.rows_per_cta * blockIdx.y) + px;
                    const uint32_t* bits = a.b.grid[IMG] + (size_t)blockIdx.z * a.b.grid_stride + ((size_t)(v / a.grid_size) * a.g.gw + cell) * a.g.gwords;
                    const int wide_res = scan_cell_bitmask(bits, a.g.gwords, dlo, dhi, hi_ok, own, othp, step16);
                    // the bitmask path ran first in evaluation order: it wins ties
                    int min_d = bw < kInitKey ? wlo + (bw & 255) - 64 : -1;
                    if (wide_res >= 0 && (min_d < 0 || (wide_res >> 16) <= (bw >> 8))) min_d = wide_res & 0xFFFF;
                    (px ? outB : outA) = min_d >= 0 ? (float)min_d : -1.0f;
                    continue;
                }
            }
        }
        // the minimum is a list entry: exact unless it lies inside the plane window
        int d_list = lds_u16(list + 2u * (uint32_t)(bl & 255));              // position 255 (nothing found) is never used below
        if (bl < bw && (unsigned)(d_list - max(wlo, 0)) <= (unsigned)(min(d_plane + radius, a.disp_max) - max(wlo, 0)) &&
            min(d_plane + radius, a.disp_max) >= max(wlo, 0)) {
            bl = exact_list_pass(list, cnt, px ? ownB : ownA, px ? oth + S : oth, step16, max(wlo, 0),
                                 min(d_plane + radius, a.disp_max), hi_ok);
            d_list = lds_u16(list + 2u * (uint32_t)(bl & 255));
        }
        const int best = min(bl, bw);
        const int min_d = best >= kInitKey ? -1 : (best & 255) < 64 ? d_list : wlo + (best & 255) - 64;   // elas.cpp:947-954
        if (act) (px ? outB : outA) = (float)min_d;
    }
}

// ROWS = image rows per CTA: 2 = a thread matches (u, v0) and (u, v0+1) together (needs an even grid_size:
// both lie in one grid cell); 1 = one row per CTA (odd grid sizes; subsampling, where only even rows exist)
#ifndef K7_MINBLOCKS
#define K7_MINBLOCKS 5
#endif
template <int RADIUS, int ROWS, bool SUB>     // plane_radius (elas.cpp:993); 0 = generic
__global__ void __launch_bounds__(kThreads, K7_MINBLOCKS)
k_matching(const __grid_constant__ MatchArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;

    const FrameGeom& g = a.g;
    const MatchBuffers& b = a.b;
    const int z = blockIdx.z;
    const int v0 = SUB ? 2 * blockIdx.y : ROWS * blockIdx.y;              // elas.cpp:1085: only even rows when subsampling
    if (SUB && (v0 >> 1) >= g.Dh) return;
    const bool rowB = ROWS == 2 && v0 + 1 < g.H;
    const int x0 = blockIdx.x * kSegW, x1 = min(x0 + kSegW, g.W);
    // strip 0 holds desc1 columns from x0 - kPad, strip 1 holds desc2 columns from x0 - disp_max - kPad,
    // cap = kSegW + disp_max + 2*kPad entries per row; only the part inside the image is copied, the rest
    // is addressable garbage that is never selected (loops that could reach it test the range)
    const int cap = kSegW + a.disp_max + 2 * kPad;
    const uint32_t row_stride = (uint32_t)cap * 16u;
    // the candidate lists come first: a thread without any candidate decodes list position 255 of its cell (value
    // unused), which then falls into the strips instead of beyond the allocation
    const uint32_t lists = smem_u32(smem_raw);                             // [2][max_cells][kGridListStride]
    const uint32_t strip0 = lists + 2u * (uint32_t)a.max_cells * (kGridListStride * 2), strip1 = strip0 + ROWS * row_stride;
    const int org0 = x0 - kPad, org1 = x0 - a.disp_max - kPad;
    const int gy = (int)__umulhi((uint32_t)v0, a.grid_magic);              // v0 / grid_size, elas.cpp:867
    const int c0 = (int)__umulhi((uint32_t)x0, a.grid_magic);
    const int ncell = (int)__umulhi((uint32_t)(x1 - 1), a.grid_magic) - c0 + 1;

    // while the copies fly: this thread's four triangle-id entries -> planes -> packed (d_plane, valid, covered)
    const int u = x0 + threadIdx.x;
    // columns that are matched at all: inside the segment, inside [2, W-2) (elas.cpp:828), even when subsampling (:1079)
    const bool col = u < x1 && u >= 2 && u < g.W - 2 && !(SUB && ((u & 1) || (u >> 1) >= g.Dw));
    int pk[2][2] = {{0, 0}, {0, 0}};
    int e[2][2];
    // the triangle-id entries first: their latency (DRAM) overlaps the barrier set-up and the TMA issue
    {
        // 32-bit element indices from the group's base pointers (a group's arrays stay far below 2^31 elements)
        const uint32_t mrow = (uint32_t)z * (uint32_t)b.map_stride + (uint32_t)v0 * (uint32_t)a.map_pitch + (uint32_t)u;
#pragma unroll
        for (int row = 0; row < ROWS; row++) {
            const bool on = col && (row == 0 || rowB);
            e[0][row] = on ? __ldg(b.map[0] + (mrow + row * a.map_pitch)) : -1;
            e[1][row] = on ? __ldg(b.map[1] + (mrow + row * a.map_pitch)) : -1;
        }
    }
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        // TMA bulk copies on one mbarrier: per row the desc1 strip [x0, x1+disp_max) and the desc2 strip
        // [x0-disp_max, x1), and the two runs of candidate lists.  Descriptor rows outside [3, H-3) are all
        // zero (k_descriptor.cu), so the reference's row clamp max(min(v, H-3), 2) (elas.cpp:834) reads the
        // same bytes as row v itself.
        const int s0hi = min(x1 + a.disp_max, g.W), s1lo = max(x0 - a.disp_max, 0);
        const uint32_t b0 = (uint32_t)(s0hi - x0) * 16u, b1 = (uint32_t)(x1 - s1lo) * 16u;
        const uint32_t bl = (uint32_t)ncell * kGridListStride * 2u;
        mbar_expect_tx(&bar, (rowB ? 2u : 1u) * (b0 + b1) + 2 * bl);
        const uint4* d1 = b.desc[0] + (size_t)z * b.desc_stride + (size_t)v0 * g.W;
        const uint4* d2 = b.desc[1] + (size_t)z * b.desc_stride + (size_t)v0 * g.W;
        tma_bulk_g2s(strip0 + kPad * 16u, d1 + x0, b0, &bar);
        tma_bulk_g2s(strip1 + (uint32_t)(s1lo - org1) * 16u, d2 + s1lo, b1, &bar);
        if (rowB) {
            tma_bulk_g2s(strip0 + row_stride + kPad * 16u, d1 + g.W + x0, b0, &bar);
            tma_bulk_g2s(strip1 + row_stride + (uint32_t)(s1lo - org1) * 16u, d2 + g.W + s1lo, b1, &bar);
        }
        const size_t cell0 = ((size_t)gy * g.gw + c0) * kGridListStride;
        tma_bulk_g2s(lists, b.lists[0] + (size_t)z * b.lists_stride + cell0, bl, &bar);
        tma_bulk_g2s(lists + (uint32_t)a.max_cells * kGridListStride * 2u, b.lists[1] + (size_t)z * b.lists_stride + cell0, bl, &bar);
    }

    {
        const uint32_t tbase = (uint32_t)z * (uint32_t)b.tri_stride;
        float4 pl[2][2];
#pragma unroll
        for (int img = 0; img < 2; img++)
#pragma unroll
            for (int row = 0; row < ROWS; row++) {
                // stale entries = other frames = uncovered
                const bool covered = (e[img][row] & ~a.map_tag_mask) == a.map_tag_bits;
                e[img][row] = covered ? (int)(tbase + (uint32_t)(e[img][row] & a.map_tag_mask)) : -1;
                pl[img][row] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (covered) pl[img][row] = __ldg(reinterpret_cast<const float4*>(&b.tri[img][(uint32_t)e[img][row]].pa));
            }
        const float fu = (float)u;
#pragma unroll
        for (int img = 0; img < 2; img++)
#pragma unroll
            for (int row = 0; row < ROWS; row++)
                pk[img][row] = pack_plane(pl[img][row], e[img][row] >= 0, fu, (float)(v0 + row));
    }
    const int p0 = a.prior4[0], p1 = a.prior4[1], p2 = a.prior4[2], p3 = a.prior4[3];
    // warp-uniform: every column of this warp keeps every disparity's warped column inside [2, W-2)
    const int uw0 = x0 + (threadIdx.x & ~31), uw1 = uw0 + 31;
    const bool safe0 = uw0 - 2 >= a.disp_max, safe1 = g.W - 3 - uw1 >= a.disp_max;
    SegCtx r;
    r.row_stride = row_stride;
    r.strip[0] = strip0 - 16u * (uint32_t)org0;          // wraps modulo 2^32; + 16*column lands inside the strip
    r.strip[1] = strip1 - 16u * (uint32_t)org1;
    r.lists = lists - (uint32_t)c0 * (kGridListStride * 2);

    mbar_wait(&bar, 0);
    float o[2][2];
    match_pair<0, RADIUS, ROWS>(a, r, u, pk[0][0], pk[0][1], safe0, p0, p1, p2, p3, o[0][0], o[0][1]);
    match_pair<1, RADIUS, ROWS>(a, r, u, pk[1][0], pk[1][1], safe1, p0, p1, p2, p3, o[1][0], o[1][1]);
    if (u >= x1 || (SUB && ((u & 1) || (u >> 1) >= g.Dw))) return;
    const uint32_t at = (uint32_t)z * (uint32_t)b.D_stride + (SUB ? (uint32_t)(v0 >> 1) * g.Dw + (u >> 1) : (uint32_t)v0 * g.W + u);
#pragma unroll
    for (int img = 0; img < 2; img++) {
        b.D[img][at] = o[img][0];
        if (ROWS == 2 && rowB) b.D[img][at + g.W] = o[img][1];
    }
}

int max_cells_per_segment(int grid_size) { return (kSegW + grid_size - 1) / grid_size + 1; }

size_t smem_bytes_for(const FrameGeom& g, int grid_size, int rows)
{
    const int dmax = g.dn - 1;
    return 2 * (size_t)rows * (kSegW + dmax + 2 * kPad) * 16 + 2 * (size_t)max_cells_per_segment(grid_size) * kGridListStride * 2;
}

template <int RADIUS, int ROWS, bool SUB>
void launch_inst(dim3 grid, size_t smem, cudaStream_t s, const MatchArgs& a)
{
    static unsigned long long optin = 0;
    if (ensure_dynamic_smem(k_matching<RADIUS, ROWS, SUB>, 200 * 1024, &optin) != cudaSuccess) return;   // the error stays pending
    k_matching<RADIUS, ROWS, SUB><<<grid, kThreads, smem, s>>>(a);
}

template <int RADIUS>
void launch_radius(int rows, bool sub, dim3 grid, size_t smem, cudaStream_t s, const MatchArgs& a)
{
    if (sub)            launch_inst<RADIUS, 1, true>(grid, smem, s, a);
    else if (rows == 2) launch_inst<RADIUS, 2, false>(grid, smem, s, a);
    else                launch_inst<RADIUS, 1, false>(grid, smem, s, a);
}

// two rows per CTA needs both rows of a pair in one grid-cell row
int rows_per_cta(const elas_b200_params& p) { return (!p.subsampling && p.grid_size % 2 == 0) ? 2 : 1; }

}  // namespace

size_t matching_smem_bytes(const FrameGeom& g, const elas_b200_params& p) { return smem_bytes_for(g, p.grid_size, rows_per_cta(p)); }

void launch_matching(const FrameGeom& g, const elas_b200_params& p, const MatchBuffers& b, int n_frames,
                     int map_tag_bits, int map_tag_shift, cudaStream_t s)
{
    MatchArgs a;
    a.g = g;
    a.disp_max = p.disp_max; a.match_texture = p.match_texture; a.grid_size = p.grid_size;
    a.max_cells = max_cells_per_segment(p.grid_size);
    a.map_pitch = map_pitch(g);
    a.map_tag_bits = map_tag_bits;
    a.map_tag_mask = (1 << map_tag_shift) - 1;
    a.grid_magic = (uint32_t)(0x100000000ull / (uint32_t)p.grid_size) + 1u;
    a.b = b;
    for (int k = 0; k < 4; k++) a.prior4[k] = b.prior_host && k < g.dn ? b.prior_host[k] : 0;
    const int rows = rows_per_cta(p);
    a.b.rows_per_cta = p.subsampling ? 2 : rows;
    const int items = p.subsampling ? (g.H + 1) / 2 : (g.H + rows - 1) / rows;
    dim3 grid((g.W + kSegW - 1) / kSegW, items, n_frames);
    const size_t smem = smem_bytes_for(g, p.grid_size, rows);
    if (g.plane_radius == 2)      launch_radius<2>(rows, p.subsampling != 0, grid, smem, s, a);      // ROBOTICS
    else if (g.plane_radius == 3) launch_radius<3>(rows, p.subsampling != 0, grid, smem, s, a);      // MIDDLEBURY
    else                          launch_radius<0>(rows, p.subsampling != 0, grid, smem, s, a);
    count_launch();
}

}  // namespace elasb
