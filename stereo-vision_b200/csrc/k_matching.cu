// K7: dense matching -- the kernel BASELINE.json's roofline metric is quoted on.
//
// Reference: Elas::computeDisparity (elas.cpp:960-1118) scan-converts every triangle and calls
// Elas::findMatch (elas.cpp:814-955) per covered pixel, once for the left and once for the right
// image.  Here the scan conversion has already produced a triangle-id map per image (k_raster), so
// the work is a flat per-pixel pass.
//
// Decomposition: one CTA per (image row v, column segment of ~420 pixels).  Both matching
// directions of a row read the SAME two descriptor rows (left image: own = desc1, other = desc2;
// right image: the reverse), so the CTA stages the desc1 strip and the desc2 strip of row
// clamp(v,2,H-3) (elas.cpp:834) in shared memory ONCE -- two TMA bulk copies (cp.async.bulk,
// contiguous 16 B/pixel rows) completing on one mbarrier -- and then produces the D1 and the D2
// pixels of the segment from shared memory.  While the copies are in flight the warps turn the
// candidate-grid bitmasks of the ~22 cells under the segment into short ascending disparity lists in
// shared memory.  Every candidate SAD is then an LDS.128 + 4 VABSDIFF4.  HBM sees each descriptor
// byte about once per row (neighbouring segments overlap by disp_max columns, absorbed by L2), the
// triangle-id maps once and the two output rows once.
//
// Per pixel (findMatch): candidates = the grid cell's disparities OUTSIDE the plane window in
// ascending order (cost = SAD), then the plane window d_plane-r..d_plane+r ascending
// (cost = SAD + prior if the triangle is valid); strict '<' keeps the first minimum (elas.cpp:790,805).
#include <cstdlib>

#include "common.cuh"

namespace elasb {
namespace {

constexpr int kSegTargetDefault = 320;  // rows wider than this are cut into ~equal segments (sweep: tools/k7_sweep*.py)

struct SegPlan { int nseg, segw; };

inline int env_int(const char* name, int dflt)
{
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

inline SegPlan plan_segments(int W)
{
    static const int kSegTarget = env_int("ELAS_B200_K7_SEG", kSegTargetDefault);
    SegPlan s;
    s.nseg = (W + kSegTarget - 1) / kSegTarget;
    s.segw = ((W + s.nseg - 1) / s.nseg + 31) & ~31;
    s.nseg = (W + s.segw - 1) / s.segw;
    return s;
}

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
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
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
    int disp_max, match_texture, grid_size, subsampling;
    int segw, max_cells, map_pitch, variant;
    int map_tag_bits, map_tag_mask;      // triangle-id map entries are tag_bits | index; other tags = stale = uncovered
    uint32_t grid_magic;                 // floor(u / grid_size) == (u * grid_magic) >> 32 for u, grid_size < 65536
    const uint4* desc[2];
    const TriRaster* tri[2];
    const int32_t* map[2];
    const uint32_t* grid[2];
    const uint16_t* lists[2];
    const int32_t* prior;
    float* D[2];
};

// Candidates are ranked by the key (cost << 8 | evaluation order): the minimum key is the lowest cost
// and, among equal costs, the candidate the reference evaluates first -- its strict '<' (elas.cpp:790,
// :805) -- so candidates need no sequential compare-and-select chain.  A candidate the reference
// skips (inside the plane window during the grid pass, or warped column outside [2, W-2),
// elas.cpp:896-899, :907-910) gets the initial key instead of a branch.  Costs stay below 2^15.
constexpr int kInitKey = (10000 << 8) | 255;
constexpr int kPad = 4;     // strip entries before/after the addressed range: window taps may step outside [0, disp_max]

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
__device__ __forceinline__ int lds_s32(uint32_t addr)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");   // ordered after plane_pass()'s stores
    return v;
}
// one accumulation chain: the candidate loops run two candidates per trip, which is the ILP
__device__ __forceinline__ int sad16_chain(const uint4& a, const uint4& b)
{
    unsigned s = 0;
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
            const int val = sad16_chain(own, lds128(oth_addr + step16 * d));
            if (val < min_val) { min_val = val; min_d = d; }
        }
    }
    return min_d < 0 ? -1 : ((min_val << 16) | min_d);
}

// Shared-memory image of one row segment (32-bit shared addresses)
struct RowCtx {
    uint32_t strip[2];      // strip[k] + 16*(column - org[k]) = descriptor k at that column
    int org[2];
    uint32_t lists;         // [2][max_cells][kGridListStride] u16
    uint32_t tmap;          // [2][segw] i32: triangle-id map entries, replaced in place by packed (d_plane, valid, covered)
    int x0, n, v, c0, gy;
};

// Turns the triangle-id entries of the row segment (both images) into what findMatch needs from the
// triangle: d_plane = (int32_t)(plane_a*u + plane_b*v + plane_c) (elas.cpp:861, evaluated left to right
// with separate roundings) and the triangle's validity flag (elas.cpp:1072).  Each thread fetches the
// planes of all its pixels together (independent 16-byte loads in flight at once, the only global-memory
// latency of the kernel after the TMA prologue) and writes the packed result back over the entry:
//   bit 0 = covered by a triangle of THIS frame, bit 1 = valid, bits 2.. = d_plane + kPlaneBias.
// d_plane is clamped to [-kPlaneBias, 2*kPlaneBias]: beyond [-radius-1, disp_max+radius+1] every value
// behaves the same (empty plane window).  A thread reads back only entries it wrote: no block barrier.
constexpr int kPlaneBias = 16384;
constexpr int kPlaneBatch = 4;

template <int kThreads>
__device__ __forceinline__ void plane_pass(const MatchArgs& a, int32_t* tmap, int n, int x0, int v)
{
    const float fv = (float)v;
#pragma unroll
    for (int img = 0; img < 2; img++) {
        const TriRaster* __restrict__ tris = a.tri[img];
        int32_t* row = tmap + img * a.segw;
        for (int i0 = threadIdx.x; i0 < n; i0 += kPlaneBatch * kThreads) {
            float4 pl[kPlaneBatch];
            bool covered[kPlaneBatch];
#pragma unroll
            for (int j = 0; j < kPlaneBatch; j++) {
                const int i = i0 + j * kThreads;
                const int e = i < n ? row[i] : -1;
                covered[j] = (e & ~a.map_tag_mask) == a.map_tag_bits;     // stale entries = other frames = uncovered
                pl[j] = covered[j] ? __ldg(reinterpret_cast<const float4*>(&tris[e & a.map_tag_mask].pa))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < kPlaneBatch; j++) {
                const int i = i0 + j * kThreads;
                if (i >= n) continue;
                const float fu = (float)(x0 + i);
                const float dp = __fadd_rn(__fadd_rn(__fmul_rn(pl[j].x, fu), __fmul_rn(pl[j].y, fv)), pl[j].z);
                const int d_plane = __float2int_rz(fminf(fmaxf(dp, (float)-kPlaneBias), (float)(2 * kPlaneBias)));
                const int valid = __float_as_int(pl[j].w) != 0;
                row[i] = covered[j] ? (((d_plane + kPlaneBias) << 2) | (valid << 1) | 1) : 0;
            }
        }
    }
}

// findMatch (elas.cpp:814-955) for the pixels of one image in this row segment
template <int IMG, int RADIUS, int kThreads, bool SUB>
__device__ __forceinline__ void match_row(const MatchArgs& a, const RowCtx& r, int p0, int p1, int p2, int p3)
{
    const FrameGeom& g = a.g;
    constexpr int step16 = IMG ? 16 : -16;                        // warped column = u + step * d, 16 bytes per column
    const int radius = RADIUS ? RADIUS : g.plane_radius;
    float* __restrict__ Drow = a.D[IMG] + (SUB ? (size_t)(r.v >> 1) * g.Dw : (size_t)r.v * g.W);
    const uint32_t own_base = r.strip[IMG] - 16u * (uint32_t)r.org[IMG];          // + 16*u
    const uint32_t oth_base = r.strip[1 - IMG] - 16u * (uint32_t)r.org[1 - IMG];  // + 16*u + step16*d
    const uint32_t tmap = r.tmap + (uint32_t)IMG * a.segw * 4u;
    const uint32_t lists = r.lists + (uint32_t)IMG * a.max_cells * (kGridListStride * 2);

    for (int i = threadIdx.x; i < r.n; i += kThreads) {
        const int u = r.x0 + i;
        if (SUB && ((u & 1) || (u >> 1) >= g.Dw)) continue;                // elas.cpp:1079
        const int e = lds_s32(tmap + 4u * i);                              // packed by plane_pass()
        float out = (float)kInvalid;                                       // elas.cpp:977-980
        if ((e & 1) && u >= 2 && u < g.W - 2) {                            // covered by a triangle; elas.cpp:828
            const uint4 own = lds128(own_base + 16u * u);
            const uint4 mid = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
            if (sad16_chain(own, mid) >= a.match_texture) {                // elas.cpp:851-859
                const bool valid = (e & 2) != 0;
                const int d_plane = (e >> 2) - kPlaneBias;
                const int dlo = max(d_plane - radius, 0);
                const int dhi = min(d_plane + radius, a.disp_max);
                const uint32_t oth = oth_base + 16u * u;                   // other descriptor at disparity d: oth + step16*d
                // the warped column stays inside [2, W-2) and d inside [0, disp_max]  <=>  0 <= d <= hi_ok
                const int hi_ok = min(a.disp_max, IMG ? g.W - 3 - u : u - 2);

                int best = kInitKey;                                       // elas.cpp:878-879
                // (i) grid candidates outside the plane window, ascending (elas.cpp:890-903, :919-932)
                const int c = (int)__umulhi((uint32_t)u, a.grid_magic) - r.c0;        // u / grid_size - c0
                const uint32_t list = lists + (uint32_t)c * (kGridListStride * 2);
                const int cnt = lds_u16(list);
                int wide = -1;
                if (cnt != 0xFFFF) {
                    const unsigned span = dhi >= dlo ? (unsigned)(dhi - dlo) : 0u;
                    const int wlo = dhi >= dlo ? dlo : -1 - a.disp_max;    // empty window: nothing matches
                    // two entries per trip; the entry after the last is a sentinel (disp_max + 1 > hi_ok)
#pragma unroll 1
                    for (int k = 1; k <= cnt; k += 2) {
                        const int da = lds_u16(list + 2u * k), db = lds_u16(list + 2u * k + 2u);
                        const int va = sad16_chain(own, lds128(oth + step16 * da));
                        const int vb = sad16_chain(own, lds128(oth + step16 * db));
                        const bool bad_a = (da > hi_ok) | ((unsigned)(da - wlo) <= span);
                        const bool bad_b = (db > hi_ok) | ((unsigned)(db - wlo) <= span);
                        const int ka = bad_a ? kInitKey : va * 256 + k;
                        const int kb = bad_b ? kInitKey : vb * 256 + k + 1;
                        best = min(best, min(ka, kb));
                    }
                } else {
                    wide = scan_cell_bitmask(a.grid[IMG] + ((size_t)r.gy * g.gw + r.c0 + c) * g.gwords, g.gwords,
                                             dlo, dhi, hi_ok, own, oth, step16);
                }
                // (ii) the plane window with the prior (elas.cpp:904-913, :934-943)
                if (RADIUS) {
                    // taps d_plane-R .. d_plane+R sit at consecutive addresses; a d_plane far outside
                    // [0, disp_max] is clamped for addressing only (every tap is then disqualified)
                    const int dc = min(max(d_plane, RADIUS - kPad), a.disp_max + kPad - RADIUS);
                    const uint32_t wbase = oth + step16 * dc;
                    const int pr0 = valid ? p0 : 0, pr1 = valid ? p1 : 0, pr2 = valid ? p2 : 0, pr3 = valid ? p3 : 0;
#define ELASB_TAP(K, PRIOR)                                                                                   \
                    {                                                                                          \
                        const int val = sad16_chain(own, lds128_off<step16 * (K)>(wbase)) + (PRIOR);           \
                        const int key = (unsigned)(d_plane + (K)) > (unsigned)hi_ok ? kInitKey                 \
                                                                                    : val * 256 + (64 + (K) + RADIUS); \
                        best = min(best, key);                                                                 \
                    }
                    if (RADIUS >= 3) ELASB_TAP(-3, pr3)
                    ELASB_TAP(-2, pr2) ELASB_TAP(-1, pr1) ELASB_TAP(0, pr0) ELASB_TAP(1, pr1) ELASB_TAP(2, pr2)
                    if (RADIUS >= 3) ELASB_TAP(3, pr3)
#undef ELASB_TAP
                } else {
                    for (int d = dlo; d <= dhi; d++) {
                        const int val = sad16_chain(own, lds128(oth + step16 * d)) + (valid ? __ldg(a.prior + abs(d - d_plane)) : 0);
                        const int key = d > hi_ok ? kInitKey : val * 256 + (64 + d - (d_plane - radius));
                        best = min(best, key);
                    }
                }
                // decode: evaluation order -> disparity
                int min_d = -1;
                if (best < kInitKey) {
                    const int ord = best & 255;
                    min_d = ord < 64 ? lds_u16(list + 2u * ord) : d_plane - radius + (ord - 64);
                }
                if (wide >= 0) {
                    // the bitmask path ran first in evaluation order: it wins ties
                    const int wval = wide >> 16, wd = wide & 0xFFFF;
                    if (min_d < 0 || wval <= (best >> 8)) min_d = wd;
                }
                out = min_d >= 0 ? (float)min_d : -1.0f;                   // elas.cpp:947-954
            }
        }
        Drow[SUB ? (u >> 1) : u] = out;
    }
}

template <int RADIUS, int kThreads, bool SUB>     // plane_radius (elas.cpp:993); 0 = generic
__global__ void __launch_bounds__(kThreads)
k_matching(const __grid_constant__ MatchArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;

    const FrameGeom& g = a.g;
    const int v = blockIdx.y;
    if (SUB && ((v & 1) || (v >> 1) >= g.Dh)) return;                     // elas.cpp:1085
    const int x0 = blockIdx.x * a.segw, x1 = min(x0 + a.segw, g.W);
    // strip 0 holds desc1 columns from x0 - kPad, strip 1 holds desc2 columns from x0 - disp_max - kPad,
    // cap = segw + disp_max + 2*kPad entries each; only the part inside the image is copied, the rest
    // is addressable garbage that is never selected (its candidates are disqualified)
    const int cap = a.segw + a.disp_max + 2 * kPad;
    uint4* strip0 = reinterpret_cast<uint4*>(smem_raw);
    uint4* strip1 = strip0 + cap;
    uint16_t* lists = reinterpret_cast<uint16_t*>(strip1 + cap);           // [2][max_cells][kGridListStride]
    int32_t* tmap = reinterpret_cast<int32_t*>(lists + 2 * a.max_cells * kGridListStride);   // [2][segw]
    RowCtx r;
    r.org[0] = x0 - kPad; r.org[1] = x0 - a.disp_max - kPad;
    r.strip[0] = smem_u32(strip0); r.strip[1] = smem_u32(strip1);
    r.lists = smem_u32(lists); r.tmap = smem_u32(tmap);
    r.x0 = x0; r.n = x1 - x0; r.v = v;
    r.gy = v / a.grid_size;                                                // elas.cpp:867
    r.c0 = x0 / a.grid_size;
    const int ncell = (x1 - 1) / a.grid_size - r.c0 + 1;

    const int vrow = max(min(v, g.H - 3), 2);                              // elas.cpp:834
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        // six TMA bulk copies on one mbarrier: two descriptor strips, two runs of candidate lists, two
        // triangle-id row segments
        const int s0hi = min(x1 + a.disp_max, g.W), s1lo = max(x0 - a.disp_max, 0);
        const uint32_t b0 = (uint32_t)(s0hi - x0) * 16u, b1 = (uint32_t)(x1 - s1lo) * 16u;
        const uint32_t bl = (uint32_t)ncell * kGridListStride * 2u;
        const uint32_t bm = (uint32_t)((r.n + 3) & ~3) * 4u;
        mbar_expect_tx(&bar, b0 + b1 + 2 * bl + 2 * bm);
        // the triangle-id rows first: the plane gather below starts from them
        tma_bulk_g2s(tmap, a.map[0] + (size_t)v * a.map_pitch + x0, bm, &bar);
        tma_bulk_g2s(tmap + a.segw, a.map[1] + (size_t)v * a.map_pitch + x0, bm, &bar);
        tma_bulk_g2s(strip0 + kPad, a.desc[0] + (size_t)vrow * g.W + x0, b0, &bar);
        tma_bulk_g2s(strip1 + (s1lo - r.org[1]), a.desc[1] + (size_t)vrow * g.W + s1lo, b1, &bar);
        const size_t cell0 = ((size_t)r.gy * g.gw + r.c0) * kGridListStride;
        tma_bulk_g2s(lists, a.lists[0] + cell0, bl, &bar);
        tma_bulk_g2s(lists + a.max_cells * kGridListStride, a.lists[1] + cell0, bl, &bar);
    }
    // prior of the plane window offsets 0..3 (elas.cpp:984-992), in registers
    const int p0 = __ldg(a.prior), p1 = g.dn > 1 ? __ldg(a.prior + 1) : 0, p2 = g.dn > 2 ? __ldg(a.prior + 2) : 0,
              p3 = g.dn > 3 ? __ldg(a.prior + 3) : 0;
    mbar_wait(&bar, 0);
    plane_pass<kThreads>(a, tmap, r.n, x0, v);
    match_row<0, RADIUS, kThreads, SUB>(a, r, p0, p1, p2, p3);
    match_row<1, RADIUS, kThreads, SUB>(a, r, p0, p1, p2, p3);
}

template <int RADIUS, int THREADS, bool SUB>
void launch_sub(dim3 grid, size_t smem, cudaStream_t s, const MatchArgs& a)
{
    static unsigned long long optin = 0;
    if (ensure_dynamic_smem(k_matching<RADIUS, THREADS, SUB>, 200 * 1024, &optin) != cudaSuccess) return;   // launch error stays pending
    k_matching<RADIUS, THREADS, SUB><<<grid, THREADS, smem, s>>>(a);
}

template <int RADIUS, int THREADS>
void launch_one(dim3 grid, size_t smem, cudaStream_t s, const MatchArgs& a)
{
    if (a.subsampling) launch_sub<RADIUS, THREADS, true>(grid, smem, s, a);
    else               launch_sub<RADIUS, THREADS, false>(grid, smem, s, a);
}

template <int THREADS>
void launch_radius(int radius, dim3 grid, size_t smem, cudaStream_t s, const MatchArgs& a)
{
    if (radius == 2)      launch_one<2, THREADS>(grid, smem, s, a);      // ROBOTICS
    else if (radius == 3) launch_one<3, THREADS>(grid, smem, s, a);      // MIDDLEBURY
    else                  launch_one<0, THREADS>(grid, smem, s, a);
}

void launch_variant(int radius, int threads, dim3 grid, size_t smem, cudaStream_t s, const MatchArgs& a)
{
    if (threads == 256)      launch_radius<256>(radius, grid, smem, s, a);
    else if (threads == 64)  launch_radius<64>(radius, grid, smem, s, a);
    else                     launch_radius<128>(radius, grid, smem, s, a);
}

int max_cells_per_segment(const FrameGeom& g, int grid_size, int segw) { return (segw + grid_size - 1) / grid_size + 1; }

size_t smem_bytes_for(const FrameGeom& g, int grid_size)
{
    const SegPlan s = plan_segments(g.W);
    const int dmax = g.dn - 1;
    const size_t strip = (size_t)((s.segw + dmax) < g.W ? (s.segw + dmax) : g.W);
    const size_t cells = (size_t)max_cells_per_segment(g, grid_size, s.segw);
    (void)strip;
    return 2 * (size_t)(s.segw + dmax + 2 * kPad) * 16 + 2 * cells * kGridListStride * 2 + 2 * (size_t)s.segw * 4;
}

}  // namespace

size_t matching_smem_bytes(const FrameGeom& g, int grid_size) { return smem_bytes_for(g, grid_size); }

void launch_matching(const FrameGeom& g, const elas_b200_params& p, const uint4* desc1,
                     const uint4* desc2, const TriRaster* tri1, const TriRaster* tri2,
                     const int32_t* map1, const int32_t* map2, const uint32_t* grid1,
                     const uint32_t* grid2, const uint16_t* lists1, const uint16_t* lists2,
                     const int32_t* prior, float* D1, float* D2, int map_tag_bits, int map_tag_shift,
                     cudaStream_t s)
{
    const SegPlan sp = plan_segments(g.W);
    MatchArgs a;
    a.g = g;
    a.disp_max = p.disp_max; a.match_texture = p.match_texture; a.grid_size = p.grid_size;
    a.subsampling = p.subsampling;
    a.segw = sp.segw;
    a.max_cells = max_cells_per_segment(g, p.grid_size, sp.segw);
    a.map_pitch = map_pitch(g);
    a.map_tag_bits = map_tag_bits;
    a.map_tag_mask = (1 << map_tag_shift) - 1;
    static const int variant = env_int("ELAS_B200_K7_VARIANT", 1);
    a.variant = variant;
    a.grid_magic = (uint32_t)(0x100000000ull / (uint32_t)p.grid_size) + 1u;
    a.desc[0] = desc1; a.desc[1] = desc2;
    a.tri[0] = tri1; a.tri[1] = tri2;
    a.map[0] = map1; a.map[1] = map2;
    a.grid[0] = grid1; a.grid[1] = grid2;
    a.lists[0] = lists1; a.lists[1] = lists2;
    a.prior = prior;
    a.D[0] = D1; a.D[1] = D2;
    dim3 grid(sp.nseg, g.H, 1);
    const size_t smem = smem_bytes_for(g, p.grid_size);
    static const int threads = env_int("ELAS_B200_K7_THREADS", 128);
    launch_variant(g.plane_radius, threads, grid, smem, s, a);
    count_launch();
}

}  // namespace elasb
