// K7: dense matching -- the kernel BASELINE.json's roofline metric is quoted on.
//
// Reference: Elas::computeDisparity (elas.cpp:960-1118) scan-converts every triangle and calls
// Elas::findMatch (elas.cpp:814-955) per covered pixel, once for the left and once for the right
// image.  Here the scan conversion has already produced a triangle-id map per image (k_raster), so
// the work is a flat per-pixel pass.
//
// Decomposition: one CTA per (image row v, column segment).  Both matching directions of a row
// read the SAME two descriptor rows (left image: own = desc1, other = desc2; right image: the
// reverse), so the CTA stages the desc1 strip and the desc2 strip of row clamp(v,2,H-3)
// (elas.cpp:834) in shared memory ONCE -- two TMA bulk copies (cp.async.bulk, contiguous
// 16 B/pixel rows) completing on one mbarrier -- and then produces the D1 and the D2 pixels of the
// segment from shared memory.  Every candidate SAD is an LDS.128 + 4 VABSDIFF4.  HBM sees each
// descriptor byte about once per row (segments overlap by disp_max columns), the triangle-id maps
// once and the two output rows once.
//
// Per pixel (findMatch): candidates = the grid cell's disparities OUTSIDE the plane window in
// ascending order (cost = SAD), then the plane window d_plane-r..d_plane+r ascending
// (cost = SAD + prior if the triangle is valid); strict '<' keeps the first minimum (elas.cpp:790,805).
// The grid cell is a bitmask (see k_grid_raster.cu): ascending order = ascending set bits.
#include "common.cuh"

namespace elasb {
namespace {

constexpr int kThreads = 256;
constexpr int kSingleSegMax = 1600;     // rows up to this width are one segment
constexpr int kSegTarget = 1024;        // wider rows are cut into ~equal segments of about this size

struct SegPlan { int nseg, segw; };

__host__ __device__ inline SegPlan plan_segments(int W)
{
    SegPlan s;
    if (W <= kSingleSegMax) { s.nseg = 1; s.segw = W; return s; }
    s.nseg = (W + kSegTarget - 1) / kSegTarget;
    s.segw = ((W + s.nseg - 1) / s.nseg + 31) & ~31;
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
    int nseg, segw;
    const uint4* desc[2];
    const TriRaster* tri[2];
    const int32_t* map[2];
    const uint32_t* grid[2];
    const int32_t* prior;
    float* D[2];
};

__global__ void __launch_bounds__(kThreads)
k_matching(const __grid_constant__ MatchArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;

    const FrameGeom& g = a.g;
    const int v = blockIdx.y;
    if (a.subsampling && ((v & 1) || (v >> 1) >= g.Dh)) return;           // elas.cpp:1085
    const int x0 = blockIdx.x * a.segw, x1 = min(x0 + a.segw, g.W);
    // strip 0 = desc1 columns [s0lo, s0hi), strip 1 = desc2 columns [s1lo, s1hi)
    const int s0lo = x0, s0hi = min(x1 + a.disp_max, g.W);
    const int s1lo = max(x0 - a.disp_max, 0), s1hi = x1;
    uint4* strip0 = reinterpret_cast<uint4*>(smem_raw);
    uint4* strip1 = strip0 + (s0hi - s0lo);

    const int vrow = max(min(v, g.H - 3), 2);                              // elas.cpp:834
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t b0 = (uint32_t)(s0hi - s0lo) * 16u, b1 = (uint32_t)(s1hi - s1lo) * 16u;
        mbar_expect_tx(&bar, b0 + b1);
        tma_bulk_g2s(strip0, a.desc[0] + (size_t)vrow * g.W + s0lo, b0, &bar);
        tma_bulk_g2s(strip1, a.desc[1] + (size_t)vrow * g.W + s1lo, b1, &bar);
    }

    const int n = x1 - x0;
    const int gy = v / a.grid_size;                                        // elas.cpp:867
    const int window = 2;
    bool waited = false;

    // items 0..n-1: left image pixels, n..2n-1: right image pixels
    for (int item = threadIdx.x; item < 2 * n; item += kThreads) {
        const int img = item >= n;
        const int u = x0 + (img ? item - n : item);
        if (a.subsampling && ((u & 1) || (u >> 1) >= g.Dw)) continue;      // elas.cpp:1079
        const int t = __ldg(a.map[img] + (size_t)v * g.W + u);             // issued before the wait
        if (!waited) { mbar_wait(&bar, 0); waited = true; }

        float out = (float)kInvalid;                                       // elas.cpp:977-980
        if (t >= 0 && u >= window && u < g.W - window) {                   // elas.cpp:828
            const uint4* own_strip = img ? strip1 : strip0;
            const uint4* oth_strip = img ? strip0 : strip1;
            const int own_lo = img ? s1lo : s0lo, oth_lo = img ? s0lo : s1lo;
            const uint4 own = own_strip[u - own_lo];
            if ((int)texture16(own) >= a.match_texture) {                  // elas.cpp:851-859
                const TriRaster* tr = a.tri[img] + t;
                const float pa = __ldg(&tr->pa), pb = __ldg(&tr->pb), pc = __ldg(&tr->pc);
                const int valid = __ldg(&tr->valid);
                // elas.cpp:861: (int32_t)(plane_a*u + plane_b*v + plane_c), evaluated left to right
                const int d_plane = __float2int_rz(
                    __fadd_rn(__fadd_rn(__fmul_rn(pa, (float)u), __fmul_rn(pb, (float)v)), pc));
                const int dlo = max(d_plane - g.plane_radius, 0);
                const int dhi = min(d_plane + g.plane_radius, g.dn - 1);
                const uint32_t* cell = a.grid[img] + ((size_t)gy * g.gw + u / a.grid_size) * g.gwords;

                int min_val = 10000, min_d = -1;                           // elas.cpp:878-879
                // (i) grid candidates outside the plane window, ascending (elas.cpp:890-903, :919-932)
                for (int w = 0; w < g.gwords; w++) {
                    uint32_t m = __ldg(cell + w);
                    // clear bits d in [dlo, dhi]
                    const int lo = dlo - 32 * w, hi = dhi - 32 * w;
                    if (hi >= 0 && lo < 32 && lo <= hi) {
                        const uint32_t upto_hi = hi >= 31 ? 0xffffffffu : ((2u << hi) - 1u);
                        const uint32_t below_lo = lo <= 0 ? 0u : ((1u << lo) - 1u);
                        m &= ~(upto_hi & ~below_lo);
                    }
                    while (m) {
                        const int d = 32 * w + __ffs(m) - 1;
                        m &= m - 1;
                        const int uw = img ? u + d : u - d;
                        if (uw < window || uw >= g.W - window) continue;
                        const int val = (int)sad16(own, oth_strip[uw - oth_lo]);
                        if (val < min_val) { min_val = val; min_d = d; }
                    }
                }
                // (ii) the plane window with the prior (elas.cpp:904-913, :934-943)
                for (int d = dlo; d <= dhi; d++) {
                    const int uw = img ? u + d : u - d;
                    if (uw < window || uw >= g.W - window) continue;
                    int val = (int)sad16(own, oth_strip[uw - oth_lo]);
                    if (valid) val += __ldg(a.prior + abs(d - d_plane));
                    if (val < min_val) { min_val = val; min_d = d; }
                }
                out = min_d >= 0 ? (float)min_d : -1.0f;                   // elas.cpp:947-954
            }
        }
        const size_t addr = a.subsampling ? (size_t)(v >> 1) * g.Dw + (u >> 1) : (size_t)v * g.W + u;
        a.D[img][addr] = out;
    }
    if (!waited) mbar_wait(&bar, 0);     // never leave with a bulk copy in flight
}

}  // namespace

size_t matching_smem_bytes(const FrameGeom& g)
{
    SegPlan s = plan_segments(g.W);
    const int dmax = g.dn - 1;
    size_t l0 = (size_t)min(s.segw + dmax, g.W), l1 = (size_t)min(s.segw + dmax, g.W);
    return (l0 + l1) * 16;
}

void launch_matching(const FrameGeom& g, const elas_b200_params& p, const uint4* desc1,
                     const uint4* desc2, const TriRaster* tri1, const TriRaster* tri2,
                     const int32_t* map1, const int32_t* map2, const uint32_t* grid1,
                     const uint32_t* grid2, const int32_t* prior, float* D1, float* D2, cudaStream_t s)
{
    static bool attr_set = false;
    const size_t smem = matching_smem_bytes(g);
    if (!attr_set || smem > 48 * 1024) {
        cudaFuncSetAttribute(k_matching, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    SegPlan sp = plan_segments(g.W);
    MatchArgs a;
    a.g = g;
    a.disp_max = p.disp_max; a.match_texture = p.match_texture; a.grid_size = p.grid_size;
    a.subsampling = p.subsampling;
    a.nseg = sp.nseg; a.segw = sp.segw;
    a.desc[0] = desc1; a.desc[1] = desc2;
    a.tri[0] = tri1; a.tri[1] = tri2;
    a.map[0] = map1; a.map[1] = map2;
    a.grid[0] = grid1; a.grid[1] = grid2;
    a.prior = prior;
    a.D[0] = D1; a.D[1] = D2;
    dim3 grid(sp.nseg, g.H, 1);
    k_matching<<<grid, kThreads, smem, s>>>(a);
    count_launch();
}

}  // namespace elasb
