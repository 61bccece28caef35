// K2: support matching on the candidate lattice (elas.cpp:322-445 called from :471-493).
//
// One CTA per (lattice row, segment of 32 lattice points): the four descriptor rows the row needs
// (v-2 and v+2 of both images, only the columns any match of the segment can touch) are staged in
// shared memory by four TMA bulk copies on one mbarrier.  One warp per lattice point: the four
// 16-byte blocks of the reference pixel (at (u+-2, v+-2), elas.cpp:329-332) live in registers; lanes
// stride over the disparity range, each lane reading the four blocks of the other image at its
// disparity with LDS.128 and taking the byte SAD with VABSDIFF4.  Each lane keeps the best and
// second-best energy in the reference's scan order; a warp-shuffle reduction merges them with the
// reference's tie-break (strict '<' while d ascends => the smaller d wins, the second-best is the
// second order statistic of the energies).  The forward match is followed, in the same warp, by the
// reverse match from (u-d, v) in the right image (elas.cpp:486-490).
//
// The lattice is calloc'ed by the reference (elas.cpp:464): row 0 and column 0 stay 0, a valid
// disparity that takes part in the filters that follow; this kernel writes them too.
#include "common.cuh"

namespace elasb {
namespace {

// Per-lane running best / second best in the reference's scan order (elas.cpp:417-428):
//   if (sum < e1) { e2 = e1; e1 = sum; d1 = d; } else if (sum < e2) e2 = sum;
// kept as key = (energy << 16 | d): a lane visits its disparities in ascending order, so among equal
// energies the smaller key is the earlier d, which is what the strict '<' keeps.  Energies stay below
// 4 * 16 * 255 < 2^15 and d below 2^12 (checked at context creation).
constexpr int kStripPad = 12;         // addressable entries before and after every staged strip (see match_point)

// Both the best and the second best are kept as KEYS: keys are unique (distinct d), ordering by key is ordering by
// (energy, d), so the energy of the second-smallest key is the second order statistic of the energies.
struct Best { unsigned key, key2; };
constexpr unsigned kNoKey = (32767u << 16) | 0xFFFFu;     // nothing evaluated: e1 = e2 = 32767 (elas.cpp:378-381)

__device__ __forceinline__ void scan_update(Best& b, int sum, int d)
{
    const unsigned key = ((unsigned)sum << 16) | (unsigned)d;
    b.key2 = min(b.key2, max(b.key, key));                // the one that does not become / stay the best
    b.key = min(b.key, key);
}

// Warp-wide (best energy, its disparity, second-best energy) with two REDUX reductions: the best key is
// the minimum key (ties between lanes: the smaller d, as in the reference's single ascending scan); the
// second order statistic of all energies is the minimum over lanes of "my second best if I hold the
// winner, else my best".
__device__ __forceinline__ void warp_merge(const Best& b, int& e1, int& d1, int& e2)
{
    const unsigned best = __reduce_min_sync(0xffffffffu, b.key);
    const unsigned mine = b.key == best ? b.key2 : b.key;
    e2 = (int)(__reduce_min_sync(0xffffffffu, mine) >> 16);
    e1 = (int)(best >> 16);
    d1 = best == kNoKey ? -1 : (int)(best & 0xFFFFu);
}

// The four descriptor rows a lattice row needs (rows v-2 and v+2 of both images), staged in shared
// memory for one segment of lattice points: rowX[img][k] = descriptor img at column org + k.
struct Strips {
    const uint4* rowA[2];    // row v-2 of desc1 / desc2
    const uint4* rowB[2];    // row v+2
    int org;                 // column of entry 0 (same for all four strips)
    int len;                 // columns staged
    const uint4* cen[2];     // row v of desc1 / desc2: the reference pixels' own descriptors (texture test)
    int cen_org[2];
};

// computeMatchingDisparity for one (u,v); all lanes of the warp call it with the same arguments.
// own_img = 0: reference pixel in the left image (forward match), 1: in the right image (reverse).
//
// The energy of disparity d adds four block SADs at the other image's columns c-2 and c+2 (rows v-2, v+2),
// c = u -/+ d (elas.cpp:329-332, :400-414).  Disparities d and d+4 therefore share a column: a lane takes
// the three disparities d, d+4, d+8 together and loads 4 columns x 2 rows instead of 6 x 2 -- shared-memory
// bandwidth (LDS.128 = 4 wavefronts per warp) is what bounds this kernel, not the SAD arithmetic.  Lane l
// of pass p owns d = dmin + 96 p + 12 (l / 4) + (l % 4) + {0, 4, 8}: every quarter warp reads eight
// columns that are distinct modulo 8, i.e. conflict-free 16-byte accesses.
__device__ __forceinline__ int match_point(const FrameGeom& g, const elas_b200_params& p, int u, int v,
                                           const Strips& st, int own_img, int lane)
{
    const int u_step = 2, window = 3, v_step = 2;
    if (!(u >= window + u_step && u <= g.W - window - 1 - u_step &&
          v >= window + v_step && v <= g.H - window - 1 - v_step)) return -1;        // :337
    if ((int)texture16(st.cen[own_img][u - st.cen_org[own_img]]) < p.support_texture) return -1;   // :358-366

    const bool right_image = own_img != 0;
    const int dmin = max(p.disp_min, 0);                                             // :384-387
    const int dmax = right_image ? min(p.disp_max, g.W - u - window - u_step)
                                 : min(p.disp_max, u - window - u_step);
    if (dmax - dmin < 10) return -1;                                                 // :390

    const uint4* ownA = st.rowA[own_img] - st.org;        // indexable by column
    const uint4* ownB = st.rowB[own_img] - st.org;
    const uint4* othA = st.rowA[1 - own_img];             // indexable by column - org (clamped below)
    const uint4* othB = st.rowB[1 - own_img];
    const uint4 a1 = ownA[u - u_step], a2 = ownA[u + u_step];                        // :369-372
    const uint4 a3 = ownB[u - u_step], a4 = ownB[u + u_step];

    const int sgn = right_image ? 1 : -1;                 // warped column = u + sgn * d
    const int lane_off = 12 * (lane >> 2) + (lane & 3);
    Best b = {kNoKey, kNoKey};                                                       // :378-381
    for (int d0 = dmin + lane_off; d0 <= dmax; d0 += 96) {                           // :396-429
        // the four columns X_j = (u + sgn*d0) + sgn*(4j - 2), j = 0..3, as strip indices; the disparities d0+4 and
        // d0+8 may lie past dmax: they are computed but not counted, their columns (at most 10 beyond the staged
        // range) fall into the kStripPad entries of addressable padding around every strip
        const int c0 = u + sgn * d0 - st.org - 2 * sgn;
        uint4 A[4], B[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            A[j] = othA[c0 + 4 * sgn * j];
            B[j] = othB[c0 + 4 * sgn * j];
        }
#pragma unroll
        for (int m = 0; m < 3; m++) {
            // block "u-2" of this disparity is column X_m (right image) or X_{m+1} (left image), block "u+2" the other one
            const int lo = right_image ? m : m + 1, hi = right_image ? m + 1 : m;
            unsigned s0 = 0, s1 = 0;
            sad16_acc(a1, right_image ? A[m] : A[m + 1], s0, s1);
            sad16_acc(a2, right_image ? A[m + 1] : A[m], s0, s1);
            sad16_acc(a3, right_image ? B[m] : B[m + 1], s0, s1);
            sad16_acc(a4, right_image ? B[m + 1] : B[m], s0, s1);
            (void)lo; (void)hi;
            const int d = d0 + 4 * m;
            if (d <= dmax) scan_update(b, (int)(s0 + s1), d);
        }
    }
    int e1, d1, e2;
    warp_merge(b, e1, d1, e2);
    // :432 -- (float)min_1_E < support_threshold * (float)min_2_E; both minima exist because the
    // range holds at least 11 disparities
    if (d1 >= 0 && (float)e1 < __fmul_rn(p.support_threshold, (float)e2)) return d1;
    return -1;
}

constexpr int kPointsPerCta = 32;     // lattice points of one row handled by a CTA (8 warps x 4 points)


__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256)
k_support(FrameGeom g, elas_b200_params p, const uint4* __restrict__ desc1_g,
          const uint4* __restrict__ desc2_g, int16_t* __restrict__ dcan_g, size_t desc_stride, size_t dcan_stride)
{
    // blockIdx.z = frame of the group
    const uint4* __restrict__ desc1 = desc1_g + (size_t)blockIdx.z * desc_stride;
    const uint4* __restrict__ desc2 = desc2_g + (size_t)blockIdx.z * desc_stride;
    int16_t* __restrict__ dcan = dcan_g + (size_t)blockIdx.z * dcan_stride;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const int lane = threadIdx.x & 31;
    const int vc = blockIdx.y, uc0 = blockIdx.x * kPointsPerCta;
    const int v = vc * g.step;
    const int npts = min(kPointsPerCta, g.Wc - uc0);
    const bool row_ok = vc >= 1 && v >= 5 && v <= g.H - 6;                           // :337 for the whole row
    if (!row_ok) {
        // calloc'ed row 0 (:464) stays 0; rows that fail the window test hold -1 (column 0 stays 0)
        if (threadIdx.x < npts) dcan[vc * g.Wc + uc0 + threadIdx.x] = (int16_t)((vc == 0 || uc0 + threadIdx.x == 0) ? 0 : -1);
        return;
    }
    // columns any match of this segment can touch: forward needs desc2 down to x0-2-disp_max, the
    // reverse match (from u-d) needs desc1 up to x1+2+disp_max
    const int x0 = uc0 * g.step, x1 = (uc0 + npts - 1) * g.step;
    const int lo = max(x0 - 2 - p.disp_max, 0), hi = min(x1 + 2 + p.disp_max + 1, g.W);
    const int len = hi - lo, cap = (kPointsPerCta - 1) * g.step + 5 + 2 * p.disp_max + 2 * kStripPad;
    uint4* s = reinterpret_cast<uint4*>(smem_raw) + kStripPad;
    Strips st;
    st.rowA[0] = s; st.rowA[1] = s + cap; st.rowB[0] = s + 2 * cap; st.rowB[1] = s + 3 * cap;
    st.org = lo; st.len = len;
    // row v itself is needed only at the reference pixels (texture test): the forward matches start from the
    // CTA's lattice points in desc1, the reverse matches from (u-d, v) in desc2
    const int cen_cap = (kPointsPerCta - 1) * g.step + 1;
    const int c1lo = max(x0 - p.disp_max, 0);
    uint4* cen0 = s + 4 * cap - kStripPad;
    uint4* cen1 = cen0 + cen_cap;
    st.cen[0] = cen0; st.cen[1] = cen1; st.cen_org[0] = x0; st.cen_org[1] = c1lo;
    __shared__ int s_next;
    __shared__ int16_t s_result[kPointsPerCta];
    if (threadIdx.x == 0) {
        s_next = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)len * 16u;
        const uint32_t b0 = (uint32_t)(x1 - x0 + 1) * 16u, b1 = (uint32_t)(x1 - c1lo + 1) * 16u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(4u * bytes + b0 + b1) : "memory");
        const size_t ra = (size_t)(v - 2) * g.W + lo, rb = (size_t)(v + 2) * g.W + lo, rc = (size_t)v * g.W;
        const uint4* src[6] = {desc1 + ra, desc2 + ra, desc1 + rb, desc2 + rb, desc1 + rc + x0, desc2 + rc + c1lo};
        const uint4* dst[6] = {st.rowA[0], st.rowA[1], st.rowB[0], st.rowB[1], cen0, cen1};
        const uint32_t nbytes[6] = {bytes, bytes, bytes, bytes, b0, b1};
#pragma unroll
        for (int k = 0; k < 6; k++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(dst[k])), "l"(src[k]), "r"(nbytes[k]), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(&bar)) : "memory");

    // lattice points are handed out dynamically: a point that fails its texture test costs almost
    // nothing, one that matches costs two full scans, so a fixed assignment leaves warps idle at the end
    for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(&s_next, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= npts) break;
        const int uc = uc0 + i;
        int result = 0;                                       // calloc'ed column 0
        if (uc >= 1) {
            const int u = uc * g.step;
            result = -1;
            const int d = match_point(g, p, u, v, st, 0, lane);                    // :482
            if (d >= 0) {
                const int d2 = match_point(g, p, u - d, v, st, 1, lane);           // :486
                if (d2 >= 0 && abs(d - d2) <= p.lr_threshold) result = d;           // :487-490
            }
        }
        if (lane == 0) s_result[i] = (int16_t)result;
    }
    __syncthreads();
    if (threadIdx.x < npts) dcan[vc * g.Wc + uc0 + threadIdx.x] = s_result[threadIdx.x];
}

}  // namespace

size_t support_smem_bytes(const FrameGeom& g, const elas_b200_params& p)
{
    const int cap = (kPointsPerCta - 1) * g.step + 5 + 2 * p.disp_max + 2 * kStripPad;
    const int cen_cap = (kPointsPerCta - 1) * g.step + 1;
    return ((size_t)4 * cap + 2 * cen_cap + p.disp_max) * 16;
}

void launch_support(const FrameGeom& g, const elas_b200_params& p, const uint4* desc1,
                    const uint4* desc2, int16_t* dcan, const GroupStrides& st, int n_frames, cudaStream_t s)
{
    static unsigned long long optin = 0;
    if (ensure_dynamic_smem(k_support, 200 * 1024, &optin) != cudaSuccess) return;
    const size_t smem = support_smem_bytes(g, p);
    dim3 grid((g.Wc + kPointsPerCta - 1) / kPointsPerCta, g.Hc, n_frames);
    k_support<<<grid, 256, smem, s>>>(g, p, desc1, desc2, dcan, st.desc, st.dcan);
    count_launch();
}

}  // namespace elasb
