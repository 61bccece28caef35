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

struct Best { int e1, d1, e2; };

__device__ __forceinline__ void scan_update(Best& b, int sum, int d)
{
    // elas.cpp:417-428
    if (sum < b.e1) { b.e2 = b.e1; b.e1 = sum; b.d1 = d; }
    else if (sum < b.e2) { b.e2 = sum; }
}

__device__ __forceinline__ Best warp_merge(Best b)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        int oe1 = __shfl_xor_sync(0xffffffffu, b.e1, off);
        int od1 = __shfl_xor_sync(0xffffffffu, b.d1, off);
        int oe2 = __shfl_xor_sync(0xffffffffu, b.e2, off);
        // d1 == -1 marks "nothing evaluated" and carries e1 = 32767
        bool other_wins = (oe1 < b.e1) || (oe1 == b.e1 && od1 >= 0 && (b.d1 < 0 || od1 < b.d1));
        if (other_wins) { b.e2 = min(oe2, b.e1); b.e1 = oe1; b.d1 = od1; }
        else            { b.e2 = min(b.e2, oe1); }
    }
    return b;
}

// The four descriptor rows a lattice row needs (rows v-2 and v+2 of both images), staged in shared
// memory for one segment of lattice points: rowX[img][k] = descriptor img at column org + k.
struct Strips {
    const uint4* rowA[2];    // row v-2 of desc1 / desc2
    const uint4* rowB[2];    // row v+2
    int org;                 // column of entry 0 (same for all four strips)
};

// computeMatchingDisparity for one (u,v); all lanes of the warp call it with the same arguments.
// own_img = 0: reference pixel in the left image (forward match), 1: in the right image (reverse).
__device__ __forceinline__ int match_point(const FrameGeom& g, const elas_b200_params& p, int u, int v,
                                           const uint4* __restrict__ own_center, const Strips& st,
                                           int own_img, int lane)
{
    const int u_step = 2, window = 3, v_step = 2;
    if (!(u >= window + u_step && u <= g.W - window - 1 - u_step &&
          v >= window + v_step && v <= g.H - window - 1 - v_step)) return -1;        // :337
    if ((int)texture16(__ldg(own_center + (size_t)v * g.W + u)) < p.support_texture) return -1;   // :358-366

    const bool right_image = own_img != 0;
    const int dmin = max(p.disp_min, 0);                                             // :384-387
    const int dmax = right_image ? min(p.disp_max, g.W - u - window - u_step)
                                 : min(p.disp_max, u - window - u_step);
    if (dmax - dmin < 10) return -1;                                                 // :390

    const uint4* ownA = st.rowA[own_img] - st.org;        // indexable by column
    const uint4* ownB = st.rowB[own_img] - st.org;
    const uint4* othA = st.rowA[1 - own_img] - st.org;
    const uint4* othB = st.rowB[1 - own_img] - st.org;
    const uint4 a1 = ownA[u - u_step], a2 = ownA[u + u_step];                        // :369-372
    const uint4 a3 = ownB[u - u_step], a4 = ownB[u + u_step];

    Best b = {32767, -1, 32767};                                                     // :378-381
    for (int d = dmin + lane; d <= dmax; d += 32) {                                  // :396-429
        const int uw = right_image ? u + d : u - d;
        int sum = sad16(a1, othA[uw - u_step]);
        sum += sad16(a2, othA[uw + u_step]);
        sum += sad16(a3, othB[uw - u_step]);
        sum += sad16(a4, othB[uw + u_step]);
        scan_update(b, sum, d);
    }
    b = warp_merge(b);
    // :432 -- (float)min_1_E < support_threshold * (float)min_2_E; both minima exist because the
    // range holds at least 11 disparities
    if (b.d1 >= 0 && (float)b.e1 < __fmul_rn(p.support_threshold, (float)b.e2)) return b.d1;
    return -1;
}

constexpr int kPointsPerCta = 32;     // lattice points of one row handled by a CTA (8 warps x 4 points)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256)
k_support(FrameGeom g, elas_b200_params p, const uint4* __restrict__ desc1,
          const uint4* __restrict__ desc2, int16_t* __restrict__ dcan)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
    const int len = hi - lo, cap = (kPointsPerCta - 1) * g.step + 5 + 2 * p.disp_max;
    uint4* s = reinterpret_cast<uint4*>(smem_raw);
    Strips st;
    st.rowA[0] = s; st.rowA[1] = s + cap; st.rowB[0] = s + 2 * cap; st.rowB[1] = s + 3 * cap;
    st.org = lo;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)len * 16u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(4u * bytes) : "memory");
        const size_t ra = (size_t)(v - 2) * g.W + lo, rb = (size_t)(v + 2) * g.W + lo;
        const uint4* src[4] = {desc1 + ra, desc2 + ra, desc1 + rb, desc2 + rb};
        const uint4* dst[4] = {st.rowA[0], st.rowA[1], st.rowB[0], st.rowB[1]};
#pragma unroll
        for (int k = 0; k < 4; k++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(dst[k])), "l"(src[k]), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(&bar)) : "memory");

    __shared__ int16_t s_result[kPointsPerCta];
    for (int i = warp; i < npts; i += 8) {
        const int uc = uc0 + i;
        int result = 0;                                       // calloc'ed column 0
        if (uc >= 1) {
            const int u = uc * g.step;
            result = -1;
            const int d = match_point(g, p, u, v, desc1, st, 0, lane);             // :482
            if (d >= 0) {
                const int d2 = match_point(g, p, u - d, v, desc2, st, 1, lane);    // :486
                if (d2 >= 0 && abs(d - d2) <= p.lr_threshold) result = d;           // :487-490
            }
        }
        if (lane == 0) s_result[i] = (int16_t)result;
    }
    // dcan is pinned HOST memory (the host stage consumes the lattice): one coalesced store per CTA
    // crosses PCIe instead of a device->host copy queued behind other slots' disparity-map copies
    __syncthreads();
    if (threadIdx.x < npts) dcan[vc * g.Wc + uc0 + threadIdx.x] = s_result[threadIdx.x];
}

}  // namespace

void launch_support(const FrameGeom& g, const elas_b200_params& p, const uint4* desc1,
                    const uint4* desc2, int16_t* dcan, cudaStream_t s)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_support, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    const int cap = (kPointsPerCta - 1) * g.step + 5 + 2 * p.disp_max;
    const size_t smem = (size_t)4 * cap * 16;
    dim3 grid((g.Wc + kPointsPerCta - 1) / kPointsPerCta, g.Hc);
    k_support<<<grid, 256, smem, s>>>(g, p, desc1, desc2, dcan);
    count_launch();
}

}  // namespace elasb
