// K2: support matching on the candidate lattice (elas.cpp:322-445 called from :471-493).
//
// One warp per lattice point.  The four 16-byte blocks of the reference pixel (at (u+-2, v+-2),
// elas.cpp:329-332) live in registers; lanes stride over the disparity range, each lane loading the
// four blocks of the other image at its disparity with 16-byte vector loads (a warp reads 4 x 512
// contiguous bytes per step) and taking the byte SAD with VABSDIFF4.  Each lane keeps the best and
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

// computeMatchingDisparity for one (u,v); all lanes of the warp call it with the same arguments.
__device__ __forceinline__ int match_point(const FrameGeom& g, const elas_b200_params& p, int u, int v,
                                           const uint4* __restrict__ own, const uint4* __restrict__ other,
                                           bool right_image, int lane)
{
    const int u_step = 2, v_step = 2, window = 3;
    if (!(u >= window + u_step && u <= g.W - window - 1 - u_step &&
          v >= window + v_step && v <= g.H - window - 1 - v_step)) return -1;        // :337
    if ((int)texture16(__ldg(own + (size_t)v * g.W + u)) < p.support_texture) return -1;   // :358-366

    const int dmin = max(p.disp_min, 0);                                             // :384-387
    const int dmax = right_image ? min(p.disp_max, g.W - u - window - u_step)
                                 : min(p.disp_max, u - window - u_step);
    if (dmax - dmin < 10) return -1;                                                 // :390

    const size_t rowA = (size_t)(v - v_step) * g.W, rowB = (size_t)(v + v_step) * g.W;
    const uint4 a1 = __ldg(own + rowA + u - u_step), a2 = __ldg(own + rowA + u + u_step);   // :369-372
    const uint4 a3 = __ldg(own + rowB + u - u_step), a4 = __ldg(own + rowB + u + u_step);

    Best b = {32767, -1, 32767};                                                     // :378-381
    for (int d = dmin + lane; d <= dmax; d += 32) {                                  // :396-429
        const int uw = right_image ? u + d : u - d;
        int sum = sad16(a1, __ldg(other + rowA + uw - u_step));
        sum += sad16(a2, __ldg(other + rowA + uw + u_step));
        sum += sad16(a3, __ldg(other + rowB + uw - u_step));
        sum += sad16(a4, __ldg(other + rowB + uw + u_step));
        scan_update(b, sum, d);
    }
    b = warp_merge(b);
    // :432 -- (float)min_1_E < support_threshold * (float)min_2_E; both minima exist because the
    // range holds at least 11 disparities
    if (b.d1 >= 0 && (float)b.e1 < __fmul_rn(p.support_threshold, (float)b.e2)) return b.d1;
    return -1;
}

__global__ void __launch_bounds__(256)
k_support(FrameGeom g, elas_b200_params p, const uint4* __restrict__ desc1,
          const uint4* __restrict__ desc2, int16_t* __restrict__ dcan)
{
    const int lane = threadIdx.x & 31;
    const int point = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (point >= g.Wc * g.Hc) return;
    const int uc = point % g.Wc, vc = point / g.Wc;
    int result = 0;                                       // calloc'ed row 0 / column 0
    if (uc >= 1 && vc >= 1) {
        const int u = uc * g.step, v = vc * g.step;
        result = -1;
        int d = match_point(g, p, u, v, desc1, desc2, false, lane);               // :482
        if (d >= 0) {
            int d2 = match_point(g, p, u - d, v, desc2, desc1, true, lane);       // :486
            if (d2 >= 0 && abs(d - d2) <= p.lr_threshold) result = d;             // :487-490
        }
    }
    if (lane == 0) dcan[point] = (int16_t)result;
}

}  // namespace

void launch_support(const FrameGeom& g, const elas_b200_params& p, const uint4* desc1,
                    const uint4* desc2, int16_t* dcan, cudaStream_t s)
{
    const int points = g.Wc * g.Hc, per_block = 8;
    k_support<<<(points + per_block - 1) / per_block, per_block * 32, 0, s>>>(g, p, desc1, desc2, dcan);
    count_launch();
}

}  // namespace elasb
