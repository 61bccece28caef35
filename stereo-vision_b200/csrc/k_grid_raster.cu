// K6: candidate grid (elas.cpp:684-780) and the triangle-id maps that replace the reference's
// per-triangle scan conversion (elas.cpp:1003-1115).
//
// Grid layout in HBM: the reference stores, per 20x20-pixel cell, a count and an ascending list of
// candidate disparities as int32[disp_max+2] (1028 B per cell at disp_max 255).  The same set is
// kept here as a bitmask of disp_max+1 bits per cell (32 B): bit d set <=> d is in the list.  The
// ascending list order the matching kernel needs is the order of the set bits.  The expansion back to
// the reference layout (for parity checks) is done on the host by the stage-dump hook.
#include <algorithm>

#include "common.cuh"

namespace elasb {
namespace {

__device__ __forceinline__ int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// elas.cpp:697-727: every support point marks d-1..d+1 in its cell, for the left image at
// (u/grid_size, v/grid_size) and for the right image at (floor((u-d)/grid_size), v/grid_size).
// scratch = two bitmask planes [2][gh*gw][gwords], zeroed before this kernel.
__device__ __forceinline__ void grid_scatter_body(int i, const FrameGeom& g, const elas_b200_params& p,
                                                  const int32_t* __restrict__ support, int n,
                                                  uint32_t* __restrict__ scratch)
{
    if (i >= n) return;
    const int u = support[3 * i], v = support[3 * i + 1], d = support[3 * i + 2];
    const int y = floor_div(v, p.grid_size);
    const int cells = g.gw * g.gh;
#pragma unroll
    for (int img = 0; img < 2; img++) {
        const int x = img ? floor_div(u - d, p.grid_size) : u / p.grid_size;     // :712, :716
        if (x < 0 || x >= g.gw || y < 0 || y >= g.gh) continue;                   // :721
        uint32_t* cell = scratch + ((size_t)img * cells + (size_t)y * g.gw + x) * g.gwords;
        for (int dd = max(d - 1, 0); dd <= min(d + 1, p.disp_max); dd++)          // :703-707
            atomicOr(cell + (dd >> 5), 1u << (dd & 31));
    }
}

// elas.cpp:732-751: the reference walks nine pointers over the flat temp1 array in lock-step, so a
// cell's "3x3 neighbourhood" is the nine FLAT offsets {-gw-1,-gw,-gw+1,-1,0,1,gw-1,gw,gw+1} (it
// wraps across row ends) and only flat cells gw+1 .. gw*gh-gw-2 are written; all others stay empty
// (SURVEY A.7).  One thread per (image, cell): ORs the nine scratch masks word by word, stores the
// cell's bitmask and, for the matching kernel, the same set as an ascending uint16 list
// (elas.cpp:754-775 packs exactly this list): list[0] = count, list[1..] = disparities; a cell with
// more than kGridListCap candidates gets count 0xFFFF and is read from its bitmask instead.
// The scatter planes are double-buffered per slot: while this frame's planes are read, the other
// buffer (used by the previous frame, needed again by the next one) is zeroed -- no memset launch.
__device__ __forceinline__ void grid_diffuse_body(int i, const FrameGeom& g, const uint32_t* __restrict__ scratch,
                                                  uint32_t* __restrict__ scratch_next,
                                                  uint32_t* __restrict__ grid1, uint32_t* __restrict__ grid2,
                                                  uint16_t* __restrict__ lists1, uint16_t* __restrict__ lists2)
{
    const int cells = g.gw * g.gh;
    if (i >= 2 * cells) return;
    const int img = i >= cells, c = img ? i - cells : i;
    uint32_t* bits = (img ? grid2 : grid1) + (size_t)c * g.gwords;
    uint16_t* list = (img ? lists2 : lists1) + (size_t)c * kGridListStride;
    const bool written = c >= g.gw + 1 && c <= cells - g.gw - 2;
    int count = 0;
    for (int w = 0; w < g.gwords; w++) {
        uint32_t m = 0;
        if (written) {
            const uint32_t* t = scratch + (size_t)img * cells * g.gwords + w;
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) m |= t[(size_t)(c + dy * g.gw + dx) * g.gwords];
        }
        bits[w] = m;
        scratch_next[(size_t)i * g.gwords + w] = 0u;
        while (m) {
            if (count < kGridListCap) list[1 + count] = (uint16_t)(32 * w + __ffs(m) - 1);
            count++;
            m &= m - 1;
        }
    }
    list[0] = count <= kGridListCap ? (uint16_t)count : (uint16_t)0xFFFF;
    if (count <= kGridListCap) list[1 + count] = (uint16_t)g.dn;      // sentinel: disp_max + 1
}

// Disparity planes (elas.cpp:605-680) and the per-triangle set-up of computeDisparity
// (elas.cpp:1006-1072), one thread per triangle of either image.  Matrix::solve (matrix.cpp:414-502)
// is Gauss-Jordan elimination with full pivoting in double precision; it is restated with explicit
// round-to-nearest intrinsics so that no multiply-add is contracted: the reference is x86-64 SSE2
// code, every operation rounds separately, and the planes must come out bit-identical.
__device__ bool solve3(double A[3][3], double b[3])
{
    int ipiv[3] = {0, 0, 0};
#pragma unroll 1
    for (int i = 0; i < 3; i++) {
        double big = 0.0;
        int irow = 0, icol = 0;
        for (int j = 0; j < 3; j++)
            if (ipiv[j] != 1)
                for (int k = 0; k < 3; k++)
                    if (ipiv[k] == 0 && fabs(A[j][k]) >= big) { big = fabs(A[j][k]); irow = j; icol = k; }
        ++ipiv[icol];
        if (irow != icol) {
            for (int l = 0; l < 3; l++) { const double t = A[irow][l]; A[irow][l] = A[icol][l]; A[icol][l] = t; }
            const double t = b[irow]; b[irow] = b[icol]; b[icol] = t;
        }
        if (fabs(A[icol][icol]) < 1e-20) return false;
        const double pivinv = __ddiv_rn(1.0, A[icol][icol]);
        A[icol][icol] = 1.0;
        for (int l = 0; l < 3; l++) A[icol][l] = __dmul_rn(A[icol][l], pivinv);
        b[icol] = __dmul_rn(b[icol], pivinv);
        for (int ll = 0; ll < 3; ll++)
            if (ll != icol) {
                const double dum = A[ll][icol];
                A[ll][icol] = 0.0;
                for (int l = 0; l < 3; l++) A[ll][l] = __dsub_rn(A[ll][l], __dmul_rn(A[icol][l], dum));
                b[ll] = __dsub_rn(b[ll], __dmul_rn(b[icol], dum));
            }
    }
    return true;
}

__device__ __forceinline__ void planes_body(int gid, const int32_t* __restrict__ support,
                                            const int32_t* __restrict__ tri1, int nt1,
                                            const int32_t* __restrict__ tri2, int nt2, TriRaster* __restrict__ out1,
                                            TriRaster* __restrict__ out2, float* __restrict__ planes1, float* __restrict__ planes2)
{
    // two threads per triangle: the even lane fits the plane in left-image coordinates (t1), the odd lane
    // the one in right-image coordinates (t2); the two solves are the long serial part of this kernel.
    // Whole warps call this together (shuffles below).
    const int k = gid & 1;
    int i = gid >> 1;
    const bool active = i < nt1 + nt2;
    const int right_image = i >= nt1;
    if (right_image) i -= nt1;
    int su[3] = {0, 1, 2}, sv[3] = {0, 0, 1}, sd[3] = {0, 0, 0};
    if (active) {
        const int32_t* tri = (right_image ? tri2 : tri1) + 3 * i;
        for (int c = 0; c < 3; c++) {
            const int32_t* s = support + 3 * (size_t)tri[c];
            su[c] = s[0]; sv[c] = s[1]; sd[c] = s[2];
        }
    }
    float mine[3];
    {
        double A[3][3], b[3];
        for (int c = 0; c < 3; c++) {
            A[c][0] = k ? su[c] - sd[c] : su[c];
            A[c][1] = sv[c];
            A[c][2] = 1.0;
            b[c] = sd[c];
        }
        const bool ok = solve3(A, b);
        for (int c = 0; c < 3; c++) mine[c] = ok ? __double2float_rn(b[c]) : 0.f;
    }
    float pl[6];
    for (int c = 0; c < 3; c++) {
        const float other = __shfl_xor_sync(0xffffffffu, mine[c], 1);
        pl[c] = k ? other : mine[c];
        pl[3 + c] = k ? mine[c] : other;
    }
    if (!active || k) return;
    float* po = (right_image ? planes2 : planes1) + 6 * (size_t)i;
    for (int c = 0; c < 6; c++) po[c] = pl[c];

    TriRaster r;
    const float pd = right_image ? pl[0] : pl[3];
    r.pa = right_image ? pl[3] : pl[0];
    r.pb = right_image ? pl[4] : pl[1];
    r.pc = right_image ? pl[5] : pl[2];
    float tu[3], tv[3];
    for (int c = 0; c < 3; c++) { tu[c] = right_image ? (float)(su[c] - sd[c]) : (float)su[c]; tv[c] = (float)sv[c]; }
    // the reference's 3-element bubble sort by u (elas.cpp:1043-1053), unrolled: (1,0) (2,0) (2,1)
#define ELASB_SWAP_IF(k, j) if (tu[k] > tu[j]) { float t_ = tu[j]; tu[j] = tu[k]; tu[k] = t_; t_ = tv[j]; tv[j] = tv[k]; tv[k] = t_; }
    ELASB_SWAP_IF(0, 1) ELASB_SWAP_IF(0, 2) ELASB_SWAP_IF(1, 2)
#undef ELASB_SWAP_IF
    const float Au = tu[0], Av = tv[0], Bu = tu[1], Bv = tv[1], Cu = tu[2], Cv = tv[2];
    float ABa = 0.f, ACa = 0.f, BCa = 0.f;               // :1061-1067
    if ((int)Au != (int)Bu) ABa = __fdiv_rn(__fsub_rn(Av, Bv), __fsub_rn(Au, Bu));
    if ((int)Au != (int)Cu) ACa = __fdiv_rn(__fsub_rn(Av, Cv), __fsub_rn(Au, Cu));
    if ((int)Bu != (int)Cu) BCa = __fdiv_rn(__fsub_rn(Bv, Cv), __fsub_rn(Bu, Cu));
    r.ABa = ABa; r.ACa = ACa; r.BCa = BCa;
    r.ABb = __fsub_rn(Av, __fmul_rn(ABa, Au));
    r.ACb = __fsub_rn(Av, __fmul_rn(ACa, Au));
    r.BCb = __fsub_rn(Bv, __fmul_rn(BCa, Bu));
    r.uA = (int)Au; r.uB = (int)Bu; r.uC = (int)Cu;
    r.valid = ((double)fabsf(r.pa) < 0.7 && (double)fabsf(pd) < 0.7) ? 2 : 0;     // :1072; bit 1 of K7's packed pixel state
    r.pad0 = min(sv[0], min(sv[1], sv[2]));     // smallest corner row: anchor of k_raster's row bands
    r.pad1 = max(sv[0], max(sv[1], sv[2]));     // largest corner row (overflow triangles walk all their bands)
    r.pad2 = 0;
    (right_image ? out2 : out1)[i] = r;
}

// Scan conversion (elas.cpp:1074-1114), one work unit per warp: 32 columns x kRasterBandRows rows of one
// triangle's bounding box (units are listed by the host stage, which knows the corner coordinates).
// The reference walks the columns u of both halves (A->B, B->C) and, per column, the half-open row
// range [min(v1,v2), max(v1,v2)) with
//   v1 = (uint32_t)(AC_a*u + AC_b),  v2 = (uint32_t)(AB_a*u + AB_b)   (separate mul and add, no FMA;
// the x86-64 conversion truncates through 64 bits, i.e. trunc toward zero for the values met here).
// Each lane owns a column and computes its row range once; the warp then sweeps the rows of the band
// together, so every store instruction touches one row segment (coalesced) with warp-uniform loop
// bounds.  The reference lets later triangles overwrite earlier ones; atomicMax on the triangle index
// gives the same winner (findMatch's early returns depend on the pixel only, never on the triangle).
//
// Map entries are (frame tag << tag_shift) | triangle index: a new frame's entries compare greater than
// anything an earlier frame left behind, so the maps are never cleared between frames (the matching
// kernel ignores entries whose tag is not the current one).
__device__ __forceinline__ void raster_unit(int t, int img, int chunk, int band, const FrameGeom& g, int subsampling,
                                            const TriRaster* __restrict__ tri1, const TriRaster* __restrict__ tri2,
                                            int32_t* __restrict__ map1, int32_t* __restrict__ map2, int tag_bits)
{
    const int lane = threadIdx.x & 31;
    const int pitch = map_pitch(g);
    const TriRaster* tri = (img ? tri2 : tri1) + t;
    int32_t* map = img ? map2 : map1;
    const float4 e0 = __ldg(reinterpret_cast<const float4*>(tri));          // ACa, ACb, ABa, ABb
    const float4 e1 = __ldg(reinterpret_cast<const float4*>(tri) + 1);      // BCa, BCb, uA, uB
    const int uA = __float_as_int(e1.z), uB = __float_as_int(e1.w), uC = __ldg(&tri->uC);
    const int u_end = min(uC, g.W);                                          // :1077, :1098
    const int u = max(uA, 0) + 32 * chunk + lane;
    int lo = 0, hi = 0;
    if (u < u_end && !(subsampling && (u & 1))) {
        const bool first = u < uB;                                           // A->B part, else B->C part
        const float fu = (float)u;
        const int v1 = __float2int_rz(__fadd_rn(__fmul_rn(e0.x, fu), e0.y));                                  // :1081, :1102
        const int v2 = __float2int_rz(__fadd_rn(__fmul_rn(first ? e0.z : e1.x, fu), first ? e0.w : e1.y));    // :1082, :1103
        lo = max(min(v1, v2), 0);
        hi = min(max(v1, v2), g.H);
    }
    // rows of this unit: band `band` of the triangle's row range; the band grid is anchored at the
    // smallest corner row, which bounds every column's range from below
    int vmin = hi > lo ? lo : g.H, vmax = hi > lo ? hi : 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        vmin = min(vmin, __shfl_xor_sync(0xffffffffu, vmin, off));
        vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, off));
    }
    const int v_anchor = max(__ldg(&tri->pad0) - 1, 0);                      // min corner row - 1 (mesh_core.h raster_box)
    const int b_lo = v_anchor + band * kRasterBandRows, b_hi = b_lo + kRasterBandRows;
    const int vlo = max(vmin, b_lo), vhi = min(vmax, b_hi);
    int32_t* col = map + u;
    for (int v = vlo; v < vhi; v++) {
        if (v >= lo && v < hi && !(subsampling && (v & 1))) atomicMax(col + (size_t)v * pitch, tag_bits | t);
    }
}

// one warp: work item w of a frame = a listed unit of the left image, of the right image, or a whole overflow triangle
__device__ __forceinline__ void raster_item(int w, const FrameHeader& h, const FrameGeom& g, int subsampling,
                                            const TriRaster* __restrict__ tri1, const TriRaster* __restrict__ tri2,
                                            const int2* __restrict__ units1, const int2* __restrict__ units2,
                                            int32_t* __restrict__ map1, int32_t* __restrict__ map2, int tag_bits)
{
    if (w < h.n_units[0] + h.n_units[1]) {
        const int2 u = w < h.n_units[0] ? __ldg(units1 + w) : __ldg(units2 + (w - h.n_units[0]));
        raster_unit(u.x & 0x3FFFFFFF, (u.x >> 30) & 1, u.y & 0xFFFF, u.y >> 16, g, subsampling, tri1, tri2, map1, map2, tag_bits);
        return;
    }
    w -= h.n_units[0] + h.n_units[1];
    const int ovf0 = h.n_tri[0] - h.ovf_from[0];
    const int img = w >= ovf0;
    const int t = img ? h.ovf_from[1] + (w - ovf0) : h.ovf_from[0] + w;
    const TriRaster* tri = (img ? tri2 : tri1) + t;
    const int uA = __ldg(&tri->uA), uC = __ldg(&tri->uC), v0 = __ldg(&tri->pad0), v1 = __ldg(&tri->pad1);
    const int chunks = (min(uC, g.W) - max(uA, 0) + 31) / 32;
    const int bands = (min(v1 + 1, g.H) - max(v0 - 1, 0) + kRasterBandRows - 1) / kRasterBandRows;
    for (int ch = 0; ch < chunks; ch++)
        for (int bd = 0; bd < bands; bd++)
            raster_unit(t, img, ch, bd, g, subsampling, tri1, tri2, map1, map2, tag_bits);
}

// K5 + the scatter half of K6 in one launch, frames of a group in blockIdx.y.  The numbers of support points and
// triangles come from the frame's header in device memory (the mesh stage runs on the GPU), so the grid is a fixed
// size and every block strides over the work: items [0, P) fit planes (two threads per triangle, whole warps),
// items [P, P + n) scatter support points into the candidate-grid bitmasks.  Both read only the frame's tables.
constexpr int kSetupThreads = 128;
__global__ void __launch_bounds__(kSetupThreads)
k_planes_scatter(FrameGeom g, elas_b200_params p, const FrameHeader* __restrict__ hdr, const int32_t* __restrict__ support,
                 const int32_t* __restrict__ tri1, const int32_t* __restrict__ tri2,
                 TriRaster* __restrict__ out1, TriRaster* __restrict__ out2, float* __restrict__ planes1,
                 float* __restrict__ planes2, uint32_t* __restrict__ scratch, GroupStrides st)
{
    const int f = blockIdx.y;
    const FrameHeader h = hdr[f];
    if (h.n_support < 3) return;                                           // elas.cpp:69-75
    support += (size_t)f * st.support;
    tri1 += (size_t)f * st.tri; tri2 += (size_t)f * st.tri;
    out1 += (size_t)f * st.traster; out2 += (size_t)f * st.traster;
    planes1 += (size_t)f * st.planes; planes2 += (size_t)f * st.planes;
    scratch += (size_t)f * st.scratch;
    const int plane_items = (2 * (h.n_tri[0] + h.n_tri[1]) + 31) & ~31;
    const int total = (plane_items + h.n_support + 31) & ~31;
    for (int w = blockIdx.x * kSetupThreads + threadIdx.x; w < total; w += gridDim.x * kSetupThreads) {
        if (w < plane_items) planes_body(w, support, tri1, h.n_tri[0], tri2, h.n_tri[1], out1, out2, planes1, planes2);
        else grid_scatter_body(w - plane_items, g, p, support, h.n_support, scratch);
    }
}

// The diffusion half of K6 + scan conversion in one launch: blocks [0, diffuse_blocks) turn the scatter
// planes into per-cell bitmasks and lists, the rest scan-convert triangle work items (one per warp, strided).
constexpr int kRasterThreads = 256;
__global__ void __launch_bounds__(kRasterThreads)
k_diffuse_raster(FrameGeom g, int subsampling, int diffuse_blocks, const FrameHeader* __restrict__ hdr,
                 const uint32_t* __restrict__ scratch, uint32_t* __restrict__ scratch_next,
                 uint32_t* __restrict__ grid1, uint32_t* __restrict__ grid2,
                 uint16_t* __restrict__ lists1, uint16_t* __restrict__ lists2,
                 const TriRaster* __restrict__ tri1, const TriRaster* __restrict__ tri2,
                 const int2* __restrict__ units1, const int2* __restrict__ units2,
                 int32_t* __restrict__ map1, int32_t* __restrict__ map2, int tag_bits, GroupStrides st)
{
    const int f = blockIdx.y;
    if ((int)blockIdx.x < diffuse_blocks) {
        grid_diffuse_body(blockIdx.x * kRasterThreads + threadIdx.x, g, scratch + (size_t)f * st.scratch,
                          scratch_next + (size_t)f * st.scratch, grid1 + (size_t)f * st.grid, grid2 + (size_t)f * st.grid,
                          lists1 + (size_t)f * st.lists, lists2 + (size_t)f * st.lists);
        return;
    }
    const FrameHeader h = hdr[f];
    if (h.n_support < 3) return;
    const int items = h.n_units[0] + h.n_units[1] + (h.n_tri[0] - h.ovf_from[0]) + (h.n_tri[1] - h.ovf_from[1]);
    const int warps = (gridDim.x - diffuse_blocks) * (kRasterThreads >> 5);
    for (int w = (blockIdx.x - diffuse_blocks) * (kRasterThreads >> 5) + (threadIdx.x >> 5); w < items; w += warps)
        raster_item(w, h, g, subsampling, tri1 + (size_t)f * st.traster, tri2 + (size_t)f * st.traster,
                    units1 + (size_t)f * (st.units / 2), units2 + (size_t)f * (st.units / 2),
                    map1 + (size_t)f * st.map, map2 + (size_t)f * st.map, tag_bits);
}

}  // namespace

void launch_planes_scatter(const FrameGeom& g, const elas_b200_params& p, const FrameHeader* hdr, const int32_t* support,
                           const int32_t* tri1, const int32_t* tri2, TriRaster* out1, TriRaster* out2,
                           float* planes1, float* planes2, uint32_t* scratch, const GroupStrides& st, int n_frames,
                           cudaStream_t s)
{
    // enough threads for a densely textured frame (a third of the lattice survives the filters at most:
    // ~2 triangles per point per image, 2 threads per triangle), strided beyond that
    const int expect = 3 * g.Wc * g.Hc;
    const int blocks = std::max(8, std::min(1024, (expect + kSetupThreads - 1) / kSetupThreads));
    ELASB_PREPARE_KERNEL(k_planes_scatter);
    k_planes_scatter<<<dim3(blocks, n_frames), kSetupThreads, 0, s>>>(g, p, hdr, support, tri1, tri2, out1, out2, planes1,
                                                                      planes2, scratch, st);
    count_launch();
}

void launch_diffuse_raster(const FrameGeom& g, int subsampling, const FrameHeader* hdr, const uint32_t* scratch,
                           uint32_t* scratch_next, uint32_t* grid1, uint32_t* grid2, uint16_t* lists1, uint16_t* lists2,
                           const TriRaster* tri1, const TriRaster* tri2, const int32_t* units1, const int32_t* units2,
                           int32_t* map1, int32_t* map2, int tag_bits, const GroupStrides& st, int n_frames, cudaStream_t s)
{
    const int diffuse_blocks = (2 * g.gw * g.gh + kRasterThreads - 1) / kRasterThreads;
    // one warp per ~32x32 piece of the image and image side is plenty of parallelism; more work is strided
    const int pieces = 2 * ((g.W + 31) / 32) * ((g.H + kRasterBandRows - 1) / kRasterBandRows);
    const int raster_blocks = std::max(16, std::min(2048, (2 * pieces + (kRasterThreads >> 5) - 1) / (kRasterThreads >> 5)));
    ELASB_PREPARE_KERNEL(k_diffuse_raster);
    k_diffuse_raster<<<dim3(diffuse_blocks + raster_blocks, n_frames), kRasterThreads, 0, s>>>(
        g, subsampling, diffuse_blocks, hdr, scratch, scratch_next, grid1, grid2, lists1, lists2, tri1, tri2,
        reinterpret_cast<const int2*>(units1), reinterpret_cast<const int2*>(units2), map1, map2, tag_bits, st);
    count_launch();
}

}  // namespace elasb
