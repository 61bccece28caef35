// Mesh stage on the device: lattice filters + support list (k_lattice, one CTA per frame) and the
// Triangle-compatible Delaunay triangulation + scan-conversion work units (k_delaunay, one CTA per frame and
// image).  The algorithms are the phases of mesh_core.h; this file adds what a CTA needs around them:
// barriers, prefix sums, a bitonic sort, and the choice between shared memory and a global scratch area.
//
// Both kernels are latency-bound single-CTA programs (tens of microseconds on < 200 KB): they exist so that a
// frame never leaves the GPU between the support search and the dense matching -- no host round trip, no CPU
// work per frame -- and they overlap with the bandwidth-heavy kernels of other frame groups.
//
// Reference: elas.cpp:174-279, :496-517 (k_lattice); :534-600 with triangle.cpp:5446-6217, :7800-7853 (k_delaunay).
#include <algorithm>

#include "common.cuh"
#include "mesh_core.h"

namespace elasb {
namespace {

constexpr int kMeshThreads = 1024;      // k_lattice
constexpr int kDelaunayThreads = 512;   // k_delaunay: the parallel parts are small, the merges are single threads

// in-place inclusive prefix sum of data[0..m) by the whole CTA; warp_sums = 32 ints of shared memory
__device__ void block_scan_inclusive(int32_t* data, int m, int32_t* warp_sums)
{
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();
    const int chunk = (m + T - 1) / T;
    const int lo = min(tid * chunk, m), hi = min(lo + chunk, m);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += data[i];
    int incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = lane < (T >> 5) ? warp_sums[lane] : 0;
        int wi = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, wi, off);
            if (lane >= off) wi += v;
        }
        warp_sums[lane] = wi - w;                    // exclusive
    }
    __syncthreads();
    int run = warp_sums[warp] + incl - sum;
    for (int i = lo; i < hi; i++) { run += data[i]; data[i] = run; }
    __syncthreads();
}

// ascending bitonic sort of npad (a power of two) 64-bit keys by the whole CTA
__device__ void block_bitonic_sort(unsigned long long* k, int npad)
{
    for (int size = 2; size <= npad; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < (npad >> 1); i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = k[lo], b = k[hi];
                if ((a > b) == up) { k[lo] = b; k[hi] = a; }
            }
        }
    __syncthreads();
}

// atomic bit operation on one 16-bit lattice element of shared memory, through its 32-bit word
__device__ __forceinline__ void atomic_or16(int16_t* p, int bits)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    atomicOr(reinterpret_cast<unsigned*>(a & ~(uintptr_t)3), ((unsigned)bits & 0xFFFFu) << ((a & 2) * 8));
}

// count0 of every cell of the unfiltered lattice (mesh::incon_count0): the one part of the inconsistency filter
// that is proportional to cells x window, spread over the whole GPU (one thread per cell) instead of one CTA
__global__ void __launch_bounds__(256)
k_lattice_count(int Wc, int Hc, int win, int thr, const int16_t* __restrict__ dcan_raw, int32_t* __restrict__ work,
                size_t dcan_stride, size_t work_stride)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= Wc * Hc) return;
    const int vc = i / Wc, uc = i - vc * Wc;
    work[blockIdx.y * work_stride + i] = mesh::incon_count0(dcan_raw + blockIdx.y * dcan_stride, Wc, Hc, uc, vc, win, thr);
}

struct LatticeArgs {
    int Wc, Hc, step, win, thr, need;
    const int16_t* dcan_raw;     // K2's lattice, [frames][Hc][Wc]
    int16_t* dcan;               // after all three filters
    int16_t* dcan_incon;         // after the inconsistency filter only (stage dump; may be null)
    int32_t* support;            // [frames][support_cap][3]
    int32_t* work;               // [frames][3][cells]: count0 (k_lattice_count), two work lists of the propagation
    FrameHeader* hdr;
    size_t dcan_stride, support_stride, work_stride;
    int cnt_in_smem;
    int lat_in_smem;             // 0: the padded working copy of a big lattice lives behind the work area (global memory, L2)
};

__global__ void __launch_bounds__(kMeshThreads)
k_lattice(const LatticeArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ int32_t warp_sums[32];
    const int f = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int cells = a.Wc * a.Hc;
    int32_t* cnt = a.work + (size_t)f * a.work_stride;
    int32_t* lists[2] = {cnt + cells, cnt + 2 * cells};
    mesh::Lattice L{a.Wc, a.Hc, a.Wc + 2 * mesh::kPadC, a.step,
                    a.lat_in_smem ? reinterpret_cast<int16_t*>(smem_raw) : reinterpret_cast<int16_t*>(cnt + 3 * (size_t)cells)};
    const size_t lat_bytes = a.lat_in_smem ? ((size_t)mesh::lat_elems(a.Wc, a.Hc) * 2 + 15) & ~(size_t)15 : 0;
    int32_t* col = reinterpret_cast<int32_t*>(smem_raw + lat_bytes);         // [Wc + 1]
    __shared__ int n_list[2];
    const int16_t* raw = a.dcan_raw + (size_t)f * a.dcan_stride;
    int16_t* out = a.dcan + (size_t)f * a.dcan_stride;
    int32_t* support = a.support + (size_t)f * a.support_stride;
    if (a.cnt_in_smem) {
        // the supporter counts take thousands of dependent atomic decrements: in shared memory when they fit
        int32_t* scnt = reinterpret_cast<int32_t*>(smem_raw + lat_bytes + (((size_t)a.Wc + 1) * 4 + 15 & ~(size_t)15));
        for (int i = tid; i < cells; i += T) scnt[i] = cnt[i];
        cnt = scnt;
    }

    mesh::lattice_load(L, raw, tid, T);
    if (tid == 0) n_list[0] = n_list[1] = 0;
    __syncthreads();
    // ---- inconsistency filter by propagation (see mesh_core.h) ---------------------------------------------------
    auto atomic_add = [](int* p, int v) { return atomicAdd(p, v); };
    auto or16 = [](int16_t* p, int bits) { atomic_or16(p, bits); };
    mesh::incon_seed(L, cnt, a.need, lists[0], &n_list[0], tid, T, atomic_add);
    for (int cur = 0;; cur ^= 1) {
        __syncthreads();                                 // list `cur` is complete
        const int n = n_list[cur];
        if (n == 0) break;
        mesh::incon_propagate(L, cnt, a.win, a.thr, a.need, lists[cur], n, lists[cur ^ 1], &n_list[cur ^ 1], tid, T, atomic_add, or16);
        __syncthreads();
        if (tid == 0) n_list[cur] = 0;                   // consumed: it is the list after next
    }
    mesh::incon_finish(L, a.dcan_incon ? a.dcan_incon + (size_t)f * a.dcan_stride : nullptr, tid, T);
    __syncthreads();
    mesh::redundant_pass(L, true, tid, T);          // elas.cpp:501
    __syncthreads();
    mesh::redundant_pass(L, false, tid, T);         // elas.cpp:502
    __syncthreads();
    mesh::support_count(L, col + 1, tid, T);
    if (tid == 0) col[0] = 0;
    block_scan_inclusive(col + 1, a.Wc, warp_sums);  // col[uc] = survivors in columns < uc
    mesh::support_write(L, col, support, tid, T);
    mesh::lattice_store(L, out, tid, T);
    if (tid == 0) {
        FrameHeader h{};
        h.n_support = col[a.Wc];
        a.hdr[f] = h;
    }
}

struct DelaunayArgs {
    int W, H, unit_cap, smem_ints;
    const int32_t* support;
    int32_t* tri[2];             // [frames][tri_cap][3]
    int32_t* units[2];           // [frames][unit_cap][2]
    FrameHeader* hdr;
    int32_t* scratch;            // [frames][2][mesh_scratch]: used when a triangulation does not fit shared memory
    size_t support_stride, tri_stride, units_stride, scratch_stride;
};

// mem = 22 n ints of working memory.  Inlined twice by the kernel, once with shared memory (the compiler then
// addresses it with LDS/STS: the merges are chains of dependent loads, their latency is the kernel's run time)
// and once with the global scratch area.
__device__ __forceinline__ void delaunay_body(const DelaunayArgs& a, int32_t* mem, FrameHeader* hdr, int n, int img, int f,
                                              int32_t* warp_sums)
{
    const int tid = threadIdx.x, T = blockDim.x;
    int32_t* x = mem; int32_t* y = mem + n; int32_t* s = mem + 2 * n; int32_t* hull = mem + 3 * n;
    int32_t* nbr = mem + 5 * n; int32_t* vtx = mem + 13 * n;               // 4 ints per triangle, 2n triangles each
    uint32_t* xy = reinterpret_cast<uint32_t*>(mem + 21 * n);
    const int32_t* support = a.support + (size_t)f * a.support_stride;
    for (int i = tid; i < n; i += T) {                                     // elas.cpp:543-559
        x[i] = img ? support[3 * i] - support[3 * i + 2] : support[3 * i];
        y[i] = support[3 * i + 1];
        xy[i] = ((uint32_t)x[i] << 16) | (uint32_t)y[i];
    }
    // ---- the two sorted id lists: (x,y) and (y,x), keys are unique (no duplicate points on this path) -----
    int32_t* R = nbr;                                                      // 16 n ints free until the triangulation starts
    mesh::Order o{n, s, R, R + n, R + 2 * n, R + 3 * n, R + 4 * n, R + 5 * n, R + 6 * n, R + 7 * n};
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(mem + ((5 * n + 8 * n + 1) & ~1));
    int npad = 1;
    while (npad < n) npad <<= 1;
    __syncthreads();
    int dup = 0;
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 0 && img == 0) {
            // the support list is in u-outer / v-inner order (elas.cpp:505-517): the left image's points are sorted by (x,y)
            for (int i = tid; i < n; i += T) o.xs[i] = i;
            continue;
        }
        for (int i = tid; i < npad; i += T)
            keys[i] = i < n ? ((unsigned long long)(pass ? ((uint32_t)y[i] << 16) | (uint32_t)x[i] : ((uint32_t)x[i] << 16) | (uint32_t)y[i]) << 32) | (uint32_t)i
                            : ~0ull;
        block_bitonic_sort(keys, npad);
        for (int i = tid; i < n; i += T) {
            (pass ? o.ys : o.xs)[i] = (int32_t)(uint32_t)keys[i];
            if (pass == 0 && i > 0 && (keys[i] >> 32) == (keys[i - 1] >> 32)) dup = 1;
        }
        __syncthreads();
    }
    if (__syncthreads_or(dup)) {
        // two support points on one pixel of this image: Triangle keeps whichever its randomised sort meets first
        // (triangle.cpp:6180-6195); contexts whose parameters allow that use the host stage instead
        if (tid == 0) { hdr->n_tri[img] = 0; hdr->n_units[img] = 0; hdr->ovf_from[img] = 0; hdr->status = ELAS_B200_E_UNSUPPORTED; }
        return;
    }
    // ---- alternating-cut vertex order (triangle.cpp:6197-6206) ---------------------------------------------
    mesh::order_init(o, tid, T);
    __syncthreads();
    for (int axis = 0;; axis ^= 1) {
        if (!__syncthreads_or(mesh::order_flags(o, axis, tid, T))) break;
        for (int i = tid; i < n; i += T) o.scan[i] = o.side[i];
        block_scan_inclusive(o.scan, n, warp_sums);
        mesh::order_scatter(o, axis, tid, T);
        __syncthreads();
        mesh::order_commit(o, axis, tid, T);
        __syncthreads();
    }
    // ---- divide and conquer, deepest level first ---------------------------------------------------------------
    mesh::Mesh m{n, xy, s, nbr, vtx, hull};
    for (int depth = mesh::tree_depth(n); depth >= 0; depth--) {
        __syncthreads();
        mesh::triangulate_depth(m, depth, tid, T);
    }
    __syncthreads();
    // ---- elements in pool order without the ghosts (triangle.cpp:7834-7853) ----------------------------------------
    const int pool = 2 * n - 2;
    int32_t* flag = nbr; int32_t* fscan = nbr + 2 * n; int32_t* count = nbr + 4 * n;     // the links are dead now
    mesh::real_flags(m, flag, tid, T);
    __syncthreads();
    for (int i = tid; i < pool; i += T) fscan[i] = flag[i];
    block_scan_inclusive(fscan, pool, warp_sums);
    const int nt = fscan[pool - 1];
    int32_t* tri = a.tri[img] + (size_t)f * a.tri_stride;
    mesh::write_triangles(m, flag, fscan, tri, tid, T);
    __syncthreads();
    // ---- scan-conversion work units ----------------------------------------------------------------------------------
    mesh::unit_counts(x, y, tri, nt, a.W, a.H, kRasterBandRows, count, tid, T);
    __syncthreads();
    int32_t* cscan = flag;
    for (int i = tid; i < nt; i += T) cscan[i] = count[i];
    block_scan_inclusive(cscan, nt, warp_sums);
    mesh::write_units(x, y, tri, nt, a.W, a.H, kRasterBandRows, img, count, cscan, a.unit_cap,
                      a.units[img] + (size_t)f * a.units_stride, tid, T);
    // the prefix sums are monotone: the triangles that do not fit are a suffix [ovf_from, nt)
    if (tid == 0) { hdr->n_tri[img] = nt; if (nt == 0 || cscan[0] > a.unit_cap) { hdr->n_units[img] = 0; hdr->ovf_from[img] = 0; } }
    for (int t = tid; t < nt; t += T)
        if (cscan[t] <= a.unit_cap && (t == nt - 1 || cscan[t + 1] > a.unit_cap)) { hdr->n_units[img] = cscan[t]; hdr->ovf_from[img] = t + 1; }
}


__global__ void __launch_bounds__(kDelaunayThreads)
k_delaunay(const DelaunayArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ int32_t warp_sums[32];
    const int img = blockIdx.x, f = blockIdx.y;
    FrameHeader* hdr = a.hdr + f;
    const int n = hdr->n_support;
    if (n < 3) {                                                           // elas.cpp:69-75
        if (threadIdx.x == 0) { hdr->n_tri[img] = 0; hdr->n_units[img] = 0; hdr->ovf_from[img] = 0; }
        return;
    }
    if (22 * n + 2 <= a.smem_ints) delaunay_body(a, reinterpret_cast<int32_t*>(smem_raw), hdr, n, img, f, warp_sums);
    else delaunay_body(a, a.scratch + ((size_t)f * 2 + img) * a.scratch_stride, hdr, n, img, f, warp_sums);
}

}  // namespace

size_t lattice_smem_bytes(const FrameGeom& g)
{
    return (((size_t)mesh::lat_elems(g.Wc, g.Hc) * 2 + 15) & ~(size_t)15) + ((size_t)g.Wc + 1) * 4;
}

// ints of k_lattice's work area per frame: supporter counts, two work lists, and room for the padded lattice
size_t lattice_work_ints(const FrameGeom& g)
{
    return 3 * (size_t)g.Wc * g.Hc + ((size_t)mesh::lat_elems(g.Wc, g.Hc) + 1) / 2 + 4;
}

// parameters / sizes the device mesh stage handles; everything else takes the host stage (host_stage.cc)
bool mesh_on_device(const FrameGeom& g, const elas_b200_params& p)
{
    return !p.add_corners &&                                  // elas.cpp:283-318 appends points outside the lattice
           2 * p.lr_threshold < g.step &&                     // no two support points on one right-image pixel
           p.incon_window_size <= mesh::kPadC && p.incon_window_size >= 0 &&
           p.disp_max <= mesh::kValueMask && g.W < 16384 && g.H < 16384;
}

void launch_lattice(const FrameGeom& g, const elas_b200_params& p, const int16_t* dcan_raw, int16_t* dcan, int16_t* dcan_incon,
                    int32_t* support, int32_t* work, FrameHeader* hdr, const GroupStrides& st, int n_frames, cudaStream_t s)
{
    static unsigned long long optin = 0;
    if (ensure_dynamic_smem(k_lattice, 224 * 1024, &optin) != cudaSuccess) return;
    ELASB_PREPARE_KERNEL(k_lattice_count);
    k_lattice_count<<<dim3((g.Wc * g.Hc + 255) / 256, n_frames), 256, 0, s>>>(g.Wc, g.Hc, p.incon_window_size, p.incon_threshold,
                                                                             dcan_raw, work, st.dcan, st.lat_work);
    // shared memory: lattice + column table, plus the supporter counts if they fit as well; a lattice beyond that
    // (4096x2160: 820x432 cells) is worked on in global memory (it stays in L2)
    size_t base = lattice_smem_bytes(g);
    const bool lat_in_smem = base <= 200 * 1024;
    if (!lat_in_smem) base = ((size_t)g.Wc + 1) * 4;
    const size_t with_counts = ((base + 15) & ~(size_t)15) + (size_t)g.Wc * g.Hc * 4;
    const bool cnt_in_smem = lat_in_smem && with_counts <= 200 * 1024;
    LatticeArgs a{g.Wc, g.Hc, g.step, p.incon_window_size, p.incon_threshold, p.incon_min_support,
                  dcan_raw, dcan, dcan_incon, support, work, hdr, st.dcan, st.support, st.lat_work, cnt_in_smem ? 1 : 0,
                  lat_in_smem ? 1 : 0};
    k_lattice<<<n_frames, kMeshThreads, cnt_in_smem ? with_counts : base, s>>>(a);
    count_launch(2);
}

void launch_delaunay(const FrameGeom& g, const int32_t* support, int32_t* tri1, int32_t* tri2, int32_t* units1, int32_t* units2,
                     int unit_cap, FrameHeader* hdr, int32_t* scratch, const GroupStrides& st, int n_frames, cudaStream_t s)
{
    static unsigned long long optin = 0;
    constexpr int kSmemMax = 200 * 1024;
    if (ensure_dynamic_smem(k_delaunay, kSmemMax, &optin) != cudaSuccess) return;
    // A triangulation works in shared memory when its 22 n ints fit, else in the global scratch area.  Frames whose
    // lattice is so large that the typical support set (about a twelfth of the lattice cells at most) cannot fit
    // anyway are launched without the shared-memory reservation: their long-running CTAs then leave the SM's
    // shared memory to the kernels of other frame groups.
    const bool may_fit = (size_t)(g.Wc * g.Hc / 12) * 22 * 4 <= (size_t)kSmemMax;
    const int smem = may_fit ? kSmemMax : 1024;
    DelaunayArgs a{g.W, g.H, unit_cap, smem / 4, support, {tri1, tri2}, {units1, units2}, hdr, scratch,
                   st.support, st.tri, st.units, st.mesh_scratch};
    k_delaunay<<<dim3(2, n_frames), kDelaunayThreads, smem, s>>>(a);
    count_launch();
}

}  // namespace elasb
