// K8-K12: post-processing of the disparity maps (elas.cpp:1122-1838), one thread per pixel.
//
// All five reference stages are sequential scans written in place; each is restated here in a form
// whose per-pixel result depends only on the stage's INPUT, so that pixels can be computed
// independently (the notes at each kernel say why that is equivalent).  Float expressions use the
// explicit _rn intrinsics: the reference is x86-64 SSE code without FMA contraction.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace elasb {
namespace {

// ---------------------------------------------------------------------------------------------
// K8  left/right consistency check, elas.cpp:1122-1204.  The reference reads copies of D1/D2 and
// writes the originals; here input and output are separate buffers.
// ---------------------------------------------------------------------------------------------
__global__ void k_lr_check(int Dw, int Dh, int subsampling, float lr_threshold,
                           const float* __restrict__ D1, const float* __restrict__ D2,
                           float* __restrict__ O1, OutTable O2_tab, size_t D_stride)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= Dw || v >= Dh) return;
    D1 += blockIdx.z * D_stride; D2 += blockIdx.z * D_stride; O1 += blockIdx.z * D_stride;
    float* __restrict__ O2 = O2_tab.p[blockIdx.z];
    const size_t row = (size_t)v * Dw, a = row + u;
    const float d1 = D1[a], d2 = D2[a];
    const float w1 = subsampling ? __fsub_rn((float)u, __fmul_rn(d1, 0.5f)) : __fsub_rn((float)u, d1);  // :1152-1161
    const float w2 = subsampling ? __fadd_rn((float)u, __fmul_rn(d2, 0.5f)) : __fadd_rn((float)u, d2);
    float o1 = (float)kInvalid, o2 = (float)kInvalid;
    if (d1 >= 0.f && w1 >= 0.f && w1 < (float)Dw)                                                       // :1164-1179
        if (!(fabsf(__fsub_rn(D2[row + (int)w1], d1)) > lr_threshold)) o1 = d1;
    if (d2 >= 0.f && w2 >= 0.f && w2 < (float)Dw)                                                       // :1182-1197
        if (!(fabsf(__fsub_rn(D1[row + (int)w2], d2)) > lr_threshold)) o2 = d2;
    O1[a] = o1; O2[a] = o2;
}

// K8 with the rows staged in shared memory: a CTA owns one row of both maps.  The L/R check of a pixel only reads
// the other map within the same row (elas.cpp:1164-1197), at a data-dependent column: from shared memory those
// gathers cost nothing, from global memory they are uncoalesced.  The checked right map optionally leaves NARROWED for
// the trip across PCIe (exact: it holds raw integer disparities or -10): as int16, or -- disp_max <= 255 -- as one
// byte per pixel plus a validity bit per pixel (rows of 32-bit ballot words), 1.125 instead of 4 bytes per pixel.
__global__ void __launch_bounds__(256)
k_lr_rows(int Dw, int subsampling, float lr_threshold,
          const float* __restrict__ D1, const float* __restrict__ D2,
          float* __restrict__ O1, OutTable O2_tab, NarrowD2 narrow, size_t D_stride)
{
    extern __shared__ float s_rows[];          // [2][Dw]: raw D1 row, raw D2 row
    float* r1 = s_rows; float* r2 = s_rows + Dw;
    const int v = blockIdx.x;
    // blockIdx.y = frame of the group
    D1 += blockIdx.y * D_stride; D2 += blockIdx.y * D_stride; O1 += blockIdx.y * D_stride;
    uint8_t* nbase = narrow.mode ? static_cast<uint8_t*>(narrow.base) + blockIdx.y * narrow.stride_bytes : nullptr;
    int16_t* __restrict__ O2_i16 = reinterpret_cast<int16_t*>(nbase);
    uint32_t* __restrict__ O2_mask = reinterpret_cast<uint32_t*>(nbase + narrow.mask_offset) + (size_t)v * narrow.mask_words_per_row;
    float* __restrict__ O2 = O2_tab.p[blockIdx.y];
    const size_t row = (size_t)v * Dw;
    for (int u = threadIdx.x; u < Dw; u += 256) { r1[u] = D1[row + u]; r2[u] = D2[row + u]; }
    __syncthreads();
    for (int u0 = 0; u0 < Dw; u0 += 256) {                 // whole warps stay in the loop (the validity ballot)
        const int u = u0 + threadIdx.x;
        const bool in = u < Dw;
        float o1 = (float)kInvalid, o2 = (float)kInvalid;
        if (in) {
            const float d1 = r1[u], d2 = r2[u];
            const float w1 = subsampling ? __fsub_rn((float)u, __fmul_rn(d1, 0.5f)) : __fsub_rn((float)u, d1);  // :1152-1161
            const float w2 = subsampling ? __fadd_rn((float)u, __fmul_rn(d2, 0.5f)) : __fadd_rn((float)u, d2);
            if (d1 >= 0.f && w1 >= 0.f && w1 < (float)Dw)                                                       // :1164-1179
                if (!(fabsf(__fsub_rn(r2[(int)w1], d1)) > lr_threshold)) o1 = d1;
            if (d2 >= 0.f && w2 >= 0.f && w2 < (float)Dw)                                                       // :1182-1197
                if (!(fabsf(__fsub_rn(r1[(int)w2], d2)) > lr_threshold)) o2 = d2;
            O1[row + u] = o1;
        }
        if (narrow.mode == 2) {
            const unsigned valid = __ballot_sync(0xffffffffu, in && o2 >= 0.f);
            if (in) nbase[row + u] = o2 >= 0.f ? (uint8_t)o2 : (uint8_t)0;
            if ((threadIdx.x & 31) == 0 && in) O2_mask[u >> 5] = valid;
        } else if (in) {
            if (narrow.mode == 1) O2_i16[row + u] = (int16_t)o2; else O2[row + u] = o2;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K9  speckle removal, elas.cpp:1208-1326.  The reference flood-fills 4-connected segments in
// which neighbouring valid pixels differ by <= speckle_sim_threshold and invalidates segments with
// fewer than speckle_size pixels.  Segments are the connected components of a symmetric relation,
// so the result does not depend on traversal order.  Only "is my component smaller than
// speckle_size" is ever asked, which allows a tile-local formulation:
//   tiles:   every 64x32 tile labels its pixels with a union-find in SHARED memory and counts the local
//            components.  A local component of >= speckle_size pixels is large whatever lies beyond the tile
//            (label KEEP); a smaller one that touches no neighbouring tile is final (label DROP).  Only the
//            small components on tile borders stay open: they become nodes (tile, local root) in global memory.
//   border:  connected pixel pairs across tile borders: node-node pairs are united (global union-find on the
//            few open nodes), a node next to a KEEP pixel is flagged as attached to a large component.
//   sizes:   every open node adds its pixel count (+ speckle_size if flagged) to its root.
//   apply:   KEEP stays, DROP goes, an open node's pixels stay iff the root's total reaches speckle_size
//            (k_post_fused, or k_seg_apply for settings outside the fused path).
// Pointer chasing through global memory is limited to the open nodes, a few percent of the pixels.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool seg_conn(float a, float b, float thr)
{
    return a >= 0.f && b >= 0.f && fabsf(__fsub_rn(a, b)) <= thr;                    // :1281, :1285
}

constexpr int kSegTW = 64, kSegTH = 32, kSegTile = kSegTW * kSegTH, kSegThreads = 256;
constexpr int kSegKeep = -2, kSegDrop = -1;      // pixel labels besides open node ids (>= 0); invalid pixels are DROP too

// find with path halving (shared memory; concurrent writers only ever replace a parent by an ancestor)
__device__ __forceinline__ int uf_find_s(int* L, int x)
{
    int p = L[x];
    while (p != x) {
        const int gp = L[p];
        if (gp != p) L[x] = gp;
        x = p; p = gp;
    }
    return x;
}
// union by minimum index with atomicMin (Playne & Hawick): a root is hooked under the smaller root; if another
// thread hooked it meanwhile, the union continues with what it was hooked to
__device__ __forceinline__ void uf_union_s(int* L, int a, int b)
{
    for (;;) {
        a = uf_find_s(L, a);
        b = uf_find_s(L, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;
    }
}

struct SegArgs {
    int Dw, Dh, tiles_x, tiles_y, speckle;
    float thr;
    const float* D;          // frames D_stride apart
    int32_t* label;          // per pixel: kSegKeep, kSegDrop or an open node id
    int32_t* nodes;          // per frame [3][node_cap]: parent, pixel count (roots: total), attached-to-large flag
    size_t D_stride, nodes_stride;
    int node_cap;            // tiles * kSegTile
};

// One 64x32 tile per CTA, 8 warps, warp w owns tile rows 4w..4w+3.
//   1. horizontal runs: every pixel points at the first pixel of its run (ballots, no atomics)
//   2. vertical unions between runs (a pair is skipped when the pair to its left joins the same two runs)
//   3. every pixel -> root; pixels per root; does the component touch a neighbouring tile
__global__ void __launch_bounds__(kSegThreads)
k_seg_tiles(const SegArgs a)
{
    __shared__ float sD[kSegTile];
    __shared__ int sL[kSegTile];          // union-find parent (local pixel index), -1 for invalid pixels
    __shared__ int sN[kSegTile];          // pixels per root; bit 30 = the component touches a neighbouring tile
    const int f = blockIdx.z, x0 = blockIdx.x * kSegTW, y0 = blockIdx.y * kSegTH, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ D = a.D + f * a.D_stride;
    int32_t* __restrict__ label = a.label + f * a.D_stride;
    int32_t* __restrict__ node_parent = a.nodes + f * a.nodes_stride;
    int32_t* __restrict__ node_count = node_parent + a.node_cap;
    int32_t* __restrict__ node_flag = node_count + a.node_cap;
    const int tile = blockIdx.y * a.tiles_x + blockIdx.x;
    // 1. rows 4 warp .. 4 warp + 3, two 32-pixel halves each
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int ly = 4 * warp + r, y = y0 + ly;
        int carry = -1;                    // run that reaches the end of the left half
        float prev_last = -1.f;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int lx = 32 * h + lane, x = x0 + lx, p = ly * kSegTW + lx;
            const float d = (x < a.Dw && y < a.Dh) ? D[(size_t)y * a.Dw + x] : -1.f;
            float dl = __shfl_up_sync(0xffffffffu, d, 1);
            if (lane == 0) dl = prev_last;
            const bool valid = d >= 0.f;
            const bool start = valid && !seg_conn(dl, d, a.thr);
            const unsigned starts = __ballot_sync(0xffffffffu, start);
            const unsigned upto = starts & (0xffffffffu >> (31 - lane));
            const int run = upto ? ly * kSegTW + 32 * h + 31 - __clz(upto) : carry;     // a run never spans an invalid pixel
            sD[p] = d;
            sL[p] = valid ? run : -1;
            sN[p] = 0;
            prev_last = __shfl_sync(0xffffffffu, d, 31);
            carry = __shfl_sync(0xffffffffu, valid ? run : -1, 31);
        }
    }
    __syncthreads();
    // 2. vertical unions
#pragma unroll
    for (int k = 0; k < kSegTile / kSegThreads; k++) {
        const int p = tid + k * kSegThreads;
        if (p < kSegTW) continue;
        const float d = sD[p], du = sD[p - kSegTW];
        if (!seg_conn(du, d, a.thr)) continue;
        if (p & (kSegTW - 1)) {
            const float dl = sD[p - 1], dul = sD[p - kSegTW - 1];
            if (seg_conn(dl, d, a.thr) && seg_conn(dul, du, a.thr) && seg_conn(dul, dl, a.thr)) continue;
        }
        uf_union_s(sL, sL[p], sL[p - kSegTW]);
    }
    __syncthreads();
    // 3. roots, counts, border flag
    int root[kSegTile / kSegThreads];
#pragma unroll
    for (int k = 0; k < kSegTile / kSegThreads; k++) {
        const int p = tid + k * kSegThreads;
        root[k] = sL[p] >= 0 ? uf_find_s(sL, sL[p]) : -1;
        if (root[k] < 0) continue;
        const int lx = p & (kSegTW - 1), ly = p / kSegTW, x = x0 + lx, y = y0 + ly;
        // on a tile edge that has a neighbouring tile behind it
        const bool edge = (lx == 0 && x > 0) || (lx == kSegTW - 1 && x + 1 < a.Dw) || (ly == 0 && y > 0) || (ly == kSegTH - 1 && y + 1 < a.Dh);
        // one add per warp and root where neighbouring lanes share it
        const unsigned same = __match_any_sync(__activemask(), root[k]);
        if (lane == __ffs(same) - 1) atomicAdd(&sN[root[k]], __popc(same));      // counts stay below 2^12
        if (edge) atomicOr(&sN[root[k]], 1 << 30);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSegTile / kSegThreads; k++) {
        const int p = tid + k * kSegThreads, x = x0 + (p & (kSegTW - 1)), y = y0 + (p / kSegTW);
        if (x >= a.Dw || y >= a.Dh) continue;
        int l = kSegDrop;
        if (root[k] >= 0) {
            const int n = sN[root[k]], count = n & 0xFFF, open = n >> 30;
            if (count >= a.speckle) l = kSegKeep;
            else if (open) {
                l = tile * kSegTile + root[k];
                if (root[k] == p) { node_parent[l] = l; node_count[l] = count; node_flag[l] = 0; }
            }
        }
        label[(size_t)y * a.Dw + x] = l;
    }
}

// global union-find on the open nodes (few, short chains): hook the larger index under the smaller
__device__ __forceinline__ int uf_find_g(int32_t* parent, int x)
{
    int p = __ldcg(parent + x);
    while (p != x) { x = p; p = __ldcg(parent + x); }
    return x;
}
__device__ __forceinline__ void uf_union_g(int32_t* parent, int a, int b)
{
    for (;;) {
        a = uf_find_g(parent, a);
        b = uf_find_g(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(parent + a, b);
        if (old == a) return;
        a = old;
    }
}

// one thread per pixel pair across a tile border: items [0, nv) = vertical borders (x = 64k: pixels (x-1,y),(x,y)),
// items [nv, nv + nh) = horizontal borders (y = 32k: pixels (x,y-1),(x,y))
__global__ void __launch_bounds__(256)
k_seg_border(const SegArgs a)
{
    const int f = blockIdx.y;
    const float* __restrict__ D = a.D + f * a.D_stride;
    const int32_t* __restrict__ label = a.label + f * a.D_stride;
    int32_t* node_parent = a.nodes + f * a.nodes_stride;
    int32_t* node_flag = node_parent + 2 * a.node_cap;
    const int nv = (a.tiles_x - 1) * a.Dh, nh = (a.tiles_y - 1) * a.Dw;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < nv + nh; i += gridDim.x * 256) {
        int pa, pb;
        if (i < nv) { const int bx = i / a.Dh, y = i - bx * a.Dh; pb = y * a.Dw + (bx + 1) * kSegTW; pa = pb - 1; }
        else { const int j = i - nv, by = j / a.Dw, x = j - by * a.Dw; pb = (by + 1) * kSegTH * a.Dw + x; pa = pb - a.Dw; }
        if (!seg_conn(D[pa], D[pb], a.thr)) continue;
        const int la = label[pa], lb = label[pb];
        if (la >= 0 && lb >= 0) uf_union_g(node_parent, la, lb);
        else if (la >= 0 && lb == kSegKeep) node_flag[la] = 1;
        else if (lb >= 0 && la == kSegKeep) node_flag[lb] = 1;
    }
}

// one thread per pixel: the root pixel of every open local component adds the component to its global root
__global__ void __launch_bounds__(256)
k_seg_sizes(const SegArgs a)
{
    const int f = blockIdx.y;
    const int32_t* __restrict__ label = a.label + f * a.D_stride;
    int32_t* node_parent = a.nodes + f * a.nodes_stride;
    int32_t* node_count = node_parent + a.node_cap;
    const int32_t* node_flag = node_count + a.node_cap;
    const int n = a.Dw * a.Dh;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const int l = label[i];
        if (l < 0) continue;
        // is this pixel the root pixel of its local component?  node id = tile * kSegTile + local index
        const int y = i / a.Dw, x = i - y * a.Dw;
        const int tile = (y / kSegTH) * a.tiles_x + x / kSegTW, local = (y % kSegTH) * kSegTW + (x % kSegTW);
        if (l != tile * kSegTile + local) continue;
        const int root = uf_find_g(node_parent, l);
        const int add = (root != l ? __ldcg(node_count + l) : 0) + (node_flag[l] ? a.speckle : 0);
        if (add) atomicAdd(node_count + root, add);
    }
}

// does the pixel with this label survive speckle removal?  (open nodes: the root's total decides)
__device__ __forceinline__ bool seg_keeps(int l, const int32_t* __restrict__ node_parent, const int32_t* __restrict__ node_count, int speckle)
{
    if (l < 0) return l == kSegKeep;
    int r = l, p;
    while ((p = node_parent[r]) != r) r = p;
    return node_count[r] >= speckle;
}

__global__ void k_seg_apply(int n, int speckle, float* __restrict__ D, const int32_t* __restrict__ label,
                            const int32_t* __restrict__ nodes, int node_cap, size_t D_stride, size_t nodes_stride)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    D += blockIdx.y * D_stride; label += blockIdx.y * D_stride; nodes += blockIdx.y * nodes_stride;
    // an invalid pixel is a segment of one (:1248-1250): it is (re)written as invalid whenever speckle_size > 1
    const bool valid = D[i] >= 0.f;
    if (valid ? !seg_keeps(label[i], nodes, nodes + node_cap, speckle) : 1 < speckle) D[i] = (float)kInvalid;   // :1309-1317
}

// ---------------------------------------------------------------------------------------------
// K10 gap interpolation, elas.cpp:1330-1530.  Row pass then column pass.  Within a pass the
// reference only ever reads pixels that were valid in the pass's input (the run's two bounding
// pixels), so each invalid pixel can find its own run: nearest valid neighbour on either side along
// the line, run length = distance between them - 1 <= ipol_gap_width, and the run must not touch
// the line's ends (:1374, :1463).  With add_corners the pass then extends the first/last valid pixel
// of the line outwards by up to ipol_gap_width pixels (:1401-1436, :1493-1528).
// ---------------------------------------------------------------------------------------------
__global__ void k_gap_pass(int Dw, int Dh, int gap, int add_corners, int vertical,
                           const float* __restrict__ in, float* __restrict__ out, size_t in_stride, size_t out_stride)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= Dw || v >= Dh) return;
    in += blockIdx.z * in_stride; out += blockIdx.z * out_stride;
    const int len = vertical ? Dh : Dw, pos = vertical ? v : u;
    const ptrdiff_t stride = vertical ? Dw : 1;
    const float* line = in + (vertical ? (size_t)u : (size_t)v * Dw);
    const float d = line[pos * stride];
    float o = d;
    if (!(d >= 0.f)) {
        int l = pos - 1, r = pos + 1;
        const int reach = min(gap, len);
        while (l >= 0 && pos - l <= reach && !(line[l * stride] >= 0.f)) l--;
        while (r < len && r - pos <= reach && !(line[r * stride] >= 0.f)) r++;
        const bool lv = l >= 0 && pos - l <= reach, rv = r < len && r - pos <= reach;
        if (lv && rv && r - l - 1 <= gap) {
            const float d1 = line[l * stride], d2 = line[r * stride];
            o = fabsf(__fsub_rn(d1, d2)) < 3.0f ? __fmul_rn(__fadd_rn(d1, d2), 0.5f) : fminf(d1, d2);   // :1379-1380
        } else if (add_corners) {
            // extrapolation: pos lies before the first / after the last valid pixel of the line
            if (rv && !lv) {            // is everything left of pos invalid?
                int k = pos - 1; while (k >= 0 && !(line[k * stride] >= 0.f)) k--;
                if (k < 0) o = line[r * stride];
            } else if (lv && !rv) {
                int k = pos + 1; while (k < len && !(line[k * stride] >= 0.f)) k++;
                if (k >= len) o = line[l * stride];
            }
        }
    }
    out[(size_t)v * Dw + u] = o;
}

// ---------------------------------------------------------------------------------------------
// K11 "adaptive mean", elas.cpp:1535-1754.  8 taps (4 with subsampling) along the line; the window
// of centre c is [c-4, c+3] ([c-2, c+1]); tap weight = max(0, 4 - M(x - x_c)) where M() is the
// reference's mis-built abs mask: a bitwise AND with 0x4F000000, i.e. weights 4/2/0.
// The reference keeps the window in a ring indexed by (position % taps) and sums SSE lanes
// lane k = slot k + slot k+4, then ((l0+l1)+l2)+l3 -- reproduced so the float sums are identical.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float masked_abs(float x)
{
    return __uint_as_float(__float_as_uint(x) & 0x4F000000u);
}

// ((q[(0-r)&3] + q[(1-r)&3]) + q[(2-r)&3]) + q[(3-r)&3]: the reference's lane order for a ring that
// starts r slots in.  The rotation is two conditional swaps stages (by 1, by 2) on registers: neighbouring
// pixels of a row have different r, a switch would make the warp walk all four orders.
__device__ __forceinline__ float ring_sum4(float q0, float q1, float q2, float q3, int r)
{
    const bool r1 = r & 1, r2 = r & 2;
    // rotate right by 1: (q0,q1,q2,q3) -> (q3,q0,q1,q2)
    const float a0 = r1 ? q3 : q0, a1 = r1 ? q0 : q1, a2 = r1 ? q1 : q2, a3 = r1 ? q2 : q3;
    // rotate right by 2
    const float b0 = r2 ? a2 : a0, b1 = r2 ? a3 : a1, b2 = r2 ? a0 : a2, b3 = r2 ? a1 : a3;
    return __fadd_rn(__fadd_rn(__fadd_rn(b0, b1), b2), b3);
}

template <int TAPS>
__device__ __forceinline__ bool mean_window(const float* __restrict__ line, ptrdiff_t stride, int c, float* result)
{
    constexpr int BACK = TAPS == 8 ? 4 : 2;          // window = [c-BACK, c+TAPS-BACK-1]
    const float xc = line[c * stride];
    float w[TAPS], f[TAPS];                          // by tap; tap k sits in ring slot (c-BACK+k) % TAPS (:1667, :1590)
#pragma unroll
    for (int k = 0; k < TAPS; k++) {
        const float x = line[(c - BACK + k) * stride];
        w[k] = fmaxf(0.f, __fsub_rn(4.0f, masked_abs(__fsub_rn(x, xc))));
        f[k] = __fmul_rn(x, w[k]);
    }
    float ws, fs;
    if (TAPS == 8) {
        // SSE lane l = slot l + slot l+4 = taps j and j+4 with j = (l - first_slot) & 3
        const int r = (c - BACK) & 3;
        ws = ring_sum4(__fadd_rn(w[0], w[4]), __fadd_rn(w[1], w[5]), __fadd_rn(w[2], w[6]), __fadd_rn(w[3], w[7]), r);
        fs = ring_sum4(__fadd_rn(f[0], f[4]), __fadd_rn(f[1], f[5]), __fadd_rn(f[2], f[6]), __fadd_rn(f[3], f[7]), r);
    } else {
        const int r = (c - BACK) & 3;
        ws = ring_sum4(w[0], w[1], w[2], w[3], r);
        fs = ring_sum4(f[0], f[1], f[2], f[3], r);
    }
    if (ws > 0.f) {
        const float d = __fdiv_rn(fs, ws);
        if (d >= 0.f) { *result = d; return true; }
    }
    return false;
}

// horizontal: in = D with invalid -> -10 (the reference's D_copy), out = D_tmp (initialised to in)
// vertical:   in = D_tmp, out = D (keeps its value where the window gives nothing)
template <int TAPS>
__global__ void k_mean_pass(int Dw, int Dh, int vertical, const float* __restrict__ in,
                            const float* keep, float* out,      // keep may alias out (vertical pass)
                            size_t in_stride, size_t keep_stride, size_t out_stride)
{
    constexpr int BACK = TAPS == 8 ? 4 : 2, FWD = TAPS - BACK - 1;
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= Dw || v >= Dh) return;
    in += blockIdx.z * in_stride; keep += blockIdx.z * keep_stride; out += blockIdx.z * out_stride;
    const size_t a = (size_t)v * Dw + u;
    float o = keep[a];
    if (!vertical) {
        // rows 3..Dh-4, centres c = u' - lag for u' in [TAPS-1, Dw)  (:1654-1663, :1577-1586)
        if (v >= 3 && v < Dh - 3 && u >= BACK && u + FWD < Dw) {
            float r;
            if (mean_window<TAPS>(in + (size_t)v * Dw, 1, u, &r)) o = r;
        }
    } else {
        if (u >= 3 && u < Dw - 3 && v >= BACK && v + FWD < Dh) {
            float r;
            if (mean_window<TAPS>(in + u, Dw, v, &r)) o = r;
        }
    }
    out[a] = o;
}

// ---------------------------------------------------------------------------------------------
// K12 separable 7-tap median, elas.cpp:1758-1838 (MIDDLEBURY preset).  Horizontal pass into a
// zero-initialised temporary (calloc, :1770), vertical pass back into D; 3-pixel border untouched.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float median7(const float* __restrict__ p, ptrdiff_t stride)
{
    float vals[7];
#pragma unroll
    for (int j = 0; j < 7; j++) {
        const float t = p[(j - 3) * stride];
        int i = j - 1;
        while (i >= 0 && vals[i] > t) { vals[i + 1] = vals[i]; i--; }
        vals[i + 1] = t;
    }
    return vals[3];
}

__global__ void k_median_pass(int Dw, int Dh, int vertical, const float* D,
                              const float* in, float* out,        // D aliases in (horizontal) or out (vertical)
                              size_t D_stride, size_t in_stride, size_t out_stride)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= Dw || v >= Dh) return;
    D += blockIdx.z * D_stride; in += blockIdx.z * in_stride; out += blockIdx.z * out_stride;
    const size_t a = (size_t)v * Dw + u;
    const bool inner = u >= 3 && u < Dw - 3 && v >= 3 && v < Dh - 3;
    if (!vertical) {
        float o = 0.f;                                       // calloc'ed D_temp outside the inner region
        if (inner) o = D[a] >= 0.f ? median7(in + a, 1) : D[a];
        out[a] = o;
    } else {
        if (inner && D[a] >= 0.f) out[a] = median7(in + a, Dw);
    }
}

// ---------------------------------------------------------------------------------------------
// Fused tail of the post-processing chain: speckle "apply" (the last step of K9), K10 gap
// interpolation (row pass, column pass) and K11 adaptive mean (row pass, column pass) in ONE kernel.
// The five stages are local stencils of each other -- row pass reach <= kFuseGap columns, column pass
// reach <= kFuseGap rows, mean window [c-BACK, c+FWD] -- so a CTA computes an output tile from an
// input tile with (BACK+kFuseGap | FWD+kFuseGap) halos entirely in shared memory:
//   A  = D after speckle removal   rows [r0-BACK-G, r0+TH+FWD+G) x cols [c0-BACK-G, c0+TW+FWD+G)
//   B  = after the gap row pass    same rows                      x cols [c0-BACK,   c0+TW+FWD)
//   C  = after the gap column pass rows [r0-BACK,   r0+TH+FWD)    x same cols          (= K10 output)
//   M  = after the mean row pass   same rows                      x cols [c0, c0+TW)   (reuses A)
//   out= after the mean column pass rows [r0, r0+TH)              x cols [c0, c0+TW)
// Pixels outside the image enter as invalid (-10): a gap run that reaches them finds no bounding
// valid pixel, which is the reference's "run touches the line end" case (elas.cpp:1374, :1463), and
// the mean windows are only evaluated where the reference evaluates them (all taps inside the image).
// Per-pixel expressions are the ones of k_gap_pass / mean_window, so results are bit-identical to
// the unfused kernels.  Used when ipol_gap_width <= kFuseGap and add_corners is off (the ROBOTICS
// family); other settings run the unfused kernels.
// ---------------------------------------------------------------------------------------------
constexpr int kFuseGap = 3;
constexpr int kFuseTW = 64, kFuseTH = 32, kFuseThreads = 512;

struct FuseArgs {
    int Dw, Dh, gap, speckle, apply;     // apply: fold k_seg_apply in (parent/size valid)
    const float* in;                     // D after the L/R check (apply) or after speckle removal; frames D_stride apart
    const int32_t* label;                // speckle labels and open nodes (k_seg_*)
    const int32_t* nodes;
    int node_cap;
    size_t nodes_stride;
    OutTable out;                        // final map of every frame of the group (must not alias in)
    float* dump_seg;                     // optional stage dumps (tests, single frame): D after speckle removal, after gap interpolation
    float* dump_gap;
    size_t D_stride;
};

__device__ __forceinline__ float gap_fill(const float* __restrict__ line, int stride, int gap)
{
    // line points at the pixel; neighbours at +-j*stride.  Same decisions as k_gap_pass without add_corners:
    // the nearest valid pixel within `gap` on either side, run length (l + r - 1) <= gap.  Written without
    // loops or branches on the neighbours (gap <= kFuseGap = 3): almost every warp holds an invalid pixel,
    // so a data-dependent search would be walked by all 32 lanes anyway.
    const float d = line[0];
    if (d >= 0.f) return d;
    const float l1 = line[-stride], l2 = line[-2 * stride], l3 = line[-3 * stride];
    const float r1 = line[stride], r2 = line[2 * stride], r3 = line[3 * stride];
    const bool vl1 = l1 >= 0.f && gap >= 1, vl2 = l2 >= 0.f && gap >= 2, vl3 = l3 >= 0.f && gap >= 3;
    const bool vr1 = r1 >= 0.f && gap >= 1, vr2 = r2 >= 0.f && gap >= 2, vr3 = r3 >= 0.f && gap >= 3;
    const int l = vl1 ? 1 : vl2 ? 2 : vl3 ? 3 : 8, r = vr1 ? 1 : vr2 ? 2 : vr3 ? 3 : 8;
    if (l + r - 1 > gap) return d;                              // no bounding pair within reach (8 = none)
    const float d1 = vl1 ? l1 : vl2 ? l2 : l3, d2 = vr1 ? r1 : vr2 ? r2 : r3;
    return fabsf(__fsub_rn(d1, d2)) < 3.0f ? __fmul_rn(__fadd_rn(d1, d2), 0.5f) : fminf(d1, d2);   // :1379-1380
}

template <int TAPS, bool MEAN>
__global__ void __launch_bounds__(kFuseThreads)
k_post_fused(const FuseArgs a_)
{
    // blockIdx.z = frame of the group
    FuseArgs a = a_;
    a.in += blockIdx.z * a.D_stride;
    if (a.apply) { a.label += blockIdx.z * a.D_stride; a.nodes += blockIdx.z * a.nodes_stride; }
    float* __restrict__ out = a.out.p[blockIdx.z];
    constexpr int BACK = MEAN ? (TAPS == 8 ? 4 : 2) : 0, FWD = MEAN ? TAPS - BACK - 1 : 0, G = kFuseGap;
    constexpr int AW = kFuseTW + BACK + FWD + 2 * G, AH = kFuseTH + BACK + FWD + 2 * G;
    constexpr int BW = kFuseTW + BACK + FWD, BH = AH;
    constexpr int CW = BW, CH = kFuseTH + BACK + FWD;
    constexpr int MW = kFuseTW, MH = CH;
    __shared__ float sA[AH * AW];      // A, later M
    __shared__ float sB[BH * BW];
    __shared__ float sC[CH * CW];
    const int c0 = blockIdx.x * kFuseTW, r0 = blockIdx.y * kFuseTH;
    const int Dw = a.Dw, Dh = a.Dh;

    // ---- A: load (+ speckle apply, elas.cpp:1309-1317) ------------------------------------------
    // pixel -> run start -> root -> size is a chain of dependent L2 loads: four elements per thread are
    // walked together so that the chains overlap
    for (int i0 = threadIdx.x; i0 < AH * AW; i0 += 4 * kFuseThreads) {
        int idx[4], lab[4];
        float d[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = i0 + k * kFuseThreads;
            const int y = i / AW, x = i - y * AW;
            const int v = r0 - BACK - G + y, u = c0 - BACK - G + x;
            idx[k] = (i < AH * AW && v >= 0 && v < Dh && u >= 0 && u < Dw) ? v * Dw + u : -1;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) d[k] = idx[k] >= 0 ? a.in[idx[k]] : (float)kInvalid;
        if (a.apply) {
#pragma unroll
            for (int k = 0; k < 4; k++) lab[k] = d[k] >= 0.f ? a.label[idx[k]] : kSegDrop;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                // an invalid pixel is a segment of one (:1248-1250); 1 < speckle_size in every supported setting of this path
                if (!(d[k] >= 0.f)) { if (1 < a.speckle) d[k] = (float)kInvalid; }
                else if (!seg_keeps(lab[k], a.nodes, a.nodes + a.node_cap, a.speckle)) d[k] = (float)kInvalid;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = i0 + k * kFuseThreads;
            if (i >= AH * AW) continue;
            sA[i] = d[k];
            if (a.dump_seg && idx[k] >= 0) {
                const int y = i / AW, x = i - y * AW;
                const int v = r0 - BACK - G + y, u = c0 - BACK - G + x;
                if (v >= r0 && v < r0 + kFuseTH && u >= c0 && u < c0 + kFuseTW) a.dump_seg[idx[k]] = d[k];
            }
        }
    }
    __syncthreads();
    // ---- B: gap row pass (elas.cpp:1350-1437) -----------------------------------------------------
    for (int i = threadIdx.x; i < BH * BW; i += kFuseThreads) {
        const int y = i / BW, x = i - y * BW;
        sB[i] = gap_fill(sA + y * AW + x + G, 1, a.gap);
    }
    __syncthreads();
    // ---- C: gap column pass (elas.cpp:1440-1529) ----------------------------------------------------
    for (int i = threadIdx.x; i < CH * CW; i += kFuseThreads) {
        const int y = i / CW, x = i - y * CW;
        const float d = gap_fill(sB + (y + G) * BW + x, BW, a.gap);
        sC[i] = d;
        if (a.dump_gap || !MEAN) {
            const int v = r0 - BACK + y, u = c0 - BACK + x;
            if (v >= r0 && v < min(r0 + kFuseTH, Dh) && u >= c0 && u < min(c0 + kFuseTW, Dw))
                (MEAN ? a.dump_gap : out)[v * Dw + u] = d;
        }
    }
    if (!MEAN) return;
    __syncthreads();
    // ---- M: mean row pass (elas.cpp:1651-1700; rows 3..Dh-4, centres BACK..Dw-FWD-1) ------------------
    float* sM = sA;
    for (int i = threadIdx.x; i < MH * MW; i += kFuseThreads) {
        const int y = i / MW, x = i - y * MW;
        const int v = r0 - BACK + y, u = c0 + x;
        const float* centre = sC + y * CW + x + BACK;
        float o = *centre;
        // An invalid centre (-10) can only collect weight from other invalid taps (a valid tap is >= 10 away,
        // weight 0), so its mean is -10 < 0 and the reference keeps the pixel as it is: nothing to compute.
        if (o >= 0.f && v >= 3 && v < Dh - 3 && u >= BACK && u + FWD < Dw) {
            float r;
            // mean_window indexes line[c * stride]: pass the line origin such that c = u
            if (mean_window<TAPS>(centre - u, 1, u, &r)) o = r;
        }
        sM[i] = o;
    }
    __syncthreads();
    // ---- out: mean column pass (elas.cpp:1703-1746; columns 3..Dw-4, centres BACK..Dh-FWD-1) ----------
    for (int i = threadIdx.x; i < kFuseTH * kFuseTW; i += kFuseThreads) {
        const int y = i / kFuseTW, x = i - y * kFuseTW;
        const int v = r0 + y, u = c0 + x;
        if (v >= Dh || u >= Dw) continue;
        float o = sC[(y + BACK) * CW + x + BACK];
        // the window's centre is the row-pass value; if that is invalid the mean is too (see above)
        if (sM[(y + BACK) * MW + x] >= 0.f && u >= 3 && u < Dw - 3 && v >= BACK && v + FWD < Dh) {
            float r;
            if (mean_window<TAPS>(sM + (y + BACK) * MW + x - (ptrdiff_t)v * MW, MW, v, &r)) o = r;
        }
        out[v * Dw + u] = o;
    }
}

inline dim3 grid2d(int Dw, int Dh, int bx, int frames) { return dim3((Dw + bx - 1) / bx, Dh, frames); }

int speckle_size_of(const elas_b200_params& p)
{
    return p.subsampling ? (int)(sqrtf((float)p.speckle_size) * 2) : p.speckle_size;   // elas.cpp:1218
}

}  // namespace

OutTable out_table(float* base, size_t stride, int n_frames)
{
    OutTable t{};
    for (int i = 0; i < n_frames && i < kMaxGroupFrames; i++) t.p[i] = base + (size_t)i * stride;
    return t;
}

void launch_lr_check(const FrameGeom& g, const elas_b200_params& p, const float* D1, const float* D2,
                     float* O1, const OutTable& O2, size_t D_stride, int n_frames, cudaStream_t s)
{
    k_lr_check<<<grid2d(g.Dw, g.Dh, 256, n_frames), 256, 0, s>>>(g.Dw, g.Dh, p.subsampling, (float)p.lr_threshold, D1, D2, O1, O2, D_stride);
    count_launch();
}

bool lr_rows_fusable(const FrameGeom& g) { return (size_t)g.Dw * 8 <= 160 * 1024; }

// the narrowed right map of one frame: mode 1 = int16 [Dh][Dw]; mode 2 = u8 [Dh][Dw], then (16-byte aligned) one validity
// bit per pixel as [Dh][mask_words_per_row] ballot words
NarrowD2 narrow_d2_layout(const FrameGeom& g, int mode, void* base, size_t stride_bytes)
{
    NarrowD2 n{};
    n.base = base; n.mode = mode; n.stride_bytes = stride_bytes;
    const size_t nd = (size_t)g.Dw * g.Dh;
    n.mask_words_per_row = (g.Dw + 31) / 32;
    n.mask_offset = (nd + 15) & ~(size_t)15;
    n.bytes = mode == 1 ? nd * 2 : mode == 2 ? n.mask_offset + (size_t)g.Dh * n.mask_words_per_row * 4 : 0;
    return n;
}

// K8 for both maps, rows staged in shared memory
void launch_lr_rows(const FrameGeom& g, const elas_b200_params& p, const float* D1, const float* D2,
                    float* O1, const OutTable& O2, const NarrowD2& narrow, size_t D_stride, int n_frames, cudaStream_t s)
{
    const size_t smem = (size_t)g.Dw * 8;
    static unsigned long long optin = 0;
    if (ensure_dynamic_smem(k_lr_rows, 160 * 1024, &optin) != cudaSuccess) return;
    k_lr_rows<<<dim3(g.Dh, n_frames), 256, smem, s>>>(g.Dw, p.subsampling, (float)p.lr_threshold, D1, D2, O1, O2, narrow, D_stride);
    count_launch();
}

size_t segment_node_ints(const FrameGeom& g)
{
    const size_t tiles = (size_t)((g.Dw + kSegTW - 1) / kSegTW) * ((g.Dh + kSegTH - 1) / kSegTH);
    return 3 * tiles * kSegTile;
}

// K9: labels + open nodes of D's components (frames D_stride / nodes_stride apart); apply = true also invalidates
// the small segments in place (settings outside the fused tail), otherwise launch_post_fused applies them
void launch_segments(const FrameGeom& g, const elas_b200_params& p, float* D, int32_t* label, int32_t* nodes,
                     size_t D_stride, size_t nodes_stride, int n_frames, cudaStream_t s, bool apply)
{
    SegArgs a;
    a.Dw = g.Dw; a.Dh = g.Dh;
    a.tiles_x = (g.Dw + kSegTW - 1) / kSegTW; a.tiles_y = (g.Dh + kSegTH - 1) / kSegTH;
    a.speckle = speckle_size_of(p);
    a.thr = p.speckle_sim_threshold;
    a.D = D; a.label = label; a.nodes = nodes; a.D_stride = D_stride; a.nodes_stride = nodes_stride;
    a.node_cap = a.tiles_x * a.tiles_y * kSegTile;
    ELASB_PREPARE_KERNEL(k_seg_tiles);
    ELASB_PREPARE_KERNEL(k_seg_border);
    ELASB_PREPARE_KERNEL(k_seg_sizes);
    k_seg_tiles<<<dim3(a.tiles_x, a.tiles_y, n_frames), kSegThreads, 0, s>>>(a);
    const int pairs = (a.tiles_x - 1) * g.Dh + (a.tiles_y - 1) * g.Dw;
    k_seg_border<<<dim3(std::max(1, (pairs + 255) / 256), n_frames), 256, 0, s>>>(a);
    k_seg_sizes<<<dim3((g.Dw * g.Dh + 1023) / 1024, n_frames), 256, 0, s>>>(a);
    count_launch(3);
    if (!apply) return;
    const int n = g.Dw * g.Dh;
    k_seg_apply<<<dim3((n + 255) / 256, n_frames), 256, 0, s>>>(n, a.speckle, D, label, nodes, a.node_cap, D_stride, nodes_stride);
    count_launch();
}

bool post_fusable(const elas_b200_params& p)
{
    const int gap = p.subsampling ? p.ipol_gap_width / 2 + 1 : p.ipol_gap_width;
    return gap <= kFuseGap && gap >= 0 && !p.add_corners;
}

// speckle apply (when label != nullptr) + gap interpolation + adaptive mean (when filter_adaptive_mean)
void launch_post_fused(const FrameGeom& g, const elas_b200_params& p, const float* in, const int32_t* label,
                       const int32_t* nodes, size_t nodes_stride, const OutTable& out, float* dump_seg, float* dump_gap,
                       size_t D_stride, int n_frames, cudaStream_t s)
{
    FuseArgs a;
    a.Dw = g.Dw; a.Dh = g.Dh;
    a.gap = p.subsampling ? p.ipol_gap_width / 2 + 1 : p.ipol_gap_width;               // :1335-1341
    a.speckle = speckle_size_of(p);
    a.apply = label != nullptr;
    a.in = in; a.label = label; a.nodes = nodes; a.nodes_stride = nodes_stride;
    a.node_cap = ((g.Dw + kSegTW - 1) / kSegTW) * ((g.Dh + kSegTH - 1) / kSegTH) * kSegTile;
    a.out = out; a.dump_seg = dump_seg; a.dump_gap = dump_gap;
    a.D_stride = D_stride;
    const dim3 grid((g.Dw + kFuseTW - 1) / kFuseTW, (g.Dh + kFuseTH - 1) / kFuseTH, n_frames);
    ELASB_PREPARE_KERNEL((k_post_fused<8, false>));
    ELASB_PREPARE_KERNEL((k_post_fused<4, true>));
    ELASB_PREPARE_KERNEL((k_post_fused<8, true>));
    if (!p.filter_adaptive_mean) k_post_fused<8, false><<<grid, kFuseThreads, 0, s>>>(a);
    else if (p.subsampling)      k_post_fused<4, true><<<grid, kFuseThreads, 0, s>>>(a);
    else                         k_post_fused<8, true><<<grid, kFuseThreads, 0, s>>>(a);
    count_launch();
}

// the unfused chain works in place on D (frames D_stride apart) with one scratch plane per frame (tmp, tmp_stride apart)
void launch_gap(const FrameGeom& g, const elas_b200_params& p, float* D, float* tmp, size_t D_stride, size_t tmp_stride,
                int n_frames, cudaStream_t s)
{
    const int gap = p.subsampling ? p.ipol_gap_width / 2 + 1 : p.ipol_gap_width;     // :1335-1341
    k_gap_pass<<<grid2d(g.Dw, g.Dh, 256, n_frames), 256, 0, s>>>(g.Dw, g.Dh, gap, p.add_corners, 0, D, tmp, D_stride, tmp_stride);
    k_gap_pass<<<grid2d(g.Dw, g.Dh, 256, n_frames), 256, 0, s>>>(g.Dw, g.Dh, gap, p.add_corners, 1, tmp, D, tmp_stride, D_stride);
    count_launch(2);
}

void launch_adaptive_mean(const FrameGeom& g, const elas_b200_params& p, float* D, float* tmp, size_t D_stride,
                          size_t tmp_stride, int n_frames, cudaStream_t s)
{
    // The reference filters a copy of D in which invalid pixels are set to -10 (elas.cpp:1553-1559).
    // Here every invalid pixel already IS -10: the L/R check writes -10 for everything it rejects
    // (elas.cpp:1172-1196) and speckle removal / gap interpolation only write -10 or valid values,
    // so the copy is D itself.  tmp = the reference's D_tmp (one plane).
    const dim3 grid = grid2d(g.Dw, g.Dh, 256, n_frames);
    if (p.subsampling) {
        k_mean_pass<4><<<grid, 256, 0, s>>>(g.Dw, g.Dh, 0, D, D, tmp, D_stride, D_stride, tmp_stride);
        k_mean_pass<4><<<grid, 256, 0, s>>>(g.Dw, g.Dh, 1, tmp, D, D, tmp_stride, D_stride, D_stride);
    } else {
        k_mean_pass<8><<<grid, 256, 0, s>>>(g.Dw, g.Dh, 0, D, D, tmp, D_stride, D_stride, tmp_stride);
        k_mean_pass<8><<<grid, 256, 0, s>>>(g.Dw, g.Dh, 1, tmp, D, D, tmp_stride, D_stride, D_stride);
    }
    count_launch(2);
}

void launch_median(const FrameGeom& g, float* D, float* tmp, size_t D_stride, size_t tmp_stride, int n_frames, cudaStream_t s)
{
    const dim3 grid = grid2d(g.Dw, g.Dh, 256, n_frames);
    k_median_pass<<<grid, 256, 0, s>>>(g.Dw, g.Dh, 0, D, D, tmp, D_stride, D_stride, tmp_stride);
    k_median_pass<<<grid, 256, 0, s>>>(g.Dw, g.Dh, 1, D, tmp, D, D_stride, tmp_stride, D_stride);
    count_launch(2);
}

}  // namespace elasb
