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

// ---------------------------------------------------------------------------------------------
// K9  speckle removal, elas.cpp:1208-1326.  The reference flood-fills 4-connected segments in
// which neighbouring valid pixels differ by <= speckle_sim_threshold and invalidates segments with
// fewer than speckle_size pixels.  Segments are the connected components of a symmetric relation,
// so the result does not depend on traversal order.  Run-based labelling:
//   rows:   every maximal horizontal run of connected pixels becomes one union-find node (its first
//           pixel); the other pixels of the run point at it and are never touched again
//   merge:  vertically connected pixel pairs union their runs; a pair is skipped when its left
//           neighbour pair already joins the same two runs
//   count:  each run adds its length to its root once (and stops adding once the root is known to be
//           large enough -- only "size < speckle_size" is ever asked)
//   apply:  pixel -> run -> root -> size
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool seg_conn(float a, float b, float thr)
{
    return a >= 0.f && b >= 0.f && fabsf(__fsub_rn(a, b)) <= thr;                    // :1281, :1285
}

// parent[] is read at L2 (ld.global.cg): other SMs hook roots concurrently.
__device__ __forceinline__ int uf_find(int32_t* parent, int x)
{
    int p = parent[x];
    while (p != x) {
        const int gp = parent[p];
        if (gp != p) parent[x] = gp;           // path halving
        x = p; p = gp;
    }
    return x;
}

// Hooking order: a root goes under the root with the smaller PRIORITY, a fixed pseudo-random
// permutation of the pixel index.  Index order would let the thousands of unions of a frame, which run
// concurrently, build chains as long as the image is high (every run of a vertical structure hooking
// under the run above it at the same moment), and every later find would walk them at L2 latency;
// with random priorities simultaneous hooks only chain along decreasing-priority sequences, whose
// expected length is logarithmic.
__device__ __forceinline__ uint32_t uf_priority(int x) { return (uint32_t)x * 0x9E3779B1u; }

__device__ __forceinline__ void uf_union(int32_t* parent, int a, int b)
{
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (uf_priority(a) < uf_priority(b)) { const int t = a; a = b; b = t; }   // hook a (larger priority) under b
        const int old = atomicCAS(parent + a, a, b);
        if (old == a) return;
        a = old;
    }
}

// one CTA per row: parent[pixel] = index of the first pixel of its run (-1 for invalid pixels).
// `row` may live in shared memory (k_lr_rows) or in global memory (k_seg_rows).
__device__ __forceinline__ void label_row_runs(const float* row, int Dw, int base, float thr,
                                               int32_t* __restrict__ parent, int32_t* __restrict__ size,
                                               int* warp_last, int* carry_s)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) *carry_s = -1;
    __syncthreads();
    for (int u0 = 0; u0 < Dw; u0 += 256) {
        const int u = u0 + threadIdx.x;
        const float d = u < Dw ? row[u] : -1.f;
        const float dl = (u > 0 && u < Dw) ? row[u - 1] : -1.f;
        const bool valid = d >= 0.f;
        const bool start = valid && !seg_conn(dl, d, thr);
        const unsigned starts = __ballot_sync(0xffffffffu, start);
        const unsigned upto = starts & (0xffffffffu >> (31 - lane));
        int last = upto ? u0 + (warp << 5) + 31 - __clz(upto) : -1;      // most recent start in this warp
        if (lane == 31) warp_last[warp] = last;
        __syncthreads();
        const int carry = *carry_s;
        if (last < 0) {
            for (int w = warp - 1; w >= 0 && last < 0; w--) last = warp_last[w];
            if (last < 0) last = carry;
        }
        if (u < Dw) {
            parent[base + u] = valid ? base + last : -1;
            if (start) size[base + u] = 0;          // sizes are only ever read and accumulated at run starts (roots)
        }
        __syncthreads();
        if (threadIdx.x == 255) *carry_s = last >= 0 ? last : carry;     // runs never span an invalid pixel
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
k_seg_rows(int Dw, float thr, const float* __restrict__ D, int32_t* __restrict__ parent, int32_t* __restrict__ size, size_t D_stride)
{
    __shared__ int warp_last[8];
    __shared__ int carry_s;
    const int v = blockIdx.x;
    D += blockIdx.y * D_stride; parent += blockIdx.y * D_stride; size += blockIdx.y * D_stride;
    label_row_runs(D + (size_t)v * Dw, Dw, v * Dw, thr, parent, size, warp_last, &carry_s);
}

// K8 + the row step of K9 in one pass: a CTA owns one row of both maps.  The L/R check of a pixel
// only reads the other map within the same row (elas.cpp:1164-1197), so both raw rows are staged in
// shared memory, the checked rows are written out, and the checked D1 row -- still in shared memory --
// is labelled into runs.
__global__ void __launch_bounds__(256)
k_lr_rows(int Dw, int subsampling, float lr_threshold, float thr,
          const float* __restrict__ D1, const float* __restrict__ D2,
          float* __restrict__ O1, OutTable O2_tab, int32_t* __restrict__ parent, int32_t* __restrict__ size,
          int16_t* __restrict__ O2_i16,      // optional: O2 narrowed (exact: raw integer disparities or -10)
          size_t D_stride)
{
    extern __shared__ float s_rows[];          // [3][Dw]: raw D1 row, raw D2 row, checked D1 row
    __shared__ int warp_last[8];
    __shared__ int carry_s;
    float* r1 = s_rows; float* r2 = s_rows + Dw; float* c1 = s_rows + 2 * Dw;
    const int v = blockIdx.x;
    // blockIdx.y = frame of the group
    D1 += blockIdx.y * D_stride; D2 += blockIdx.y * D_stride; O1 += blockIdx.y * D_stride;
    parent += blockIdx.y * D_stride; size += blockIdx.y * D_stride;
    if (O2_i16) O2_i16 += blockIdx.y * D_stride;
    float* __restrict__ O2 = O2_tab.p[blockIdx.y];
    const size_t row = (size_t)v * Dw;
    for (int u = threadIdx.x; u < Dw; u += 256) { r1[u] = D1[row + u]; r2[u] = D2[row + u]; }
    __syncthreads();
    for (int u = threadIdx.x; u < Dw; u += 256) {
        const float d1 = r1[u], d2 = r2[u];
        const float w1 = subsampling ? __fsub_rn((float)u, __fmul_rn(d1, 0.5f)) : __fsub_rn((float)u, d1);  // :1152-1161
        const float w2 = subsampling ? __fadd_rn((float)u, __fmul_rn(d2, 0.5f)) : __fadd_rn((float)u, d2);
        float o1 = (float)kInvalid, o2 = (float)kInvalid;
        if (d1 >= 0.f && w1 >= 0.f && w1 < (float)Dw)                                                       // :1164-1179
            if (!(fabsf(__fsub_rn(r2[(int)w1], d1)) > lr_threshold)) o1 = d1;
        if (d2 >= 0.f && w2 >= 0.f && w2 < (float)Dw)                                                       // :1182-1197
            if (!(fabsf(__fsub_rn(r1[(int)w2], d2)) > lr_threshold)) o2 = d2;
        O1[row + u] = o1;
        if (O2_i16) O2_i16[row + u] = (int16_t)o2; else O2[row + u] = o2;
        c1[u] = o1;
    }
    __syncthreads();
    label_row_runs(c1, Dw, v * Dw, thr, parent, size, warp_last, &carry_s);
}

// k_seg_merge and k_seg_count are chains of dependent L2 accesses (find, CAS): their warps are stalled
// almost all the time.  They run as a modest grid-stride grid (kSegCtasPerSm CTAs per SM) instead of one
// thread per pixel, so that they occupy a quarter of an SM's thread slots while the pipeline's other
// kernels (other slots' frames) use the issue slots they leave idle.
constexpr int kSegCtasPerSm = 2;

__global__ void __launch_bounds__(256)
k_seg_merge(int Dw, int Dh, float thr, const float* __restrict__ D, int32_t* parent, size_t D_stride)
{
    const int n = Dw * (Dh - 1);
    D += blockIdx.y * D_stride; parent += blockIdx.y * D_stride;
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x) {
        const int u = a % Dw, b = a + Dw;
        const float da = D[a], db = D[b];
        if (!seg_conn(da, db, thr)) continue;
        bool a_start = true, b_start = true;
        if (u > 0) {
            const float la = D[a - 1], lb = D[b - 1];
            a_start = !seg_conn(la, da, thr);
            b_start = !seg_conn(lb, db, thr);
            if (!a_start && !b_start && seg_conn(la, lb, thr)) continue;   // the pair to the left joins the same runs
        }
        uf_union(parent, a_start ? a : __ldcg(parent + a), b_start ? b : __ldcg(parent + b));
    }
}

__global__ void __launch_bounds__(256)
k_seg_count(int Dw, int Dh, float thr, int speckle, const float* __restrict__ D,
            int32_t* parent, int32_t* size, size_t D_stride)
{
    const int n = Dw * Dh;
    D += blockIdx.y * D_stride; parent += blockIdx.y * D_stride; size += blockIdx.y * D_stride;
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x) {
        const int u = a % Dw;
        const float d = D[a];
        if (!(d >= 0.f)) continue;
        if (u + 1 < Dw && seg_conn(d, D[a + 1], thr)) continue;           // not the last pixel of its run
        const bool is_start = !(u > 0 && seg_conn(D[a - 1], d, thr));
        const int start = is_start ? a : __ldcg(parent + a);
        const int root = uf_find(parent, start);
        __stcg(parent + start, root);                                           // every run ends up one hop from its root
        if (__ldcg(size + root) < speckle) atomicAdd(size + root, a - start + 1);
    }
}

__global__ void k_seg_apply(int n, int speckle, float* __restrict__ D, const int32_t* __restrict__ parent,
                            const int32_t* __restrict__ size, size_t D_stride)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    D += blockIdx.y * D_stride; parent += blockIdx.y * D_stride; size += blockIdx.y * D_stride;
    const float d = D[i];
    if (d >= 0.f) {
        int root = parent[i];                      // pixel -> run start -> (usually one hop) -> root
        for (int up = parent[root]; up != root; up = parent[root]) root = up;
        if (size[root] < speckle) D[i] = (float)kInvalid;                              // :1309-1317
    } else if (1 < speckle) D[i] = (float)kInvalid;   // an invalid pixel is a segment of one (:1248-1250)
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
    const int32_t* parent;
    const int32_t* size;
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
    if (a.apply) { a.parent += blockIdx.z * a.D_stride; a.size += blockIdx.z * a.D_stride; }
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
        int idx[4], root[4];
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
            for (int k = 0; k < 4; k++) root[k] = d[k] >= 0.f ? a.parent[idx[k]] : -1;      // run start
#pragma unroll
            for (int k = 0; k < 4; k++) if (root[k] >= 0) root[k] = a.parent[root[k]];      // its root (k_seg_count left it one hop away)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (root[k] < 0) continue;
                for (int up = a.parent[root[k]]; up != root[k]; up = a.parent[root[k]]) root[k] = up;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (root[k] >= 0) { if (a.size[root[k]] < a.speckle) d[k] = (float)kInvalid; }
                else if (1 < a.speckle) d[k] = (float)kInvalid;     // an invalid pixel is a segment of one (:1248-1250)
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

bool lr_rows_fusable(const FrameGeom& g) { return (size_t)g.Dw * 12 <= 160 * 1024; }

// K8 for both maps + the run labelling of D1 (launch_segments(..., rows_done = true) continues from there)
void launch_lr_rows(const FrameGeom& g, const elas_b200_params& p, const float* D1, const float* D2,
                    float* O1, const OutTable& O2, int32_t* parent, int32_t* size, int16_t* O2_i16, size_t D_stride,
                    int n_frames, cudaStream_t s)
{
    const size_t smem = (size_t)g.Dw * 12;
    static unsigned long long optin = 0;
    if (ensure_dynamic_smem(k_lr_rows, 160 * 1024, &optin) != cudaSuccess) return;
    k_lr_rows<<<dim3(g.Dh, n_frames), 256, smem, s>>>(g.Dw, p.subsampling, (float)p.lr_threshold, p.speckle_sim_threshold,
                                                      D1, D2, O1, O2, parent, size, O2_i16, D_stride);
    count_launch();
}

void launch_segments(const FrameGeom& g, const elas_b200_params& p, float* D, int32_t* parent,
                     int32_t* size, size_t D_stride, int n_frames, cudaStream_t s, bool apply, bool rows_done)
{
    const int n = g.Dw * g.Dh;
    const int speckle = speckle_size_of(p);
    const float thr = p.speckle_sim_threshold;
    if (!rows_done) { k_seg_rows<<<dim3(g.Dh, n_frames), 256, 0, s>>>(g.Dw, thr, D, parent, size, D_stride); count_launch(); }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static int sms_of[64] = {0};
    if (!sms_of[dev & 63]) { cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); sms_of[dev & 63] = sms; }
    static const int per_sm = [] { const char* e = std::getenv("ELAS_B200_SEG_CTAS_PER_SM"); return e ? std::atoi(e) : kSegCtasPerSm; }();
    const int seg_grid = std::min(std::max(1, sms_of[dev & 63] * per_sm / std::max(1, std::min(n_frames, 4))), (n + 255) / 256);
    ELASB_PREPARE_KERNEL(k_seg_merge);
    ELASB_PREPARE_KERNEL(k_seg_count);
    k_seg_merge<<<dim3(seg_grid, n_frames), 256, 0, s>>>(g.Dw, g.Dh, thr, D, parent, D_stride);
    k_seg_count<<<dim3(seg_grid, n_frames), 256, 0, s>>>(g.Dw, g.Dh, thr, speckle, D, parent, size, D_stride);
    if (!apply) { count_launch(2); return; }
    k_seg_apply<<<dim3((n + 255) / 256, n_frames), 256, 0, s>>>(n, speckle, D, parent, size, D_stride);
    count_launch(3);
}

bool post_fusable(const elas_b200_params& p)
{
    const int gap = p.subsampling ? p.ipol_gap_width / 2 + 1 : p.ipol_gap_width;
    return gap <= kFuseGap && gap >= 0 && !p.add_corners;
}

// speckle apply (when parent != nullptr) + gap interpolation + adaptive mean (when filter_adaptive_mean)
void launch_post_fused(const FrameGeom& g, const elas_b200_params& p, const float* in, const int32_t* parent,
                       const int32_t* size, const OutTable& out, float* dump_seg, float* dump_gap, size_t D_stride,
                       int n_frames, cudaStream_t s)
{
    FuseArgs a;
    a.Dw = g.Dw; a.Dh = g.Dh;
    a.gap = p.subsampling ? p.ipol_gap_width / 2 + 1 : p.ipol_gap_width;               // :1335-1341
    a.speckle = speckle_size_of(p);
    a.apply = parent != nullptr;
    a.in = in; a.parent = parent; a.size = size; a.out = out; a.dump_seg = dump_seg; a.dump_gap = dump_gap;
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
