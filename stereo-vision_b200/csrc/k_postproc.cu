// K8-K12: post-processing of the disparity maps (elas.cpp:1122-1838), one thread per pixel.
//
// All five reference stages are sequential scans written in place; each is restated here in a form
// whose per-pixel result depends only on the stage's INPUT, so that pixels can be computed
// independently (the notes at each kernel say why that is equivalent).  Float expressions use the
// explicit _rn intrinsics: the reference is x86-64 SSE code without FMA contraction.
#include "common.cuh"

namespace elasb {
namespace {

// ---------------------------------------------------------------------------------------------
// K8  left/right consistency check, elas.cpp:1122-1204.  The reference reads copies of D1/D2 and
// writes the originals; here input and output are separate buffers.
// ---------------------------------------------------------------------------------------------
__global__ void k_lr_check(int Dw, int Dh, int subsampling, float lr_threshold,
                           const float* __restrict__ D1, const float* __restrict__ D2,
                           float* __restrict__ O1, float* __restrict__ O2)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= Dw || v >= Dh) return;
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
// so the result does not depend on traversal order: union-find over the pixel lattice.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(int32_t* parent, int x)
{
    int p = parent[x];
    while (p != x) { x = p; p = parent[x]; }
    return x;
}

__device__ __forceinline__ void uf_union(int32_t* parent, int a, int b)
{
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }        // hook the larger root under the smaller
        const int old = atomicMin(parent + a, b);
        if (old == a) return;
        a = old;
    }
}

__global__ void k_seg_init(int n, int32_t* __restrict__ parent, int32_t* __restrict__ size)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { parent[i] = i; size[i] = 0; }
}

__global__ void k_seg_link(int Dw, int Dh, float thr, const float* __restrict__ D, int32_t* parent)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= Dw || v >= Dh) return;
    const int a = v * Dw + u;
    const float d = D[a];
    if (!(d >= 0.f)) return;                                                        // :1281
    if (u + 1 < Dw) {
        const float e = D[a + 1];
        if (e >= 0.f && fabsf(__fsub_rn(d, e)) <= thr) uf_union(parent, a, a + 1);  // :1285
    }
    if (v + 1 < Dh) {
        const float e = D[a + Dw];
        if (e >= 0.f && fabsf(__fsub_rn(d, e)) <= thr) uf_union(parent, a, a + Dw);
    }
}

__global__ void k_seg_count(int n, const float* __restrict__ D, int32_t* parent, int32_t* size)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n && D[i] >= 0.f;
    int root = -1;
    if (valid) { root = uf_find(parent, i); parent[i] = root; }
    // one atomic per distinct root per warp: most of the image is a handful of large segments
    const unsigned peers = __match_any_sync(0xffffffffu, root);
    if (valid && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(size + root, __popc(peers));
}

__global__ void k_seg_apply(int n, int speckle, float* __restrict__ D, const int32_t* __restrict__ parent,
                            const int32_t* __restrict__ size)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d = D[i];
    if (d >= 0.f) { if (size[parent[i]] < speckle) D[i] = (float)kInvalid; }          // :1309-1317
    else if (1 < speckle) D[i] = (float)kInvalid;   // an invalid pixel is a segment of one (:1248-1250)
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
                           const float* __restrict__ in, float* __restrict__ out)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= Dw || v >= Dh) return;
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

template <int TAPS>
__device__ __forceinline__ bool mean_window(const float* __restrict__ line, ptrdiff_t stride, int c, float* result)
{
    constexpr int BACK = TAPS == 8 ? 4 : 2;          // window = [c-BACK, c+TAPS-BACK-1]
    const float xc = line[c * stride];
    float w[TAPS], f[TAPS];
#pragma unroll
    for (int k = 0; k < TAPS; k++) {
        const int pos = c - BACK + k;
        const float x = line[pos * stride];
        const float wk = fmaxf(0.f, __fsub_rn(4.0f, masked_abs(__fsub_rn(x, xc))));
        const int slot = pos & (TAPS - 1);           // val[u % taps], :1667 / :1590
        w[slot] = wk;
        f[slot] = __fmul_rn(x, wk);
    }
    float ws, fs;
    if (TAPS == 8) {
        ws = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(w[0], w[4]), __fadd_rn(w[1], w[5])), __fadd_rn(w[2], w[6])), __fadd_rn(w[3], w[7]));
        fs = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(f[0], f[4]), __fadd_rn(f[1], f[5])), __fadd_rn(f[2], f[6])), __fadd_rn(f[3], f[7]));
    } else {
        ws = __fadd_rn(__fadd_rn(__fadd_rn(w[0], w[1]), w[2]), w[3]);
        fs = __fadd_rn(__fadd_rn(__fadd_rn(f[0], f[1]), f[2]), f[3]);
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
                            const float* keep, float* out)      // keep may alias out (vertical pass)
{
    constexpr int BACK = TAPS == 8 ? 4 : 2, FWD = TAPS - BACK - 1;
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= Dw || v >= Dh) return;
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

__global__ void k_invalid_to_m10(int n, const float* __restrict__ in, float* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float d = in[i]; out[i] = d >= 0.f ? d : (float)kInvalid; }
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
                              const float* in, float* out)        // D aliases in (horizontal) or out (vertical)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= Dw || v >= Dh) return;
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

inline dim3 grid2d(int Dw, int Dh, int bx) { return dim3((Dw + bx - 1) / bx, Dh, 1); }

}  // namespace

void launch_lr_check(const FrameGeom& g, const elas_b200_params& p, const float* D1, const float* D2,
                     float* O1, float* O2, cudaStream_t s)
{
    k_lr_check<<<grid2d(g.Dw, g.Dh, 256), 256, 0, s>>>(g.Dw, g.Dh, p.subsampling, (float)p.lr_threshold, D1, D2, O1, O2);
    count_launch();
}

void launch_segments(const FrameGeom& g, const elas_b200_params& p, float* D, int32_t* parent,
                     int32_t* size, cudaStream_t s)
{
    const int n = g.Dw * g.Dh;
    int speckle = p.speckle_size;
    if (p.subsampling) speckle = (int)(sqrtf((float)p.speckle_size) * 2);            // :1218
    k_seg_init<<<(n + 255) / 256, 256, 0, s>>>(n, parent, size);
    k_seg_link<<<grid2d(g.Dw, g.Dh, 256), 256, 0, s>>>(g.Dw, g.Dh, p.speckle_sim_threshold, D, parent);
    k_seg_count<<<(n + 255) / 256, 256, 0, s>>>(n, D, parent, size);
    k_seg_apply<<<(n + 255) / 256, 256, 0, s>>>(n, speckle, D, parent, size);
    count_launch(4);
}

void launch_gap(const FrameGeom& g, const elas_b200_params& p, float* D, float* tmp, cudaStream_t s)
{
    const int gap = p.subsampling ? p.ipol_gap_width / 2 + 1 : p.ipol_gap_width;     // :1335-1341
    k_gap_pass<<<grid2d(g.Dw, g.Dh, 256), 256, 0, s>>>(g.Dw, g.Dh, gap, p.add_corners, 0, D, tmp);
    k_gap_pass<<<grid2d(g.Dw, g.Dh, 256), 256, 0, s>>>(g.Dw, g.Dh, gap, p.add_corners, 1, tmp, D);
    count_launch(2);
}

void launch_adaptive_mean(const FrameGeom& g, const elas_b200_params& p, float* D, float* tmp,
                          cudaStream_t s)
{
    // tmp holds two planes: [0] = D_copy (invalid -> -10), [1] = D_tmp
    const int n = g.Dw * g.Dh;
    float* copy = tmp;
    float* dtmp = tmp + n;
    k_invalid_to_m10<<<(n + 255) / 256, 256, 0, s>>>(n, D, copy);
    if (p.subsampling) {
        k_mean_pass<4><<<grid2d(g.Dw, g.Dh, 256), 256, 0, s>>>(g.Dw, g.Dh, 0, copy, copy, dtmp);
        k_mean_pass<4><<<grid2d(g.Dw, g.Dh, 256), 256, 0, s>>>(g.Dw, g.Dh, 1, dtmp, D, D);
    } else {
        k_mean_pass<8><<<grid2d(g.Dw, g.Dh, 256), 256, 0, s>>>(g.Dw, g.Dh, 0, copy, copy, dtmp);
        k_mean_pass<8><<<grid2d(g.Dw, g.Dh, 256), 256, 0, s>>>(g.Dw, g.Dh, 1, dtmp, D, D);
    }
    count_launch(3);
}

void launch_median(const FrameGeom& g, float* D, float* tmp, cudaStream_t s)
{
    k_median_pass<<<grid2d(g.Dw, g.Dh, 256), 256, 0, s>>>(g.Dw, g.Dh, 0, D, D, tmp);
    k_median_pass<<<grid2d(g.Dw, g.Dh, 256), 256, 0, s>>>(g.Dw, g.Dh, 1, D, tmp, D);
    count_launch(2);
}

}  // namespace elasb
