// the two per-pixel loops that consume D1 inside stereomapper's StereoThread,
// run where D1 already is -- in HBM -- instead of after a device->host copy:
//   k_colormap   HSV colour map of min(D1/200, 1)                       stereothread.cpp:116-147
//   k_reproject  back-projection z = f*b/d, x = (u-cu)*b/d, y = (v-cv)*b/d through the 3x4 pose,
//                intensity image I1/255 with the border gain ramp       stereothread.cpp:180-255
//   k_fuse_*     fusion of the current map with the previous one             stereothread.cpp:290-437
// One thread per pixel.  Expression types follow the C++ source (double where it promotes to double,
// separate roundings: the reference is x86-64 SSE code built at -O0, stereomapper.pro:145-150).
#include "common.cuh"

namespace elasb {
namespace {

__global__ void __launch_bounds__(256)
k_colormap(int n, const float* __restrict__ D1, float* __restrict__ out)
{
    __shared__ float s_rgb[3 * 256];          // a thread's r,g,b are 12 bytes apart: staged so that the CTA writes 3 KB in order
    const int i0 = blockIdx.x * 256, i = i0 + threadIdx.x;
    float r = 0.f, g = 0.f, b = 0.f;
    if (i < n) {
        float val = __fdiv_rn(D1[i], 200.f);                                      // :117, :130
        if (1.0f < val) val = 1.0f;                                               // std::min
        if (!(val <= 0.f)) {
            const float h2 = __double2float_rn(__dmul_rn(6.0, __dsub_rn(1.0, (double)val)));                 // :137
            const float x = __double2float_rn(__dsub_rn(1.0, fabs(__dsub_rn((double)fmodf(h2, 2.0f), 1.0)))); // :138
            if      (0.f <= h2 && h2 < 1.f)  { r = 1.f; g = x; }                  // :139-144
            else if (1.f <= h2 && h2 < 2.f)  { r = x; g = 1.f; }
            else if (2.f <= h2 && h2 < 3.f)  { g = 1.f; b = x; }
            else if (3.f <= h2 && h2 < 4.f)  { g = x; b = 1.f; }
            else if (4.f <= h2 && h2 < 5.f)  { r = x; b = 1.f; }
            else if (5.f <= h2 && h2 <= 6.f) { r = 1.f; b = x; }
        }
    }
    s_rgb[3 * threadIdx.x] = r; s_rgb[3 * threadIdx.x + 1] = g; s_rgb[3 * threadIdx.x + 2] = b;
    __syncthreads();
    const int count = 3 * min(256, n - i0);
    for (int k = threadIdx.x; k < count; k += 256) out[3 * (size_t)i0 + k] = s_rgb[k];
}

}  // namespace

// Intensity and gain values have few distinct inputs (256 grey levels, <= 200 ramp positions): the
// host evaluates the reference's double expressions once per call and the kernel looks them up.
struct ViewArgs {
    int W, H, pitch, margin;
    float f, cu, cv, base, max_dist;
    float h[12];                    // pose rows 0..2, narrowed to float like the reference's hcf.. (:201-204)
    float intensity[256];           // (float)(((float)i)/255.0), :198
    float gain[200];                // g of ramp position i, :240
};

namespace {

__global__ void k_reproject(const __grid_constant__ ViewArgs a, const uint8_t* __restrict__ img,
                            const float* __restrict__ D1, float* __restrict__ I, float* __restrict__ D,
                            float* __restrict__ X, float* __restrict__ Y, float* __restrict__ Z)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= a.W) return;
    const size_t p = (size_t)v * a.W + u;
    float d = D1[p];
    float X_ = 0.f, Y_ = 0.f, Z_ = 0.f;                                        // the reference leaves these unwritten
    if (d > 0.f) {                                                             // :212
        const float z = __fdiv_rn(__fmul_rn(a.f, a.base), d);                  // :214
        if ((double)z > 0.1 && z < a.max_dist) {                               // :215
            const float x = __fdiv_rn(__fmul_rn(__fsub_rn((float)u, a.cu), a.base), d);   // :217-218
            const float y = __fdiv_rn(__fmul_rn(__fsub_rn((float)v, a.cv), a.base), d);
            X_ = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.h[0], x), __fmul_rn(a.h[1], y)), __fmul_rn(a.h[2], z)), a.h[3]);
            Y_ = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.h[4], x), __fmul_rn(a.h[5], y)), __fmul_rn(a.h[6], z)), a.h[7]);
            Z_ = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.h[8], x), __fmul_rn(a.h[9], y)), __fmul_rn(a.h[10], z)), a.h[11]);
        } else {
            d = -1.f;                                                          // :225
        }
    }
    D[p] = d; X[p] = X_; Y[p] = Y_; Z[p] = Z_;
    // intensity with the gain ramp on the image border (:231-252): rows i and H-1-i for the columns
    // between the margins, columns i and W-1-i for the rows between the margins; corners untouched
    float in = a.intensity[img[(size_t)v * a.pitch + u]];
    const bool mid_u = u >= a.margin && u < a.W - a.margin, mid_v = v >= a.margin && v < a.H - a.margin;
    int ramp = -1;
    if (mid_u && !mid_v) ramp = v < a.margin ? v : a.H - 1 - v;
    else if (mid_v && !mid_u) ramp = u < a.margin ? u : a.W - 1 - u;
    if (ramp >= 0 && ramp < a.margin) {
        float t = __fmul_rn(a.gain[ramp], in);
        t = t < 0.f ? 0.f : t;                                                 // std::max, std::min
        in = 1.f < t ? 1.f : t;
    }
    I[p] = in;
}

}  // namespace

void launch_colormap(int n, const float* D1, float* out, cudaStream_t s)
{
    k_colormap<<<(n + 255) / 256, 256, 0, s>>>(n, D1, out);
    count_launch();
}

void launch_reproject(int W, int H, const uint8_t* img, int pitch, const float* D1, const elas_b200_view& view,
                      float* I, float* D, float* X, float* Y, float* Z, cudaStream_t s)
{
    ViewArgs a;
    a.W = W; a.H = H; a.pitch = pitch;
    a.margin = W / 2 < 200 ? W / 2 : 200;                                      // :232
    if (H / 2 < a.margin) a.margin = H / 2;
    a.f = view.f; a.cu = view.cu; a.cv = view.cv; a.base = view.base; a.max_dist = view.max_dist;
    for (int k = 0; k < 12; k++) a.h[k] = (float)view.H[k];
    for (int i = 0; i < 256; i++) a.intensity[i] = (float)(((float)i) / 255.0);
    float gain_inv = 1;                                                         // :233-237
    if (view.gain) gain_inv = (float)(1.0 / view.gain);
    for (int i = 0; i < 200; i++)
        a.gain[i] = i < a.margin ? (float)(((float)(a.margin - i) * gain_inv + (float)i * 1.0) / (float)a.margin) : 1.f;
    k_reproject<<<dim3((W + 255) / 256, H), 256, 0, s>>>(a, img, D1, I, D, X, Y, Z);
    count_launch();
}


// ---------------------------------------------------------------------------------------------
// StereoThread::addDisparityMapToReconstruction, stereothread.cpp:290-437.
// The reference walks the PREVIOUS map in u-outer / v-inner order; every valid point is projected into the
// current image and either averaged with the current point there (if closer than 0.2 in L1), or creates a point
// where the current map has none, or is kept as a point of its own.  Several previous points can land on one
// current pixel and each sees what the ones before it left there, so the order matters:
//   project:  one thread per previous pixel: target pixel (or "keep"), pushed on the target's list (atomicExch)
//   apply:    one thread per current pixel with a list: its sources in ascending scan order, one after the other
//   compact:  the kept previous points / the valid current points, in scan order (three-pass prefix sum)
// The previous map of a call is the fused current map of the call before (the reference's own hand-over frees
// what it has just copied, :433-434).
// ---------------------------------------------------------------------------------------------
namespace {

struct FuseMapArgs {
    int W, H;
    float hfc2[4];              // row 2 of inv(pose), narrowed to float (:303-304)
    float pfc[3][4];            // K * inv(pose)[0..2][:] (:307-310)
    float max_dist;
    float *pI, *pD, *pX, *pY, *pZ, *cI, *cD, *cX, *cY, *cZ;
    int32_t* target;            // per previous pixel in scan order: >= 0 target pixel, kKeep, kNone, kAdded
    int32_t* head;              // per current pixel: first source of its list (-1: none)
    int32_t* next;              // per previous pixel (scan order): next source on the same list
    int32_t* long_targets;      // current pixels whose list is longer than kFuseShort
    int32_t* scratch;           // their lists, sorted (2 * W * H)
    int32_t* counters;          // [0] long targets, [1] scratch ints handed out
};
constexpr int kFuseKeep = -1, kFuseNone = -2, kFuseAdded = -3;

// (int32_t) of a float on x86-64 (cvttss2si): out of range and NaN give INT32_MIN
__device__ __forceinline__ int x86_float_to_int(float q)
{
    return (q >= -2147483648.0f && q < 2147483648.0f) ? __float2int_rz(q) : (int)0x80000000;
}
__device__ __forceinline__ float affine3(const float* c, float x, float y, float z)
{
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c[0], x), __fmul_rn(c[1], y)), __fmul_rn(c[2], z)), c[3]);
}

__global__ void __launch_bounds__(256)
k_fuse_project(const __grid_constant__ FuseMapArgs a)
{
    const int s = blockIdx.x * 256 + threadIdx.x;              // scan order: u outer, v inner (:315-318)
    if (s >= a.W * a.H) return;
    const int u = s / a.H, v = s - u * a.H, addr = v * a.W + u;
    const float d = a.pD[addr];
    int t = kFuseNone;
    if (d > 0.f) {                                              // :325
        t = kFuseKeep;
        const float x = a.pX[addr], y = a.pY[addr], z = a.pZ[addr];
        const float z2 = affine3(a.hfc2, x, y, z);              // :333
        if ((double)z2 > 0.1 && z2 < a.max_dist) {              // :336
            const float w2 = affine3(a.pfc[2], x, y, z);        // :339-341
            const int u2 = x86_float_to_int(__fdiv_rn(affine3(a.pfc[0], x, y, z), w2));
            const int v2 = x86_float_to_int(__fdiv_rn(affine3(a.pfc[1], x, y, z), w2));
            if (u2 >= 0 && u2 < a.W && v2 >= 0 && v2 < a.H) {   // :344
                t = v2 * a.W + u2;
                a.next[s] = atomicExch(a.head + t, s);
            }
        }
    }
    a.target[s] = t;
}

// one previous point (scan position s) meets the running state of its target pixel (:352-383)
struct FusePixel { float X, Y, Z, I, D; };
__device__ __forceinline__ void fuse_apply_one(const FuseMapArgs& a, int s, FusePixel& c)
{
    const int u = s / a.H, v = s - u * a.H, addr = v * a.W + u;
    const float x = a.pX[addr], y = a.pY[addr], z = a.pZ[addr];
    bool added = false;
    if (c.D > 0.f) {                                                                         // :354
        const float dist = __fadd_rn(__fadd_rn(fabsf(__fsub_rn(x, c.X)), fabsf(__fsub_rn(y, c.Y))), fabsf(__fsub_rn(z, c.Z)));
        if ((double)dist < 0.2) {                                                            // :359
            c.X = __fmul_rn(__fadd_rn(c.X, x), 0.5f);                                        // (float)((X + x) / 2.0), exact halving
            c.Y = __fmul_rn(__fadd_rn(c.Y, y), 0.5f);
            c.Z = __fmul_rn(__fadd_rn(c.Z, z), 0.5f);
            c.I = __fmul_rn(__fadd_rn(c.I, a.pI[addr]), 0.5f);
            added = true;
        }
    } else {                                                                                 // :371-378
        c.X = x; c.Y = y; c.Z = z; c.I = a.pI[addr]; c.D = 1.f;
        added = true;
    }
    if (added) a.pD[addr] = -1.f;                                                            // :383
    a.target[s] = added ? kFuseAdded : kFuseKeep;
}
__device__ __forceinline__ FusePixel fuse_load(const FuseMapArgs& a, int addr2)
{
    return FusePixel{a.cX[addr2], a.cY[addr2], a.cZ[addr2], a.cI[addr2], a.cD[addr2]};
}
__device__ __forceinline__ void fuse_store(const FuseMapArgs& a, int addr2, const FusePixel& c)
{
    a.cX[addr2] = c.X; a.cY[addr2] = c.Y; a.cZ[addr2] = c.Z; a.cI[addr2] = c.I; a.cD[addr2] = c.D;
}

// Lists are in arrival order; the reference takes a pixel's sources in ascending scan order.  Up to kFuseShort
// sources (practically every pixel: a handful unless the camera backs away fast) are sorted in registers by the
// pixel's own thread; longer lists are handed to k_fuse_apply_long.
constexpr int kFuseShort = 16;

__global__ void __launch_bounds__(256)
k_fuse_apply(const __grid_constant__ FuseMapArgs a)
{
    const int addr2 = blockIdx.x * 256 + threadIdx.x;
    if (addr2 >= a.W * a.H) return;
    int k = a.head[addr2];
    if (k < 0) return;
    int src[kFuseShort], n = 0;
    for (; k >= 0 && n < kFuseShort; k = a.next[k]) {          // insertion into an ascending array
        int at = n++;
#pragma unroll
        for (int j = kFuseShort - 1; j > 0; j--)
            if (j <= at && src[j - 1] > k) { src[j] = src[j - 1]; at = j - 1; }
        src[at] = k;
    }
    if (k >= 0) { a.long_targets[atomicAdd(a.counters, 1)] = addr2; return; }
    FusePixel c = fuse_load(a, addr2);
#pragma unroll
    for (int j = 0; j < kFuseShort; j++)
        if (j < n) fuse_apply_one(a, src[j], c);
    fuse_store(a, addr2, c);
}

// Long lists (degenerate geometry: up to every previous point on one pixel): one CTA per such pixel copies the list
// into a power-of-two segment of the scratch area, sorts it with a bitonic network in global memory and one
// thread walks the chain of averages (inherently sequential: each source sees what the one before left).
__global__ void __launch_bounds__(1024)
k_fuse_apply_long(const __grid_constant__ FuseMapArgs a)
{
    __shared__ int seg_at, seg_n;
    const int n_long = a.counters[0];
    for (int w = blockIdx.x; w < n_long; w += gridDim.x) {
        const int addr2 = a.long_targets[w];
        if (threadIdx.x == 0) {
            int n = 0;
            for (int k = a.head[addr2]; k >= 0; k = a.next[k]) n++;
            int n2 = 1;
            while (n2 < n) n2 <<= 1;
            const int at = atomicAdd(a.counters + 1, n2);      // segments add up to less than 2 * W * H
            n = 0;
            for (int k = a.head[addr2]; k >= 0; k = a.next[k]) a.scratch[at + n++] = k;
            for (; n < n2; n++) a.scratch[at + n] = 0x7FFFFFFF;
            seg_at = at; seg_n = n2;
        }
        __syncthreads();
        int* seg = a.scratch + seg_at;
        const int n2 = seg_n;
        for (int size = 2; size <= n2; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = threadIdx.x; i < n2; i += 1024) {
                    const int j = i ^ stride;
                    if (j > i) {
                        const int x = seg[i], y = seg[j];
                        if ((x > y) == ((i & size) == 0)) { seg[i] = y; seg[j] = x; }
                    }
                }
                __syncthreads();
            }
        if (threadIdx.x == 0) {
            FusePixel c = fuse_load(a, addr2);
            for (int i = 0; i < n2 && seg[i] != 0x7FFFFFFF; i++) fuse_apply_one(a, seg[i], c);
            fuse_store(a, addr2, c);
        }
        __syncthreads();
    }
}

// ---- compaction in scan order: MODE 0 = previous points that were kept (:384, :389, :394), 1 = valid current points (:414-430)
template <int MODE>
__device__ __forceinline__ bool fuse_flag(const FuseMapArgs& a, int s)
{
    if (MODE == 0) return a.target[s] == kFuseKeep;
    const int u = s / a.H, v = s - u * a.H;
    return a.cD[v * a.W + u] > 0.f;
}
constexpr int kScanBlock = 1024;

template <int MODE>
__global__ void __launch_bounds__(256)
k_fuse_count(const __grid_constant__ FuseMapArgs a, int32_t* __restrict__ block_sums)
{
    __shared__ int warp_n[8];
    const int n = a.W * a.H, base = blockIdx.x * kScanBlock;
    int c = 0;
    for (int k = 0; k < kScanBlock / 256; k++) {
        const int s = base + k * 256 + threadIdx.x;
        c += s < n && fuse_flag<MODE>(a, s);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) warp_n[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) t += warp_n[w]; block_sums[blockIdx.x] = t; }
}

// exclusive prefix sum of block_sums[0..nb) in place by one CTA; the total goes to *total
__global__ void __launch_bounds__(1024)
k_fuse_scan_blocks(int32_t* block_sums, int nb, int32_t* total)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int v = i < nb ? block_sums[i] : 0;
        int incl = v;
        for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = warp_sums[lane];
            int wi = w;
            for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, off); if (lane >= off) wi += t; }
            warp_sums[lane] = wi - w;
        }
        __syncthreads();
        const int excl = carry + warp_sums[warp] + incl - v;
        if (i < nb) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

template <int MODE>
__global__ void __launch_bounds__(256)
k_fuse_scatter(const __grid_constant__ FuseMapArgs a, const int32_t* __restrict__ block_offsets, float4* __restrict__ points)
{
    __shared__ int warp_n[8];
    const int n = a.W * a.H, base = blockIdx.x * kScanBlock, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int at = block_offsets[blockIdx.x];
    for (int k = 0; k < kScanBlock / 256; k++) {                // 256 consecutive scan positions per step
        const int s = base + k * 256 + threadIdx.x;
        const bool f = s < n && fuse_flag<MODE>(a, s);
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_n[warp] = __popc(m);
        __syncthreads();
        int before = 0, all = 0;
        for (int w = 0; w < 8; w++) { if (w < warp) before += warp_n[w]; all += warp_n[w]; }
        if (f) {
            const int u = s / a.H, v = s - u * a.H, addr = v * a.W + u;
            const float4 p = MODE == 0 ? make_float4(a.pX[addr], a.pY[addr], a.pZ[addr], a.pI[addr])
                                       : make_float4(a.cX[addr], a.cY[addr], a.cZ[addr], a.cI[addr]);
            points[at + before + __popc(m & ((1u << lane) - 1))] = p;
        }
        at += all;
        __syncthreads();
    }
}

}  // namespace

// Matrix::inv of a 4x4 pose (libviso2/src/matrix.cpp:593-604 = eye(4).solve(A), :648-757): Gauss-Jordan elimination
// with full pivoting in double; false when a pivot is below 1e-20.  Host code (-ffp-contract=off).
static bool inv4_gauss_jordan(double A[4][4], double B[4][4])
{
    int ipiv[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) B[i][j] = i == j;
    int irow = 0, icol = 0;
    for (int i = 0; i < 4; i++) {
        double big = 0.0;
        for (int j = 0; j < 4; j++)
            if (ipiv[j] != 1)
                for (int k = 0; k < 4; k++)
                    if (ipiv[k] == 0 && fabs(A[j][k]) >= big) { big = fabs(A[j][k]); irow = j; icol = k; }
        ++ipiv[icol];
        if (irow != icol)
            for (int l = 0; l < 4; l++) {
                double t = A[irow][l]; A[irow][l] = A[icol][l]; A[icol][l] = t;
                t = B[irow][l]; B[irow][l] = B[icol][l]; B[icol][l] = t;
            }
        if (fabs(A[icol][icol]) < 1e-20) return false;
        const double pivinv = 1.0 / A[icol][icol];
        A[icol][icol] = 1.0;
        for (int l = 0; l < 4; l++) A[icol][l] *= pivinv;
        for (int l = 0; l < 4; l++) B[icol][l] *= pivinv;
        for (int ll = 0; ll < 4; ll++)
            if (ll != icol) {
                const double dum = A[ll][icol];
                A[ll][icol] = 0.0;
                for (int l = 0; l < 4; l++) A[ll][l] -= A[icol][l] * dum;
                for (int l = 0; l < 4; l++) B[ll][l] -= B[icol][l] * dum;
            }
    }
    return true;
}

size_t fuse_work_ints(int W, int H)
{
    const size_t n = (size_t)W * H;
    return 3 * n + (n + kScanBlock - 1) / kScanBlock + 8 + 2 * n + n / kFuseShort + 8 + 2;
}

// prev may be null (no previous map: only the current points are listed).  work = fuse_work_ints() int32.
// counts[0..1] (device memory) receive the numbers of previous / current points.
bool launch_fuse(int W, int H, const elas_b200_view& view, float* const* prev, float* const* cur, int32_t* work,
                 float* points_prev, float* points_curr, int32_t* counts, cudaStream_t s)
{
    FuseMapArgs a{};
    a.W = W; a.H = H; a.max_dist = view.max_dist;
    const size_t n = (size_t)W * H;
    const int nb = (int)((n + kScanBlock - 1) / kScanBlock), blocks = (int)((n + 255) / 256);
    a.target = work; a.head = work + n; a.next = work + 2 * n;
    int32_t* block_sums = work + 3 * n;
    a.scratch = block_sums + nb + 8; a.long_targets = a.scratch + 2 * n; a.counters = a.long_targets + n / kFuseShort + 8;
    a.cI = cur[0]; a.cD = cur[1]; a.cX = cur[2]; a.cY = cur[3]; a.cZ = cur[4];
    if (prev) {
        a.pI = prev[0]; a.pD = prev[1]; a.pX = prev[2]; a.pY = prev[3]; a.pZ = prev[4];
        double A[4][4] = {{0}}, Hi[4][4];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) A[r][c] = view.H[4 * r + c];
        A[3][3] = 1.0;
        if (!inv4_gauss_jordan(A, Hi)) return false;
        for (int j = 0; j < 4; j++) a.hfc2[j] = (float)Hi[2][j];
        const double K[3][3] = {{view.f, 0, view.cu}, {0, view.f, view.cv}, {0, 0, 1}};      // stereothread.cpp:450-455
        for (int i = 0; i < 3; i++)                                                              // Matrix::operator*, matrix.cpp:396-418
            for (int j = 0; j < 4; j++) {
                double acc = 0.0;
                for (int k = 0; k < 3; k++) acc += K[i][k] * Hi[k][j];
                a.pfc[i][j] = (float)acc;
            }
        cudaMemsetAsync(a.head, 0xFF, n * 4, s);
        cudaMemsetAsync(a.counters, 0, 8, s);
        k_fuse_project<<<blocks, 256, 0, s>>>(a);
        k_fuse_apply<<<blocks, 256, 0, s>>>(a);
        k_fuse_apply_long<<<32, 1024, 0, s>>>(a);
        k_fuse_count<0><<<nb, 256, 0, s>>>(a, block_sums);
        k_fuse_scan_blocks<<<1, 1024, 0, s>>>(block_sums, nb, counts);
        k_fuse_scatter<0><<<nb, 256, 0, s>>>(a, block_sums, reinterpret_cast<float4*>(points_prev));
        count_launch(6);
    } else {
        cudaMemsetAsync(counts, 0, 4, s);
    }
    k_fuse_count<1><<<nb, 256, 0, s>>>(a, block_sums);
    k_fuse_scan_blocks<<<1, 1024, 0, s>>>(block_sums, nb, counts + 1);
    k_fuse_scatter<1><<<nb, 256, 0, s>>>(a, block_sums, reinterpret_cast<float4*>(points_curr));
    count_launch(3);
    return true;
}

}  // namespace elasb
