// the two per-pixel loops that consume D1 inside stereomapper's StereoThread,
// run where D1 already is -- in HBM -- instead of after a device->host copy:
//   k_colormap   HSV colour map of min(D1/200, 1)                       stereothread.cpp:116-147
//   k_reproject  back-projection z = f*b/d, x = (u-cu)*b/d, y = (v-cv)*b/d through the 3x4 pose,
//                intensity image I1/255 with the border gain ramp       stereothread.cpp:180-255
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

}  // namespace elasb
