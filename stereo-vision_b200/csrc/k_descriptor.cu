// K1: 3x3 Sobel responses + 16-byte descriptor, fused, both images in one launch.
//
// Reference: filter::sobel3x3 (filter.cpp:408-416 = convolve_cols_3x3 :372-405 +
// convolve_101_row_3x3_16bit :227-267 + convolve_121_row_3x3_16bit :176-222) followed by
// Descriptor::createDescriptor (descriptor.cpp:48-121).  The reference materialises two int16 and
// two uint8 planes per image; here one CTA stages a 64x16 pixel tile (+3 halo) of the uint8 image in
// shared memory, derives du/dv for the tile (+2 / +1 halo) in shared memory and writes each
// pixel's 16 descriptor bytes with a single 16-byte store.  HBM traffic: 1 B/px in, 16 B/px out.
//
// Border rule: pixels outside v in [3,H-3), u in [3,W-3) are zero (the reference leaves
// them uninitialised); with half resolution only rows 4,6,8,.. < H-3 are computed (descriptor.cpp:54).
#include "common.cuh"

namespace elasb {
namespace {

constexpr int TW = 64, TH = 16;            // output tile
constexpr int IW = TW + 8, IH = TH + 6;    // image tile: cols u0-4 .. u0+TW+3 (word aligned), rows v0-3 .. v0+TH+2
constexpr int UW = TW + 4, UH = TH + 4;    // du tile: cols u0-2 .. , rows v0-2 ..
constexpr int VW = TW + 4, VH = TH + 2;    // dv tile: cols u0-1 .. u0+TW (pitch padded), rows v0-1 ..

__device__ __forceinline__ int sat8(int x) { return min(max(x, 0), 255); }

__global__ void __launch_bounds__(256)
k_descriptor(FrameGeom g, int half, const uint8_t* __restrict__ img1, const uint8_t* __restrict__ img2,
             uint4* __restrict__ desc1, uint4* __restrict__ desc2)
{
    __shared__ __align__(16) uint8_t sI[IH][IW];
    __shared__ uint8_t sU[UH][UW];
    __shared__ uint8_t sV[VH][VW];

    const uint8_t* __restrict__ img = blockIdx.z ? img2 : img1;
    uint4* __restrict__ desc = blockIdx.z ? desc2 : desc1;
    const int u0 = blockIdx.x * TW, v0 = blockIdx.y * TH;
    const int tid = threadIdx.x;

    // stage the image tile, one aligned 32-bit word per load; outside the padded image = 0
    for (int i = tid; i < IH * (IW / 4); i += 256) {
        int r = i / (IW / 4), cw = i % (IW / 4);
        int v = v0 - 3 + r, u = u0 - 4 + 4 * cw;
        uint32_t w = 0;
        if (v >= 0 && v < g.H && u >= 0 && u < g.bpl)
            w = *reinterpret_cast<const uint32_t*>(img + (size_t)v * g.bpl + u);
        *reinterpret_cast<uint32_t*>(&sI[r][4 * cw]) = w;
    }
    __syncthreads();

    // du(u,v) = sat8(((S(u-1,v) - S(u+1,v)) >> 2) + 128),  S = I(v-1) + 2 I(v) + I(v+1)
    for (int i = tid; i < UH * UW; i += 256) {
        int r = i / UW, c = i % UW;               // (u0-2+c, v0-2+r) -> sI row r+1, col c+2
        int ir = r + 1, ic = c + 2;
        int Sl = sI[ir - 1][ic - 1] + 2 * sI[ir][ic - 1] + sI[ir + 1][ic - 1];
        int Sr = sI[ir - 1][ic + 1] + 2 * sI[ir][ic + 1] + sI[ir + 1][ic + 1];
        sU[r][c] = (uint8_t)sat8(((Sl - Sr) >> 2) + 128);
    }
    // dv(u,v) = sat8(((T(u-1,v) + 2 T(u,v) + T(u+1,v)) >> 2) + 128),  T = I(v-1) - I(v+1)
    for (int i = tid; i < VH * (TW + 2); i += 256) {
        int r = i / (TW + 2), c = i % (TW + 2);   // (u0-1+c, v0-1+r) -> sI row r+2, col c+3
        int ir = r + 2, ic = c + 3;
        int Tl = sI[ir - 1][ic - 1] - sI[ir + 1][ic - 1];
        int Tc = sI[ir - 1][ic] - sI[ir + 1][ic];
        int Tr = sI[ir - 1][ic + 1] - sI[ir + 1][ic + 1];
        sV[r][c] = (uint8_t)sat8(((Tl + 2 * Tc + Tr) >> 2) + 128);
    }
    __syncthreads();

    // gather 12 du + 4 dv taps (descriptor.cpp:101-116), one 16-byte store per pixel
    for (int i = tid; i < TW * TH; i += 256) {
        int r = i / TW, c = i % TW;
        int u = u0 + c, v = v0 + r;
        if (u >= g.W || v >= g.H) continue;
        uint4 o = make_uint4(0, 0, 0, 0);
        bool inside = v >= 3 && v < g.H - 3 && u >= 3 && u < g.W - 3;
        if (half) inside = inside && v >= 4 && !(v & 1);
        if (inside) {
            const int ur = r + 2, uc = c + 2;     // this pixel in sU
            const int vr = r + 1, vc = c + 1;     // this pixel in sV
            o.x = sU[ur - 2][uc] | (sU[ur - 1][uc - 2] << 8) | (sU[ur - 1][uc] << 16) | (sU[ur - 1][uc + 2] << 24);
            o.y = sU[ur][uc - 1] | (sU[ur][uc] << 8) | (sU[ur][uc] << 16) | (sU[ur][uc + 1] << 24);
            o.z = sU[ur + 1][uc - 2] | (sU[ur + 1][uc] << 8) | (sU[ur + 1][uc + 2] << 16) | (sU[ur + 2][uc] << 24);
            o.w = sV[vr - 1][vc] | (sV[vr][vc - 1] << 8) | (sV[vr][vc + 1] << 16) | (sV[vr + 1][vc] << 24);
        }
        desc[(size_t)v * g.W + u] = o;
    }
}

}  // namespace

void launch_descriptor(const FrameGeom& g, int half, const uint8_t* img1, const uint8_t* img2,
                       uint4* desc1, uint4* desc2, cudaStream_t s)
{
    dim3 grid((g.W + TW - 1) / TW, (g.H + TH - 1) / TH, 2);
    k_descriptor<<<grid, 256, 0, s>>>(g, half, img1, img2, desc1, desc2);
    count_launch();
}

}  // namespace elasb
