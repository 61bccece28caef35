// K1: 3x3 Sobel responses + 16-byte descriptor, fused, both images in one launch.
//
// Reference: filter::sobel3x3 (filter.cpp:408-416 = convolve_cols_3x3 :372-405 +
// convolve_101_row_3x3_16bit :227-267 + convolve_121_row_3x3_16bit :176-222) followed by
// Descriptor::createDescriptor (descriptor.cpp:48-121).  The reference materialises two int16 and
// two uint8 planes per image; here one CTA stages a 64x16 pixel tile (+3 halo) of the uint8 image in
// shared memory -- ONE tensor-map TMA copy (cp.async.bulk.tensor.3d: column, row, frame of the group's image
// buffer; the halo outside the image arrives as zeros, no bounds tests) -- derives du/dv for the tile
// (+2 / +1 halo) in shared memory and writes each pixel's 16 descriptor bytes with a single 16-byte store.
// HBM traffic: 1 B/px in, 16 B/px out.
//
// Both stages work on groups of four horizontally adjacent pixels per thread: the Sobel stage shares
// the column sums S = I(v-1)+2I(v)+I(v+1) and T = I(v-1)-I(v+1) between neighbours (6 columns for 4
// outputs), the gather stage reads each du/dv row it needs as one 8-byte window (two aligned 32-bit
// loads) and assembles the 16 output words with byte permutes (PRMT) -- 16 loads and ~40 permutes for
// four pixels instead of 64 byte loads and 48 shift/or.
//
// Border rule: pixels outside v in [3,H-3), u in [3,W-3) are zero (the reference leaves
// them uninitialised); with half resolution only rows 4,6,8,.. < H-3 are computed (descriptor.cpp:54).
#include "common.cuh"

namespace elasb {
namespace {

constexpr int TW = 64, TH = 16;            // output tile
constexpr int IW = TW + 32, IH = TH + 6;   // image tile: cols u0-16 .. u0+TW+15 (a TMA box starts and ends on 16-byte boundaries of the row), rows v0-3 .. v0+TH+2
constexpr int IX = 16;                     // tile column of image column u0
constexpr int GW = (TW + 4 + 3) / 4;       // du/dv groups of four columns per row: cols u0-2 .. u0-2+4*GW-1
constexpr int UP = 4 * GW + 4;             // du/dv tile pitch in bytes (cols u0-2 ..), a multiple of 4, plus one spare word
constexpr int UH = TH + 4;                 // du rows v0-2 .. v0+TH+1; dv is kept for the same rows (v0-1 .. v0+TH used)

__device__ __forceinline__ int sat8(int x) { return min(max(x, 0), 255); }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
// byte i (0..7) of the 8-byte window {lo, hi}
struct Win { uint32_t lo, hi; };
__device__ __forceinline__ Win load_win(const uint8_t* row, int idx)      // idx multiple of 4
{
    Win w;
    w.lo = *reinterpret_cast<const uint32_t*>(row + idx);
    w.hi = *reinterpret_cast<const uint32_t*>(row + idx + 4);
    return w;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256)
k_descriptor(FrameGeom g, int half, const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2,
             uint4* __restrict__ desc1, uint4* __restrict__ desc2, size_t desc_stride)
{
    __shared__ __align__(128) uint8_t sI[IH][IW];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(16) uint8_t sU[UH][UP];
    __shared__ __align__(16) uint8_t sV[UH][UP];
    __shared__ uint4 sO[TH][TW];               // finished descriptors, 16-byte chunks XOR-swizzled within 128-byte lines

    // blockIdx.z = 2 * frame + image
    uint4* __restrict__ desc = ((blockIdx.z & 1) ? desc2 : desc1) + (size_t)(blockIdx.z >> 1) * desc_stride;
    const int u0 = blockIdx.x * TW, v0 = blockIdx.y * TH;
    const int tid = threadIdx.x;

    // stage the image tile: one TMA tensor copy of the box [u0-16, u0+80) x [v0-3, v0+19) x {frame} (the innermost start
    // coordinate must be a multiple of 16 bytes: tools/micro/tma_tile_probe.cu); whatever lies
    // outside the padded image [0, bpl) x [0, H) is filled with zeros by the copy engine
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(IH * IW) : "memory");
        // the descriptor must be addressed where it lies in the kernel's parameter space (no pointer select: that would
        // make the compiler copy the map into local memory, which the TMA unit cannot read)
        if (blockIdx.z & 1)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(&sI[0][0])), "l"(&tm2), "r"(u0 - IX), "r"(v0 - 3), "r"((int)(blockIdx.z >> 1)), "r"(smem_u32(&bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(&sI[0][0])), "l"(&tm1), "r"(u0 - IX), "r"(v0 - 3), "r"((int)(blockIdx.z >> 1)), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(&bar)) : "memory");

    // du(u,v) = sat8(((S(u-1,v) - S(u+1,v)) >> 2) + 128),  S = I(v-1) + 2 I(v) + I(v+1)
    // dv(u,v) = sat8(((T(u-1,v) + 2 T(u,v) + T(u+1,v)) >> 2) + 128),  T = I(v-1) - I(v+1)
    // one item = four columns u0-2+4q .. +3 of row v0-2+r: six columns of S and T
    for (int i = tid; i < UH * GW; i += 256) {
        const int r = i / GW, q = i - r * GW;
        const int ir = r + 1, ic = 4 * q + IX - 3;     // column u0-2+4q-1 sits at sI column (u0-3+4q) - (u0-IX)
        int S[6], T[6];
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const int a = sI[ir - 1][ic + k], b = sI[ir][ic + k], c = sI[ir + 1][ic + k];
            S[k] = a + 2 * b + c;
            T[k] = a - c;
        }
        uint32_t du = 0, dv = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            du |= (uint32_t)sat8(((S[k] - S[k + 2]) >> 2) + 128) << (8 * k);
            dv |= (uint32_t)sat8(((T[k] + 2 * T[k + 1] + T[k + 2]) >> 2) + 128) << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(&sU[r][4 * q]) = du;
        *reinterpret_cast<uint32_t*>(&sV[r][4 * q]) = dv;
    }
    __syncthreads();

    // gather 12 du + 4 dv taps (descriptor.cpp:101-116) for four pixels (u0+4q .. +3, v0+r); tile column of
    // pixel k = 4q+2+k, so the window that starts at tile column 4q holds columns u-2 .. u+5 of the group
    {
        const int r = tid >> 4, q = tid & 15;
        const int v = v0 + r, ub = u0 + 4 * q;
        const int ur = r + 2;                                       // this row in sU / sV
        const Win um2 = load_win(sU[ur - 2], 4 * q), um1 = load_win(sU[ur - 1], 4 * q), uc = load_win(sU[ur], 4 * q),
                  up1 = load_win(sU[ur + 1], 4 * q), up2 = load_win(sU[ur + 2], 4 * q);
        const Win vm1 = load_win(sV[ur - 1], 4 * q), vc = load_win(sV[ur], 4 * q), vp1 = load_win(sV[ur + 1], 4 * q);
        bool row_ok = v >= 3 && v < g.H - 3;
        if (half) row_ok = row_ok && v >= 4 && !(v & 1);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int u = ub + k;
            uint4 o = make_uint4(0, 0, 0, 0);
            if (row_ok && u >= 3 && u < g.W - 3) {
                // window byte index of column u+dx is k+2+dx; prmt selector nibbles pick bytes 0-3 of the
                // first operand and 4-7 of the second
                const uint32_t c = k + 2;
                // a single byte c of a window comes from its low word (c < 4) or its high word: fixed per k
#define ELASB_WORD(w) (c < 4 ? (w).lo : (w).hi)
                // o.x = du(0,-2), du(-2,-1), du(0,-1), du(2,-1)
                const uint32_t x123 = prmt(um1.lo, um1.hi, (c - 2) << 4 | c << 8 | (c + 2) << 12);      // bytes 1..3
                o.x = prmt(ELASB_WORD(um2), x123, (c & 3) | 5 << 4 | 6 << 8 | 7 << 12);
                // o.y = du(-1,0), du(0,0), du(0,0), du(1,0)
                o.y = prmt(uc.lo, uc.hi, (c - 1) | c << 4 | c << 8 | (c + 1) << 12);
                // o.z = du(-2,1), du(0,1), du(2,1), du(0,2)
                const uint32_t z012 = prmt(up1.lo, up1.hi, (c - 2) | c << 4 | (c + 2) << 8);             // bytes 0..2
                o.z = prmt(z012, ELASB_WORD(up2), 0 | 1 << 4 | 2 << 8 | (4 + (c & 3)) << 12);
                // o.w = dv(0,-1), dv(-1,0), dv(1,0), dv(0,1)
                const uint32_t w03 = prmt(ELASB_WORD(vm1), ELASB_WORD(vp1), (c & 3) | (4 + (c & 3)) << 12);  // bytes 0 and 3
                const uint32_t w12 = prmt(vc.lo, vc.hi, (c - 1) << 4 | (c + 1) << 8);                    // bytes 1..2
                o.w = prmt(w03, w12, 0 | 5 << 4 | 6 << 8 | 3 << 12);
#undef ELASB_WORD
            }
            // a thread's four descriptors are 64 bytes apart from its neighbour's: stored directly, every
            // 32-byte sector would be written in two halves by two instructions.  They go through shared
            // memory instead (chunk index XOR-swizzled: conflict-free both ways) and leave as whole rows.
            const int pcol = 4 * q + k;
            sO[r][pcol ^ ((pcol >> 3) & 7)] = o;
        }
    }
    __syncthreads();
    for (int i = tid; i < TH * TW; i += 256) {
        const int r = i / TW, pcol = i - r * TW;
        const int v = v0 + r, u = u0 + pcol;
        if (v < g.H && u < g.W) desc[(size_t)v * g.W + u] = sO[r][pcol ^ ((pcol >> 3) & 7)];
    }
}

}  // namespace

void descriptor_tile_box(int box[2]) { box[0] = IW; box[1] = IH; }

void launch_descriptor(const FrameGeom& g, int half, const CUtensorMap& tm1, const CUtensorMap& tm2,
                       uint4* desc1, uint4* desc2, const GroupStrides& st, int n_frames, cudaStream_t s)
{
    dim3 grid((g.W + TW - 1) / TW, (g.H + TH - 1) / TH, 2 * n_frames);
    ELASB_PREPARE_KERNEL(k_descriptor);
    k_descriptor<<<grid, 256, 0, s>>>(g, half, tm1, tm2, desc1, desc2, st.desc);
    count_launch();
}

}  // namespace elasb
