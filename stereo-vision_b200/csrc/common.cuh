// Shared device/host definitions of the B200 dense-stereo path.
// Vocabulary follows the reference (libelas/src/elas.{h,cpp}): descriptors, candidate lattice,
// support points, triangles, disparity planes, candidate grid, disparity maps.
#pragma once
#include <cuda.h>            // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/elas_b200.h"

namespace elasb {

constexpr int kInvalid = -10;        // elas.cpp:977-980: disparity maps are pre-filled with -10

// candidate grid, list form: per cell kGridListStride uint16 = {count, d0, d1, ..., sentinel}.
// The entry after the last one is disp_max+1, a disparity no pixel can take, so the matching kernel
// walks the list two entries at a time without a remainder case.
// The stride is 144 bytes, not 128: lanes of a warp that sit in neighbouring cells then read their
// k-th entries from different shared-memory banks.
constexpr int kGridListStride = 72;
constexpr int kGridListCap = 62;

constexpr int kRasterBandRows = 32;   // k_raster work unit: 32 columns x this many rows of a triangle's bounding box

struct FrameGeom;
// triangle-id maps are int32 [H][map_pitch]: rows padded to 16 bytes so the matching kernel can TMA them
__host__ __device__ inline int map_pitch_of(int W) { return (W + 3) & ~3; }

// Everything a kernel needs to know about one frame geometry + parameter block.
struct FrameGeom {
    int W, H;            // image size (dims[0], dims[1])
    int bpl;             // padded image pitch, elas.cpp:37
    int Dw, Dh;          // disparity map size (W/2 x H/2 with subsampling, elas.h:83-85)
    int step;            // candidate lattice stride (5, or 6 with subsampling, elas.cpp:453-457)
    int Wc, Hc;          // candidate lattice size, elas.cpp:460-463
    int gw, gh;          // candidate grid size, elas.cpp:98-99
    int gwords;          // 32-bit words per grid cell bitmask = ceil((disp_max+1)/32)
    int dn;              // disp_max + 1 (number of disparities, elas.cpp:819)
    int plane_radius;    // elas.cpp:993
};

// One triangle of one image, prepared by the host stage in the reference's own float arithmetic
// (elas.cpp:1006-1072): edge lines v = a*u + b, integer corner columns, plane and validity.
struct __align__(16) TriRaster {
    float ACa, ACb, ABa, ABb;                 // 16 B: edge lines A-C and A-B
    float BCa, BCb; int uA, uB;               // 16 B: edge line B-C, (int32_t)A_u, (int32_t)B_u after the sort by u
    float pa, pb, pc; int valid;              // 16 B: plane of this image; valid = 2 if |plane_a| < 0.7 && |plane_d| < 0.7 (elas.cpp:1072), else 0
    int   uC, pad0, pad1, pad2;               // 16 B: (int32_t)C_u
};
static_assert(sizeof(TriRaster) == 64, "TriRaster is one 64-byte record");

// Per-frame results of the mesh stage, in device memory (kernels downstream size their loops from it) and
// copied to pinned host memory with the frame's maps (status, statistics).
struct FrameHeader {
    int32_t n_support;       // support points (elas.cpp:505-517); < 3: the frame has no triangulation (elas.cpp:69-75)
    int32_t n_tri[2];        // triangles of the left / right image
    int32_t n_units[2];      // scan-conversion work units of each image
    int32_t ovf_from[2];     // triangles [ovf_from, n_tri) did not fit the unit list: one warp scan-converts each on its own
    int32_t status;          // 0, or an ELAS_B200_E_* code raised on the device
};
static_assert(sizeof(FrameHeader) == 32, "FrameHeader is 32 bytes");

// Element strides between consecutive frames of a group (frames batched per launch, blockIdx.z / .y)
struct GroupStrides {
    size_t img, desc, dcan, support, tri, units, traster, planes, scratch, grid, lists, map, D, mesh_scratch, lat_work, seg_nodes;
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned sad16(const uint4& a, const uint4& b)
{
    // VABSDIFF4.U8.ACC x4: sum of |a_i - b_i| over 16 bytes (_mm_sad_epu8 both halves, elas.cpp:787-789);
    // two accumulation chains of two so the XU-pipe latencies overlap
    unsigned s = 0, t = 0;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(s) : "r"(a.x), "r"(b.x));
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(t) : "r"(a.y), "r"(b.y));
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(s) : "r"(a.z), "r"(b.z));
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(t) : "r"(a.w), "r"(b.w));
    return s + t;
}

// same, accumulating into two running sums (several blocks share the chains)
__device__ __forceinline__ void sad16_acc(const uint4& a, const uint4& b, unsigned& s, unsigned& t)
{
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(s) : "r"(a.x), "r"(b.x));
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(t) : "r"(a.y), "r"(b.y));
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(s) : "r"(a.z), "r"(b.z));
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(t) : "r"(a.w), "r"(b.w));
}

// sum |desc[i] - 128| (elas.cpp:358-362, :851-855)
__device__ __forceinline__ unsigned texture16(const uint4& a)
{
    const uint4 mid = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
    return sad16(a, mid);
}
#endif  // __CUDACC__

__host__ __device__ inline int map_pitch(const FrameGeom& g) { return map_pitch_of(g.W); }

// ---- kernel launchers (one translation unit each) --------------------------------------------
// Every launcher takes the pointers of FRAME 0 of a frame group, the group's strides and the number of frames
// batched into the launch (a grid dimension): one launch chain serves up to kMaxGroupFrames frames.
constexpr int kMaxGroupFrames = 16;
struct OutTable { float* p[kMaxGroupFrames]; };      // where each frame's finished map goes (group buffer or the caller's)
OutTable out_table(float* base, size_t stride, int n_frames);

// K1  Sobel + descriptor, both images (filter.cpp:408-416, descriptor.cpp:48-121)
// (the image tiles arrive by tensor-map TMA: tm1 / tm2 describe the group's two image buffers as [frames][H][bpl] uint8
// with the box descriptor_tile_box() = {columns, rows}; elas_b200.cu encodes them once per group)
void descriptor_tile_box(int box[2]);
void launch_descriptor(const FrameGeom& g, int half, const CUtensorMap& tm1, const CUtensorMap& tm2,
                       uint4* desc1, uint4* desc2, const GroupStrides& st, int n_frames, cudaStream_t s);
// K2  support matching on the candidate lattice, forward + reverse (elas.cpp:322-445, :471-493)
void launch_support(const FrameGeom& g, const elas_b200_params& p, const uint4* desc1,
                    const uint4* desc2, int16_t* dcan, const GroupStrides& st, int n_frames, cudaStream_t s);
// K3  lattice filters + support list (elas.cpp:174-279, :496-517), one CTA per frame; fills hdr[f].n_support
bool mesh_on_device(const FrameGeom& g, const elas_b200_params& p);
size_t lattice_work_ints(const FrameGeom& g);
void launch_lattice(const FrameGeom& g, const elas_b200_params& p, const int16_t* dcan_raw, int16_t* dcan, int16_t* dcan_incon,
                    int32_t* support, int32_t* work, FrameHeader* hdr, const GroupStrides& st, int n_frames, cudaStream_t s);
// K4  Delaunay triangulation of both images + scan-conversion work units (elas.cpp:534-600, triangle.cpp), one CTA
//     per frame and image; fills hdr[f].n_tri / n_units / ovf_from
void launch_delaunay(const FrameGeom& g, const int32_t* support, int32_t* tri1, int32_t* tri2, int32_t* units1, int32_t* units2,
                     int unit_cap, FrameHeader* hdr, int32_t* scratch, const GroupStrides& st, int n_frames, cudaStream_t s);
// K5 + K6 scatter: disparity planes + per-triangle raster records (elas.cpp:605-680, :1006-1072) and the
// support points' d-1..d+1 marks in the candidate-grid scatter planes (elas.cpp:697-727), one launch.
// scratch = this frame's scatter planes [2][gh*gw][gwords], all zero on entry.
void launch_planes_scatter(const FrameGeom& g, const elas_b200_params& p, const FrameHeader* hdr, const int32_t* support,
                           const int32_t* tri1, const int32_t* tri2, TriRaster* out1, TriRaster* out2,
                           float* planes1, float* planes2, uint32_t* scratch, const GroupStrides& st, int n_frames,
                           cudaStream_t s);
// K6 diffusion (elas.cpp:732-775) + triangle-id maps by scan conversion with last-writer-wins
// (elas.cpp:1074-1114), one launch.  scratch_next (the group's other scatter buffer) is zeroed for the
// next frame.  Map entries are tag_bits | triangle index (see k_grid_raster.cu).
void launch_diffuse_raster(const FrameGeom& g, int subsampling, const FrameHeader* hdr, const uint32_t* scratch,
                           uint32_t* scratch_next, uint32_t* grid1, uint32_t* grid2, uint16_t* lists1, uint16_t* lists2,
                           const TriRaster* tri1, const TriRaster* tri2, const int32_t* units1, const int32_t* units2,
                           int32_t* map1, int32_t* map2, int tag_bits, const GroupStrides& st, int n_frames, cudaStream_t s);
// K7  dense matching, both images (elas.cpp:814-955, :960-1118), n_frames frames per launch (blockIdx.z):
// every pointer addresses frame 0 of a group, *_stride = elements between consecutive frames
struct MatchBuffers {
    const uint4* desc[2];
    const TriRaster* tri[2];
    const int32_t* map[2];
    const uint32_t* grid[2];
    const uint16_t* lists[2];
    const int32_t* prior;
    const int32_t* prior_host;           // the same table in host memory (its first entries travel as kernel arguments)
    float* D[2];
    size_t desc_stride, tri_stride, map_stride, grid_stride, lists_stride, D_stride;
    int rows_per_cta;                    // set by launch_matching
};
void launch_matching(const FrameGeom& g, const elas_b200_params& p, const MatchBuffers& b, int n_frames,
                     int map_tag_bits, int map_tag_shift, cudaStream_t s);
size_t matching_smem_bytes(const FrameGeom& g, const elas_b200_params& p);
size_t support_smem_bytes(const FrameGeom& g, const elas_b200_params& p);
// K8  left/right consistency (elas.cpp:1122-1204); O2 = per-frame destination of the checked right map
void launch_lr_check(const FrameGeom& g, const elas_b200_params& p, const float* D1, const float* D2,
                     float* O1, const OutTable& O2, size_t D_stride, int n_frames, cudaStream_t s);
// K9  speckle removal (elas.cpp:1208-1326): tile-local labelling in shared memory + a global union-find on the few
// small components that touch tile borders (k_postproc.cu).  label = one int32 per pixel, nodes = segment_node_ints()
// int32 per frame.  apply = false: stop when the sizes are known; launch_post_fused then applies them
size_t segment_node_ints(const FrameGeom& g);
void launch_segments(const FrameGeom& g, const elas_b200_params& p, float* D, int32_t* label, int32_t* nodes,
                     size_t D_stride, size_t nodes_stride, int n_frames, cudaStream_t s, bool apply = true);
// K8 with both rows staged in shared memory.  The checked right map may leave narrowed for the copy to the host
// (mode 0: not at all; see narrow_d2_layout in k_postproc.cu)
struct NarrowD2 { void* base; int mode; int mask_words_per_row; size_t stride_bytes, mask_offset, bytes; };
NarrowD2 narrow_d2_layout(const FrameGeom& g, int mode, void* base, size_t stride_bytes);
bool lr_rows_fusable(const FrameGeom& g);
void launch_lr_rows(const FrameGeom& g, const elas_b200_params& p, const float* D1, const float* D2,
                    float* O1, const OutTable& O2, const NarrowD2& narrow, size_t D_stride, int n_frames, cudaStream_t s);
// K9 apply + K10 + K11 in one tiled kernel (ipol_gap_width <= 3, no add_corners); out must not alias in
bool post_fusable(const elas_b200_params& p);
void launch_post_fused(const FrameGeom& g, const elas_b200_params& p, const float* in, const int32_t* label,
                       const int32_t* nodes, size_t nodes_stride, const OutTable& out, float* dump_seg, float* dump_gap,
                       size_t D_stride, int n_frames, cudaStream_t s);
// K10 gap interpolation (elas.cpp:1330-1530), K11 adaptive mean (elas.cpp:1535-1754), K12 median (elas.cpp:1758-1838):
// in place on D with one scratch plane per frame
void launch_gap(const FrameGeom& g, const elas_b200_params& p, float* D, float* tmp, size_t D_stride, size_t tmp_stride,
                int n_frames, cudaStream_t s);
void launch_adaptive_mean(const FrameGeom& g, const elas_b200_params& p, float* D, float* tmp, size_t D_stride,
                          size_t tmp_stride, int n_frames, cudaStream_t s);
void launch_median(const FrameGeom& g, float* D, float* tmp, size_t D_stride, size_t tmp_stride, int n_frames, cudaStream_t s);

// D1's consumers in StereoThread: colour map (stereothread.cpp:116-147), back-projection (:180-255)
void launch_colormap(int n, const float* D1, float* out, cudaStream_t s);
void launch_reproject(int W, int H, const uint8_t* img, int pitch, const float* D1, const elas_b200_view& view,
                      float* I, float* D, float* X, float* Y, float* Z, cudaStream_t s);

// Opt-in to more than 48 KB of dynamic shared memory.  The attribute belongs to (kernel, DEVICE): a
// process may own contexts on several GPUs, so it is set once per device the kernel is launched on
// (idempotent, so concurrent first launches from several worker threads are harmless).
template <class Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, unsigned long long* done_mask)
{
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(done_mask, __ATOMIC_ACQUIRE) & bit) return cudaSuccess;
    err = bytes > 0 ? cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) : cudaSuccess;
    // Every kernel of the path asks for the same L1 / shared-memory split (all shared): chains of different frame
    // groups run concurrently on different streams, and CTAs of kernels that want different splits cannot share an SM
    if (err == cudaSuccess) err = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (err == cudaSuccess) __atomic_fetch_or(done_mask, bit, __ATOMIC_RELEASE);
    return err;
}
// kernels without dynamic shared memory: the carve-out preference only
#define ELASB_PREPARE_KERNEL(kernel)                                                        \
    do {                                                                                    \
        static unsigned long long prepared__ = 0;                                           \
        if (::elasb::ensure_dynamic_smem(kernel, 0, &prepared__) != cudaSuccess) return;    \
    } while (0)

// fusion of the current map with the previous one (stereothread.cpp:290-437); maps are device pointers {I, D, X, Y, Z}
size_t fuse_work_ints(int W, int H);
bool launch_fuse(int W, int H, const elas_b200_view& view, float* const* prev, float* const* cur, int32_t* work,
                 float* points_prev, float* points_curr, int32_t* counts, cudaStream_t s);

// feature filters of libviso2's Matcher (filter.cpp:474-530), one launch
void launch_matcher_filters(const uint8_t* in, int w, int h, uint8_t* du, uint8_t* dv, int16_t* f1, int16_t* f2, cudaStream_t s);

// number of kernel launches issued through the launchers above (process-wide, relaxed)
long long launches_issued();
void count_launch(int n = 1);

}  // namespace elasb
