// C ABI (include/elas_b200.h) + per-device context of the B200 dense-stereo path.
//
// A context owns `n_slots` frame slots.  A slot is everything one frame needs -- CUDA streams, all
// device buffers (images, descriptors, tables, triangle-id maps, disparity maps, scratch), pinned host
// memory for the candidate lattice and the small tables, and a HostStage -- allocated once: nothing is
// allocated per frame (the reference mallocs ~20 buffers per Elas::process call).  A frame is
//        GPU phase A (images in, K1 descriptors, K2 support search -> lattice in pinned host memory)
//     -> host stage  (lattice filters, Delaunay, raster units; host_stage.cc)
//     -> GPU phase B (tables in, planes, grid, triangle-id maps, K7 matching, K8-K12, maps out)
// Frames are independent (stereothread.cpp:113 builds a fresh Elas per frame), so the batch calls keep
// many slots in flight: a pool of host workers (not tied to slots) advances whichever slot can move,
// see worker_main().
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <emmintrin.h>
#include <sched.h>

#include "common.cuh"
#include "host_stage.h"
#include "host_widen.h"

namespace elasb {

static std::atomic<long long> g_launches{0};
long long launches_issued() { return g_launches.load(std::memory_order_relaxed); }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace elasb

using namespace elasb;

#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t err__ = (expr);                                                                \
        if (err__ != cudaSuccess) {                                                                \
            std::fprintf(stderr, "elas_b200: %s failed at %s:%d: %s\n", #expr, __FILE__, __LINE__, \
                         cudaGetErrorString(err__));                                               \
            return ELAS_B200_E_CUDA;                                                               \
        }                                                                                          \
    } while (0)

namespace {

struct StageTimer {
    std::vector<std::pair<std::string, cudaEvent_t>> marks;   // event recorded AFTER the named stage
    cudaEvent_t begin = nullptr;
    std::vector<std::pair<std::string, float>> last;          // (stage, ms)
};

// One frame of a call: the caller's four buffers
struct FrameIO {
    const uint8_t* I1; const uint8_t* I2;
    float* D1; float* D2;
    int bytes_per_line; bool device_io;
};

// A frame group: everything up to `cap` frames need, allocated once.  The frames of a group go through ONE
// launch chain (every kernel takes the frame as a grid dimension), on the group's stream.
struct Group {
    int cap = 1;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;        // disparity maps go out here
    cudaEvent_t ev_done = nullptr;             // kernels finished
    cudaEvent_t ev_out = nullptr;              // the group's maps and headers are in the caller's / pinned buffers
    // device, [cap] frames each (strides: elas_b200_ctx::st)
    uint8_t* d_img[2] = {nullptr, nullptr};
    CUtensorMap tm_img[2];                     // d_img[k] as a [frames][H][bpl] uint8 tensor for K1's tile loads (TMA)
    uint4* d_desc[2] = {nullptr, nullptr};
    int16_t* d_dcan_raw = nullptr;             // K2's candidate lattice
    int16_t* d_dcan = nullptr;                 // after the lattice filters
    int16_t* d_dcan_incon = nullptr;           // after the inconsistency filter (allocated when stages are captured)
    int32_t* d_support = nullptr;              // (u,v,d) triples
    int32_t* d_tri[2] = {nullptr, nullptr};    // (c1,c2,c3) triples
    int32_t* d_units[2] = {nullptr, nullptr};  // scan-conversion work units
    int32_t* d_mesh_scratch = nullptr;         // k_delaunay working memory beyond shared memory
    int32_t* d_lat_work = nullptr;             // k_lattice: supporter counts + two work lists
    FrameHeader* d_hdr = nullptr;
    TriRaster* d_traster[2] = {nullptr, nullptr};
    float* d_planes[2] = {nullptr, nullptr};   // (t1a,t1b,t1c,t2a,t2b,t2c) per triangle
    uint32_t* d_grid_scratch = nullptr;
    uint32_t* d_grid[2] = {nullptr, nullptr};      // candidate grid, bitmask form [gh*gw][gwords]
    uint16_t* d_lists[2] = {nullptr, nullptr};     // candidate grid, list form [gh*gw][kGridListStride]
    int32_t* d_map[2] = {nullptr, nullptr};
    float* d_raw[2] = {nullptr, nullptr};      // K7 output; plane 0 is reused for the final left map
    float* d_D[2] = {nullptr, nullptr};        // after the L/R check and post-processing
    float* d_tmp = nullptr;                    // two planes of scratch per frame
    int32_t* d_seg_label = nullptr;            // speckle removal: label per pixel
    int32_t* d_seg_nodes = nullptr;            // ... and the open components on tile borders
    uint8_t* d_D2_narrow = nullptr;            // D2 after the L/R check narrowed for the host-output path (NarrowD2), narrow_stride bytes per frame
    // pinned host
    uint8_t* h_img[2] = {nullptr, nullptr};    // staging for images that arrive in pageable host memory
    uint8_t* h_D2_narrow = nullptr;            // landing buffer; widened to float into the caller's D2 by the worker
    FrameHeader* h_hdr = nullptr;
    int16_t* h_dcan = nullptr;                 // host-stage path only
    int32_t* h_tables = nullptr;               // host-stage path only: [support | tri1 | tri2 | units] of one frame
    HostStage host;                            // host-stage path only
    // the call in flight
    int n = 0;                                 // frames in this launch chain
    int first_frame = -1;                      // index of frame 0 in the batch
    std::vector<FrameIO> io;
    std::vector<float*> expand_D2;             // caller's D2 awaiting widening at finish (null: nothing to do)
    int last_n = 0;                            // frames of the last completed chain (elas_b200_time_matching)
    int d2_mode = 0;                           // how the chain in flight narrowed D2 for the copy out (NarrowD2::mode)
    float* d_view = nullptr;                   // colour map / back-projection outputs (5 planes), allocated on first use
    float* d_fuse = nullptr;                   // map fusion: previous + current map (10 planes), two point lists, work area
    float* last_D1 = nullptr;                  // frame 0's final left map of the last chain, if it lives in the group's buffers
    int map_tag = 0;                           // frame tag of the triangle-id map entries (k_grid_raster.cu)
    int scratch_phase = 0;                     // which of the two grid scatter buffers this chain uses
    int scratch_dirty_lo[2] = {0, 0}, scratch_dirty_hi[2] = {0, 0};   // frames [lo, hi) of a scatter buffer may hold marks of an earlier chain
    bool tables_valid = false;
    // introspection (single-frame calls)
    bool capture = false;
    std::map<std::string, std::vector<uint8_t>> stages;
    StageTimer timer;
};

}  // namespace

struct elas_b200_ctx {
    int device = 0;
    elas_b200_params p{};
    FrameGeom g{};
    GroupStrides st{};
    int support_cap = 0, tri_cap = 0, unit_cap = 0;
    int map_tag_shift = 0, map_tag_max = 0;  // map entry = tag << shift | triangle index
    bool mesh_device = true;                 // lattice filters + Delaunay on the GPU (k_mesh.cu); false: host stage (host_stage.cc)
    int32_t* d_prior = nullptr;
    std::vector<int32_t> prior_host;
    void* d_flush = nullptr;                 // > L2-sized buffer for elas_b200_time_matching
    size_t flush_bytes = 0;
    bool timing = false;
    int narrow_d2 = 2;                       // host output: D2 crosses PCIe narrowed when that is exact: 2 = u8 + validity bits where disp_max <= 255, else int16; 1 = int16; 0 = float32 (ELAS_B200_NARROW_D2)
    long long launches_at_create = 0;
    // host-side wall time, summed over all groups (nanoseconds)
    std::atomic<long long> ns_submit{0}, ns_host{0}, ns_wait{0}, ns_finish{0}, frames{0};
    std::vector<std::unique_ptr<Group>> groups;

    // batch calls: a pool of workers, worker w drives groups w, w + W, w + 2W, ...
    struct Job {
        int n = 0;
        const uint8_t* const* I1 = nullptr; const uint8_t* const* I2 = nullptr;
        float* const* D1 = nullptr; float* const* D2 = nullptr;
        int bpl = 0; bool device_io = false; int32_t* status = nullptr;
        std::atomic<int> next{0}, done{0}, worst{0};
        int active = 0;                      // workers currently holding this job (guarded by mu)
    };
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    Job* job = nullptr;
    uint64_t job_seq = 0;
    bool stopping = false;
    std::mutex batch_mu;                     // one call at a time
};

namespace {

inline long long now_ns()
{
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

FrameGeom make_geom(const elas_b200_params& p, int W, int H)
{
    FrameGeom g{};
    g.W = W; g.H = H;
    g.bpl = W + 15 - (W - 1) % 16;                                         // elas.cpp:37
    g.Dw = p.subsampling ? W / 2 : W;
    g.Dh = p.subsampling ? H / 2 : H;
    g.step = p.candidate_stepsize + (p.subsampling ? p.candidate_stepsize % 2 : 0);   // :453-457
    g.Wc = (W + g.step - 1) / g.step;                                      // :462-463
    g.Hc = (H + g.step - 1) / g.step;
    g.gw = (int)std::ceil((float)W / (float)p.grid_size);                  // :98-99
    g.gh = (int)std::ceil((float)H / (float)p.grid_size);
    g.dn = p.disp_max + 1;
    g.gwords = (g.dn + 31) / 32;
    g.plane_radius = (int)std::fmax((float)std::ceil(p.sigma * p.sradius), 2.0f);     // :993
    return g;
}

// prior table, elas.cpp:984-992 (float exp/log, as the reference's C++ overloads resolve them)
std::vector<int32_t> make_prior(const elas_b200_params& p, int dn)
{
    std::vector<int32_t> P(dn);
    const float two_sigma_squared = 2 * p.sigma * p.sigma;
    for (int dd = 0; dd < dn; dd++) {
        const float tmp = -std::log(p.gamma + std::exp(-dd * dd / two_sigma_squared)) + std::log(p.gamma);
        P[dd] = (int32_t)(tmp / p.beta);
    }
    return P;
}

// bytes between the narrowed right maps of consecutive frames: room for int16 per pixel (the u8 + validity-bit layout
// is smaller), a multiple of 16 so that every frame's words stay aligned
inline size_t narrow_stride(const GroupStrides& st) { return (2 * st.D + 15) & ~(size_t)15; }

void free_group(Group& s)
{
    for (int k = 0; k < 2; k++) {
        cudaFree(s.d_img[k]); cudaFreeHost(s.h_img[k]); cudaFree(s.d_desc[k]); cudaFree(s.d_tri[k]); cudaFree(s.d_units[k]);
        cudaFree(s.d_traster[k]); cudaFree(s.d_grid[k]); cudaFree(s.d_lists[k]);
        cudaFree(s.d_map[k]); cudaFree(s.d_raw[k]); cudaFree(s.d_D[k]); cudaFree(s.d_planes[k]);
    }
    cudaFree(s.d_dcan_raw); cudaFree(s.d_dcan); cudaFree(s.d_dcan_incon); cudaFree(s.d_support); cudaFree(s.d_mesh_scratch); cudaFree(s.d_lat_work);
    cudaFree(s.d_hdr); cudaFree(s.d_view); cudaFree(s.d_fuse); cudaFree(s.d_D2_narrow); cudaFreeHost(s.h_D2_narrow); cudaFreeHost(s.h_hdr);
    cudaFree(s.d_grid_scratch); cudaFree(s.d_tmp); cudaFree(s.d_seg_label); cudaFree(s.d_seg_nodes);
    cudaFreeHost(s.h_dcan); cudaFreeHost(s.h_tables);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    if (s.ev_out) cudaEventDestroy(s.ev_out);
    if (s.copy_stream) cudaStreamDestroy(s.copy_stream);
    for (auto& m : s.timer.marks) cudaEventDestroy(m.second);
    if (s.timer.begin) cudaEventDestroy(s.timer.begin);
    if (s.stream) cudaStreamDestroy(s.stream);
}

// ints of one frame's host-path tables: [support | tri1 | tri2 | units]
size_t table_ints(const elas_b200_ctx* c) { return 3 * ((size_t)c->support_cap + 2 * (size_t)c->tri_cap) + 2 * (size_t)c->unit_cap + 8; }

// Tensor map of a group's image buffer for k_descriptor: dims (column, row, frame), box = one image tile of one frame.
// cuTensorMapEncodeTiled is a driver entry point; it is fetched through the runtime so that the library keeps
// linking against cudart only.
int32_t encode_image_tensor_map(CUtensorMap* tm, uint8_t* base, int bpl, int H, int frames, size_t frame_stride)
{
    using Encode = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Encode encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q{};
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) return ELAS_B200_E_CUDA;
        encode = reinterpret_cast<Encode>(fn);
    }
    int box[2];
    descriptor_tile_box(box);
    const cuuint64_t dims[3] = {(cuuint64_t)bpl, (cuuint64_t)H, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)bpl, (cuuint64_t)frame_stride};          // bytes; bpl is a multiple of 16
    const cuuint32_t boxdim[3] = {(cuuint32_t)box[0], (cuuint32_t)box[1], 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, boxdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        std::fprintf(stderr, "elas_b200: cuTensorMapEncodeTiled failed with %d (bpl %d, H %d, frames %d)\n", (int)r, bpl, H, frames);
        return ELAS_B200_E_CUDA;
    }
    return ELAS_B200_OK;
}

int32_t alloc_group(elas_b200_ctx* c, Group& s, int cap)
{
    const FrameGeom& g = c->g;
    const GroupStrides& st = c->st;
    const size_t n = (size_t)cap;
    s.cap = cap;
    s.io.resize(cap); s.expand_D2.assign(cap, nullptr);
    CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    // a worker waits for a group's maps in cudaEventSynchronize: blocking, so that it leaves its core to the others
    CK(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming | cudaEventBlockingSync));
    for (int k = 0; k < 2; k++) {
        CK(cudaMalloc(&s.d_img[k], n * st.img));
        CK(cudaMemset(s.d_img[k], 0, n * st.img));                          // padding columns stay 0 (elas.cpp:42-43)
        if (int32_t rc = encode_image_tensor_map(&s.tm_img[k], s.d_img[k], c->g.bpl, c->g.H, cap, st.img)) return rc;
        CK(cudaMallocHost(&s.h_img[k], n * st.img));
        std::memset(s.h_img[k], 0, n * st.img);
        CK(cudaMalloc(&s.d_desc[k], n * st.desc * 16));
        CK(cudaMalloc(&s.d_tri[k], n * st.tri * 4));
        CK(cudaMalloc(&s.d_units[k], n * st.units * 4));
        CK(cudaMalloc(&s.d_traster[k], n * st.traster * sizeof(TriRaster)));
        CK(cudaMalloc(&s.d_planes[k], n * st.planes * 4));
        CK(cudaMalloc(&s.d_grid[k], n * st.grid * 4));
        CK(cudaMalloc(&s.d_lists[k], n * st.lists * 2));
        CK(cudaMalloc(&s.d_map[k], n * st.map * 4));
        CK(cudaMemset(s.d_map[k], 0xFF, n * st.map * 4));                   // -1 = not covered by any triangle
        CK(cudaMalloc(&s.d_raw[k], n * st.D * 4));
        CK(cudaMalloc(&s.d_D[k], n * st.D * 4));
    }
    CK(cudaMalloc(&s.d_dcan_raw, n * st.dcan * 2));
    CK(cudaMalloc(&s.d_dcan, n * st.dcan * 2));
    CK(cudaMalloc(&s.d_support, n * st.support * 4));
    CK(cudaMalloc(&s.d_hdr, n * sizeof(FrameHeader)));
    CK(cudaMemset(s.d_hdr, 0, n * sizeof(FrameHeader)));
    CK(cudaMallocHost(&s.h_hdr, n * sizeof(FrameHeader)));
    if (c->mesh_device) {
        CK(cudaMalloc(&s.d_mesh_scratch, n * 2 * st.mesh_scratch * 4));
        CK(cudaMalloc(&s.d_lat_work, n * st.lat_work * 4));
    }
    else {
        CK(cudaMallocHost(&s.h_dcan, n * st.dcan * 2));
        CK(cudaMallocHost(&s.h_tables, table_ints(c) * 4));
    }
    CK(cudaMalloc(&s.d_grid_scratch, n * st.scratch * 4));                  // per frame two buffers of [2][cells] words
    CK(cudaMemset(s.d_grid_scratch, 0, n * st.scratch * 4));
    CK(cudaMalloc(&s.d_tmp, n * 2 * st.D * 4));
    CK(cudaMalloc(&s.d_seg_label, n * st.D * 4));
    CK(cudaMalloc(&s.d_seg_nodes, n * st.seg_nodes * 4));
    CK(cudaMalloc(&s.d_D2_narrow, n * narrow_stride(st)));
    CK(cudaMallocHost(&s.h_D2_narrow, n * narrow_stride(st)));
    return ELAS_B200_OK;
}

// ---- introspection helpers -------------------------------------------------------------------

int32_t grab(Group& s, const char* name, const void* dptr, size_t bytes)
{
    std::vector<uint8_t>& v = s.stages[name];
    v.resize(bytes);
    CK(cudaMemcpyAsync(v.data(), dptr, bytes, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    return ELAS_B200_OK;
}

void grab_host(Group& s, const char* name, const void* ptr, size_t bytes)
{
    std::vector<uint8_t>& v = s.stages[name];
    v.assign((const uint8_t*)ptr, (const uint8_t*)ptr + bytes);
}

void mark(elas_b200_ctx* c, Group& s, const char* name)
{
    if (!c->timing) return;
    cudaEvent_t e = nullptr;
    for (auto& m : s.timer.marks) if (m.first == name) e = m.second;
    if (!e) { cudaEventCreate(&e); s.timer.marks.emplace_back(name, e); }
    cudaEventRecord(e, s.stream);
}

// expands a per-cell bitmask grid to the reference's int32 [gh][gw][disp_max+2] lists (elas.cpp:754-775)
std::vector<uint8_t> expand_grid(const FrameGeom& g, const std::vector<uint8_t>& raw)
{
    const uint32_t* bits = reinterpret_cast<const uint32_t*>(raw.data());
    const size_t cells = (size_t)g.gw * g.gh;
    std::vector<uint8_t> out(cells * (g.dn + 1) * 4, 0);
    int32_t* o = reinterpret_cast<int32_t*>(out.data());
    for (size_t cidx = 0; cidx < cells; cidx++) {
        int k = 1;
        for (int d = 0; d < g.dn; d++)
            if (bits[cidx * g.gwords + (d >> 5)] >> (d & 31) & 1u) o[cidx * (g.dn + 1) + k++] = d;
        o[cidx * (g.dn + 1)] = k - 1;
    }
    return out;
}

// what kind of memory a caller's pointer is (launch errors are checked where the launches are issued, so
// clearing the error of a failed query here cannot swallow one)
struct PtrKind { bool pageable, device, mapped_host, device_alias_only; void* device_alias; };
PtrKind classify(const void* ptr, int device)
{
    PtrKind k{false, false, false, false, nullptr};
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) { cudaGetLastError(); k.pageable = true; return k; }
    k.pageable = attr.type == cudaMemoryTypeUnregistered;
    // only memory of THIS device is written in place by the kernels; anything else takes the copy path
    k.device = (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) && attr.device == device;
    k.mapped_host = attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr;
    // device or managed memory that is NOT this device's: reachable by copies only, and not CPU-addressable for sure
    k.device_alias_only = (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) && !k.device;
    k.device_alias = attr.devicePointer;
    return k;
}

// elas.cpp:35-56: W bytes of every row go into the 16-byte-aligned zero-padded copy; when the caller's
// pitch already equals that padded pitch the reference memcpy's the whole block (:44-48) -- one 1-D copy
int32_t copy_image_in(const elas_b200_ctx* c, uint8_t* dst, const uint8_t* src, int pitch, cudaStream_t st, uint8_t* staging)
{
    const FrameGeom& g = c->g;
    if (staging && classify(src, c->device).pageable) {
        // Pageable host memory (stereomapper's IplImage buffers are malloc'ed): the driver would stage such a
        // copy itself at a few GB/s; rows go into the group's pinned staging image instead (padding columns
        // stay 0 unless the caller's pitch is the padded pitch, elas.cpp:44-55) and leave as one pinned copy
        if (pitch == g.bpl) std::memcpy(staging, src, (size_t)g.bpl * g.H);
        else for (int v = 0; v < g.H; v++) std::memcpy(staging + (size_t)v * g.bpl, src + (size_t)v * pitch, (size_t)g.W);
        CK(cudaMemcpyAsync(dst, staging, (size_t)g.bpl * g.H, cudaMemcpyHostToDevice, st));
        return ELAS_B200_OK;
    }
    if (pitch == g.bpl) CK(cudaMemcpyAsync(dst, src, (size_t)g.bpl * g.H, cudaMemcpyDefault, st));
    else CK(cudaMemcpy2DAsync(dst, g.bpl, src, pitch, g.W, g.H, cudaMemcpyDefault, st));
    return ELAS_B200_OK;
}

MatchBuffers match_buffers(const elas_b200_ctx* c, const Group& s)
{
    MatchBuffers b{};
    for (int k = 0; k < 2; k++) {
        b.desc[k] = s.d_desc[k]; b.tri[k] = s.d_traster[k]; b.map[k] = s.d_map[k]; b.grid[k] = s.d_grid[k];
        b.lists[k] = s.d_lists[k]; b.D[k] = s.d_raw[k];
    }
    b.prior = c->d_prior;
    b.prior_host = c->prior_host.data();
    b.desc_stride = c->st.desc; b.tri_stride = c->st.traster; b.map_stride = c->st.map; b.grid_stride = c->st.grid;
    b.lists_stride = c->st.lists; b.D_stride = c->st.D;
    return b;
}

// ---- host-stage path (parameters the device mesh stage does not take): lattice to the host, filters +
// Delaunay on the CPU (host_stage.cc), tables and header back.  One frame after the other.
int32_t mesh_on_host(elas_b200_ctx* c, Group& s)
{
    const FrameGeom& g = c->g;
    const elas_b200_params& p = c->p;
    const size_t lat = (size_t)g.Wc * g.Hc;
    CK(cudaMemcpyAsync(s.h_dcan, s.d_dcan_raw, (size_t)s.n * c->st.dcan * 2, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    const long long t0 = now_ns();
    for (int f = 0; f < s.n; f++) {
        int16_t* dcan = s.h_dcan + (size_t)f * c->st.dcan;
        if (s.capture) grab_host(s, "dcan_raw", dcan, lat * 2);
        const int n = s.host.run(g, p, dcan, s.capture, false);
        if (s.capture) grab_host(s, "dcan_incon", s.host.dcan_incon.data(), s.host.dcan_incon.size() * 2);
        FrameHeader h{};
        h.n_support = n;
        if (n >= 3) {
            const int nt1 = (int)s.host.tri[0].size() / 3, nt2 = (int)s.host.tri[1].size() / 3;
            if (n > c->support_cap || nt1 > c->tri_cap || nt2 > c->tri_cap) return ELAS_B200_E_UNSUPPORTED;
            h.n_tri[0] = nt1; h.n_tri[1] = nt2;
            // the host lists the units of both images in one array (the image is a bit of the unit): they go into
            // the left image's list; if there are too many, every triangle is scan-converted by a warp of its own
            const int n_units = (int)s.host.units.size() / 2;
            const bool fits = n_units <= 2 * c->unit_cap;
            h.n_units[0] = fits ? std::min(n_units, c->unit_cap) : 0;
            h.n_units[1] = fits ? n_units - h.n_units[0] : 0;
            h.ovf_from[0] = fits ? nt1 : 0; h.ovf_from[1] = fits ? nt2 : 0;
            int32_t* t = s.h_tables;
            std::memcpy(t, s.host.support.data(), (size_t)n * 12);
            std::memcpy(t + 3 * (size_t)c->support_cap, s.host.tri[0].data(), (size_t)nt1 * 12);
            std::memcpy(t + 3 * ((size_t)c->support_cap + c->tri_cap), s.host.tri[1].data(), (size_t)nt2 * 12);
            CK(cudaMemcpyAsync(s.d_support + (size_t)f * c->st.support, t, (size_t)n * 12, cudaMemcpyHostToDevice, s.stream));
            CK(cudaMemcpyAsync(s.d_tri[0] + (size_t)f * c->st.tri, t + 3 * (size_t)c->support_cap, (size_t)nt1 * 12, cudaMemcpyHostToDevice, s.stream));
            CK(cudaMemcpyAsync(s.d_tri[1] + (size_t)f * c->st.tri, t + 3 * ((size_t)c->support_cap + c->tri_cap), (size_t)nt2 * 12, cudaMemcpyHostToDevice, s.stream));
            if (fits) {
                int32_t* u = t + 3 * ((size_t)c->support_cap + 2 * (size_t)c->tri_cap);
                std::memcpy(u, s.host.units.data(), (size_t)n_units * 8);
                CK(cudaMemcpyAsync(s.d_units[0] + (size_t)f * c->st.units, u, (size_t)h.n_units[0] * 8, cudaMemcpyHostToDevice, s.stream));
                if (h.n_units[1]) CK(cudaMemcpyAsync(s.d_units[1] + (size_t)f * c->st.units, u + 2 * (size_t)h.n_units[0], (size_t)h.n_units[1] * 8, cudaMemcpyHostToDevice, s.stream));
            }
        }
        s.h_hdr[f] = h;
        CK(cudaMemcpyAsync(s.d_dcan + (size_t)f * c->st.dcan, dcan, lat * 2, cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(s.d_hdr + f, s.h_hdr + f, sizeof(FrameHeader), cudaMemcpyHostToDevice, s.stream));
        // h_tables is reused by the next frame of the group
        if (f + 1 < s.n || true) CK(cudaStreamSynchronize(s.stream));
    }
    c->ns_host += now_ns() - t0;
    return ELAS_B200_OK;
}

// ---- one launch chain: s.n frames (s.io) from images to maps, everything enqueued on the group's streams -----------
int32_t submit_group(elas_b200_ctx* c, Group& s)
{
    const FrameGeom& g = c->g;
    const elas_b200_params& p = c->p;
    const GroupStrides& gs = c->st;
    const int n = s.n;
    const size_t N = (size_t)g.W * g.H, ND = (size_t)g.Dw * g.Dh;
    cudaStream_t st = s.stream;
    const long long t0 = now_ns();
    if (s.capture) s.stages.clear();
    s.tables_valid = false;
    if (c->timing) {
        if (!s.timer.begin) cudaEventCreate(&s.timer.begin);
        cudaEventRecord(s.timer.begin, st);
    }
    // ---- images in, descriptors, support search ----------------------------------------------------------------------
    for (int f = 0; f < n; f++) {
        const FrameIO& io = s.io[f];
        if (int32_t rc = copy_image_in(c, s.d_img[0] + (size_t)f * gs.img, io.I1, io.bytes_per_line, st, io.device_io ? nullptr : s.h_img[0] + (size_t)f * gs.img)) return rc;
        if (int32_t rc = copy_image_in(c, s.d_img[1] + (size_t)f * gs.img, io.I2, io.bytes_per_line, st, io.device_io ? nullptr : s.h_img[1] + (size_t)f * gs.img)) return rc;
    }
    mark(c, s, "copy_in");
    launch_descriptor(g, p.subsampling, s.tm_img[0], s.tm_img[1], s.d_desc[0], s.d_desc[1], gs, n, st);
    mark(c, s, "descriptor");
    launch_support(g, p, s.d_desc[0], s.d_desc[1], s.d_dcan_raw, gs, n, st);
    mark(c, s, "support");
    CK(cudaGetLastError());                      // launch-configuration errors are not sticky: report them with THIS group
    if (s.capture) {
        if (int32_t rc = grab(s, "desc1", s.d_desc[0], N * 16)) return rc;
        if (int32_t rc = grab(s, "desc2", s.d_desc[1], N * 16)) return rc;
        if (int32_t rc = grab(s, "dcan_raw", s.d_dcan_raw, (size_t)g.Wc * g.Hc * 2)) return rc;
        int32_t lat[2] = {g.Wc, g.Hc};
        grab_host(s, "lattice_dims", lat, sizeof lat);
    }
    // ---- mesh stage: lattice filters, support list, Delaunay x2, raster units ---------------------------------------------
    if (c->mesh_device) {
        if (s.capture && !s.d_dcan_incon) CK(cudaMalloc(&s.d_dcan_incon, (size_t)s.cap * gs.dcan * 2));
        launch_lattice(g, p, s.d_dcan_raw, s.d_dcan, s.capture ? s.d_dcan_incon : nullptr, s.d_support, s.d_lat_work, s.d_hdr, gs, n, st);
        mark(c, s, "lattice");
        launch_delaunay(g, s.d_support, s.d_tri[0], s.d_tri[1], s.d_units[0], s.d_units[1], c->unit_cap, s.d_hdr,
                        s.d_mesh_scratch, gs, n, st);
        mark(c, s, "delaunay");
        CK(cudaGetLastError());
        if (s.capture) if (int32_t rc = grab(s, "dcan_incon", s.d_dcan_incon, (size_t)g.Wc * g.Hc * 2)) return rc;
    } else {
        if (int32_t rc = mesh_on_host(c, s)) return rc;
        mark(c, s, "host_stage");
    }
    if (s.capture) {
        FrameHeader h{};
        CK(cudaMemcpyAsync(&h, s.d_hdr, sizeof h, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (int32_t rc = grab(s, "dcan", s.d_dcan, (size_t)g.Wc * g.Hc * 2)) return rc;
        if (int32_t rc = grab(s, "support", s.d_support, (size_t)std::min(h.n_support, c->support_cap) * 12)) return rc;
        if (int32_t rc = grab(s, "tri1", s.d_tri[0], (size_t)h.n_tri[0] * 12)) return rc;
        if (int32_t rc = grab(s, "tri2", s.d_tri[1], (size_t)h.n_tri[1] * 12)) return rc;
        int32_t gd[3] = {p.disp_max + 2, g.gw, g.gh};
        grab_host(s, "grid_dims", gd, sizeof gd);
        grab_host(s, "header", &h, sizeof h);
    }
    // ---- planes, candidate grid, triangle-id maps, dense matching ------------------------------------------------------------
    const size_t scratch_words = 2 * (size_t)g.gw * g.gh * g.gwords;
    const int ph = s.scratch_phase;
    uint32_t* scratch_cur = s.d_grid_scratch + (size_t)ph * scratch_words;
    uint32_t* scratch_next = s.d_grid_scratch + (size_t)(1 - ph) * scratch_words;
    s.scratch_phase ^= 1;
    // A chain zeroes the OTHER buffer for its own n frames only.  After a short chain (the tail of a batch) the
    // frames beyond it still hold the marks of the chain before: clear them before a longer chain reads them.
    if (s.scratch_dirty_lo[ph] < std::min(s.scratch_dirty_hi[ph], n)) {
        const int lo = s.scratch_dirty_lo[ph], hi = s.scratch_dirty_hi[ph];
        CK(cudaMemset2DAsync(scratch_cur + (size_t)lo * gs.scratch, gs.scratch * 4, 0, scratch_words * 4, (size_t)(hi - lo), st));
        s.scratch_dirty_lo[ph] = s.scratch_dirty_hi[ph] = 0;
    }
    s.scratch_dirty_hi[ph] = std::max(s.scratch_dirty_hi[ph], n);            // this chain scatters into frames [0, n) ...
    s.scratch_dirty_lo[ph] = 0;
    if (s.scratch_dirty_hi[1 - ph] <= n) s.scratch_dirty_lo[1 - ph] = s.scratch_dirty_hi[1 - ph] = 0;      // ... and zeroes them in the other buffer
    else s.scratch_dirty_lo[1 - ph] = std::max(s.scratch_dirty_lo[1 - ph], n);
    if (++s.map_tag > c->map_tag_max) {
        // tag space used up: start over from cleared maps
        for (int k = 0; k < 2; k++) CK(cudaMemsetAsync(s.d_map[k], 0xFF, (size_t)s.cap * gs.map * 4, st));
        s.map_tag = 1;
    }
    const int tag_bits = s.map_tag << c->map_tag_shift;
    launch_planes_scatter(g, p, s.d_hdr, s.d_support, s.d_tri[0], s.d_tri[1], s.d_traster[0], s.d_traster[1], s.d_planes[0],
                          s.d_planes[1], scratch_cur, gs, n, st);                          // elas.cpp:87-88, :697-727
    mark(c, s, "planes+scatter");
    if (s.capture) {
        const FrameHeader& h = *reinterpret_cast<const FrameHeader*>(s.stages["header"].data());
        if (int32_t rc = grab(s, "planes1", s.d_planes[0], (size_t)h.n_tri[0] * 24)) return rc;
        if (int32_t rc = grab(s, "planes2", s.d_planes[1], (size_t)h.n_tri[1] * 24)) return rc;
    }
    launch_diffuse_raster(g, p.subsampling, s.d_hdr, scratch_cur, scratch_next, s.d_grid[0], s.d_grid[1], s.d_lists[0],
                          s.d_lists[1], s.d_traster[0], s.d_traster[1], s.d_units[0], s.d_units[1], s.d_map[0],
                          s.d_map[1], tag_bits, gs, n, st);                                 // :732-775, :1074-1114
    mark(c, s, "diffuse+raster");
    launch_matching(g, p, match_buffers(c, s), n, tag_bits, c->map_tag_shift, st);
    mark(c, s, "matching");
    CK(cudaGetLastError());
    s.tables_valid = true;
    if (s.capture) {
        const size_t cells = (size_t)g.gw * g.gh * g.gwords;
        if (int32_t rc = grab(s, "grid1_bits", s.d_grid[0], cells * 4)) return rc;
        if (int32_t rc = grab(s, "grid2_bits", s.d_grid[1], cells * 4)) return rc;
        if (int32_t rc = grab(s, "D1_raw", s.d_raw[0], ND * 4)) return rc;
        if (int32_t rc = grab(s, "D2_raw", s.d_raw[1], ND * 4)) return rc;
    }
    // ---- L/R check and post-processing ------------------------------------------------------------------------------------------
    const int n_post = p.postprocess_only_left ? 1 : 2;                                  // elas.cpp:121-159
    const bool fused_post = post_fusable(p) && !p.filter_median;
    const bool rows_fused = lr_rows_fusable(g);
    // Where a finished map goes: a caller's buffer in THIS device's memory is written by the last kernel that
    // touches the map (D2 without post-processing is final after the L/R check); host buffers are reached by
    // copies from the group's buffers.
    bool direct[2][kMaxGroupFrames] = {};
    bool any_direct_d2 = false, all_host = true;
    for (int f = 0; f < n; f++)
        for (int k = 0; k < 2; k++) {
            float* user = k ? s.io[f].D2 : s.io[f].D1;
            if (!user) continue;                     // batch calls may leave D2 out (stereomapper never reads it)
            // device memory of any GPU (this one is written in place by the fused tail, others and the unfused chain
            // are reached by copies): never handed to the CPU-side widening
            const bool on_device = s.io[f].device_io || classify(user, c->device).device || classify(user, c->device).device_alias_only;
            direct[k][f] = fused_post && (s.io[f].device_io || classify(user, c->device).device);
            if (k == 1 && direct[k][f]) any_direct_d2 = true;
            if (on_device) all_host = false;
        }
    // Copy path, D2 final after the L/R check: its values are raw integer disparities or -10, so it crosses
    // PCIe as int16 (half the bytes) and the worker widens it into the caller's float map.
    // ... or, when every disparity fits a byte, as u8 plus one validity bit per pixel (1.125 bytes per pixel)
    const bool d2_narrow = all_host && !any_direct_d2 && n_post == 1 && rows_fused && !s.capture && !c->timing && c->narrow_d2;
    const int d2_mode = !d2_narrow ? 0 : (p.disp_max <= 255 && c->narrow_d2 > 1) ? 2 : 1;
    const NarrowD2 narrow = narrow_d2_layout(g, d2_mode, s.d_D2_narrow, narrow_stride(gs));
    s.d2_mode = d2_mode;
    OutTable lr_d2 = out_table(s.d_D[1], gs.D, n);
    if (n_post == 1) for (int f = 0; f < n; f++) if (direct[1][f]) lr_d2.p[f] = s.io[f].D2;
    if (rows_fused) launch_lr_rows(g, p, s.d_raw[0], s.d_raw[1], s.d_D[0], lr_d2, narrow, gs.D, n, st);
    else launch_lr_check(g, p, s.d_raw[0], s.d_raw[1], s.d_D[0], lr_d2, gs.D, n, st);       // elas.cpp:116
    mark(c, s, "lr_check");
    if (s.capture) {
        if (int32_t rc = grab(s, "D1_lr", s.d_D[0], ND * 4)) return rc;
        if (int32_t rc = grab(s, "D2_lr", lr_d2.p[0], ND * 4)) return rc;
    }
    // final_map[k] = where frame f's finished map k lives
    OutTable final_map[2] = {out_table(s.d_D[0], gs.D, n), lr_d2};
    if (fused_post) {
        // speckle sizes (K9 rows/merge/count), then ONE kernel for speckle apply + gap interpolation +
        // adaptive mean; it reads d_D and writes the final map (d_raw is dead after the L/R check)
        for (int k = 0; k < n_post; k++) {
            launch_segments(g, p, s.d_D[k], s.d_seg_label, s.d_seg_nodes, gs.D, gs.seg_nodes, n, st, false);
            mark(c, s, k ? "segments2" : "segments");
            final_map[k] = out_table(s.d_raw[k], gs.D, n);
            for (int f = 0; f < n; f++) if (direct[k][f]) final_map[k].p[f] = k ? s.io[f].D2 : s.io[f].D1;
            launch_post_fused(g, p, s.d_D[k], s.d_seg_label, s.d_seg_nodes, gs.seg_nodes, final_map[k],
                              s.capture ? s.d_tmp : nullptr, s.capture ? s.d_tmp + ND : nullptr, gs.D, n, st);
            if (s.capture) {
                if (int32_t rc = grab(s, k ? "D2_seg" : "D1_seg", s.d_tmp, ND * 4)) return rc;
                if (int32_t rc = grab(s, k ? "D2_gap" : "D1_gap", p.filter_adaptive_mean ? s.d_tmp + ND : final_map[k].p[0], ND * 4)) return rc;
            }
        }
        mark(c, s, "apply+gap+mean");
        if (s.capture && n_post == 1) {
            if (int32_t rc = grab(s, "D2_seg", lr_d2.p[0], ND * 4)) return rc;
            if (int32_t rc = grab(s, "D2_gap", lr_d2.p[0], ND * 4)) return rc;
        }
    } else {
        // the unfused chain works in place on the group's buffers (settings outside the ROBOTICS family)
        for (int k = 0; k < n_post; k++) launch_segments(g, p, s.d_D[k], s.d_seg_label, s.d_seg_nodes, gs.D, gs.seg_nodes, n, st, true);
        mark(c, s, "segments");
        if (s.capture) {
            if (int32_t rc = grab(s, "D1_seg", s.d_D[0], ND * 4)) return rc;
            if (int32_t rc = grab(s, "D2_seg", s.d_D[1], ND * 4)) return rc;
        }
        for (int k = 0; k < n_post; k++) launch_gap(g, p, s.d_D[k], s.d_tmp, gs.D, 2 * gs.D, n, st);
        mark(c, s, "gap");
        if (s.capture) {
            if (int32_t rc = grab(s, "D1_gap", s.d_D[0], ND * 4)) return rc;
            if (int32_t rc = grab(s, "D2_gap", s.d_D[1], ND * 4)) return rc;
        }
        if (p.filter_adaptive_mean) {
            for (int k = 0; k < n_post; k++) launch_adaptive_mean(g, p, s.d_D[k], s.d_tmp, gs.D, 2 * gs.D, n, st);
            mark(c, s, "adaptive_mean");
        }
        if (s.capture) {
            if (int32_t rc = grab(s, "D1_mean", s.d_D[0], ND * 4)) return rc;
            if (int32_t rc = grab(s, "D2_mean", s.d_D[1], ND * 4)) return rc;
        }
        if (p.filter_median) {
            for (int k = 0; k < n_post; k++) launch_median(g, s.d_D[k], s.d_tmp, gs.D, 2 * gs.D, n, st);
            mark(c, s, "median");
        }
    }
    if (s.capture) {
        if (fused_post) {
            if (int32_t rc = grab(s, "D1_mean", final_map[0].p[0], ND * 4)) return rc;
            if (int32_t rc = grab(s, "D2_mean", final_map[1].p[0], ND * 4)) return rc;
        }
        if (int32_t rc = grab(s, "D1", final_map[0].p[0], ND * 4)) return rc;
        if (int32_t rc = grab(s, "D2", final_map[1].p[0], ND * 4)) return rc;
    }
    CK(cudaGetLastError());
    // a group-owned copy of frame 0's final left map serves elas_b200_colormap / _reproject with D1 == NULL
    s.last_D1 = direct[0][0] ? nullptr : final_map[0].p[0];
    // ---- maps and headers out: nothing to do for maps the kernels stored in place; the others are copied on the
    // group's copy stream (stage timing keeps them in line) ---------------------------------------------------------------------
    cudaStream_t out_stream = st;
    if (!c->timing) {
        CK(cudaEventRecord(s.ev_done, st));
        CK(cudaStreamWaitEvent(s.copy_stream, s.ev_done, 0));
        out_stream = s.copy_stream;
    }
    for (int f = 0; f < n; f++) {
        s.expand_D2[f] = nullptr;
        for (int k = 0; k < 2; k++) {
            float* user = k ? s.io[f].D2 : s.io[f].D1;
            if (direct[k][f] || !user) continue;
            if (k == 1 && d2_mode) {
                CK(cudaMemcpyAsync(s.h_D2_narrow + f * narrow_stride(gs), s.d_D2_narrow + f * narrow_stride(gs), narrow.bytes, cudaMemcpyDeviceToHost, out_stream));
                s.expand_D2[f] = user;
            } else {
                CK(cudaMemcpyAsync(user, final_map[k].p[f], ND * 4, cudaMemcpyDefault, out_stream));
            }
        }
    }
    CK(cudaMemcpyAsync(s.h_hdr, s.d_hdr, (size_t)n * sizeof(FrameHeader), cudaMemcpyDeviceToHost, out_stream));
    mark(c, s, "copy_out");
    CK(cudaEventRecord(s.ev_out, out_stream));
    c->ns_submit += now_ns() - t0;
    return ELAS_B200_OK;
}

// ---- the group's maps are in the caller's buffers: per-frame status into status_out[0..n) ------------------------------
int32_t finish_group(elas_b200_ctx* c, Group& s, int32_t* status_out)
{
    const long long t0 = now_ns();
    cudaError_t waited = cudaEventSynchronize(s.ev_out);
    if (waited == cudaSuccess) waited = cudaGetLastError();
    if (waited != cudaSuccess) {
        // the chain failed: nothing landed, nothing to widen (the callers' maps keep whatever they held)
        std::fill(s.expand_D2.begin(), s.expand_D2.end(), nullptr);
        CK(waited);
    }
    const long long t1 = now_ns();
    c->ns_wait += t1 - t0; c->frames += s.n;
    const size_t ND = (size_t)c->g.Dw * c->g.Dh;
    for (int f = 0; f < s.n; f++) {
        const FrameHeader& h = s.h_hdr[f];
        status_out[f] = h.status < 0 ? h.status : h.n_support < 3 ? ELAS_B200_E_FEW_SUPPORT : ELAS_B200_OK;
        if (s.expand_D2[f]) {
            const uint8_t* landed = s.h_D2_narrow + f * narrow_stride(c->st);
            if (s.d2_mode == 2) {
                const NarrowD2 lay = narrow_d2_layout(c->g, 2, nullptr, 0);
                widen_u8_mask_to_f32(landed, reinterpret_cast<const uint32_t*>(landed + lay.mask_offset), lay.mask_words_per_row,
                                     s.expand_D2[f], c->g.Dw, c->g.Dh);
            } else widen_i16_to_f32(reinterpret_cast<const int16_t*>(landed), s.expand_D2[f], ND);
            s.expand_D2[f] = nullptr;
        }
    }
    s.last_n = s.n;
    c->ns_finish += now_ns() - t1;
    if (c->timing) {
        CK(cudaStreamSynchronize(s.stream));
        s.timer.last.clear();
        cudaEvent_t prev = s.timer.begin;
        // marks are appended in first-use order, which is pipeline order
        for (auto& m : s.timer.marks) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, prev, m.second) == cudaSuccess) s.timer.last.emplace_back(m.first, ms);
            prev = m.second;
        }
    }
    return ELAS_B200_OK;
}

// one frame, start to finish (the drop-in call and elas_b200_process_ctx)
int32_t run_frame(elas_b200_ctx* c, Group& s, const uint8_t* I1, const uint8_t* I2, float* D1,
                  float* D2, int bytes_per_line, bool device_io)
{
    s.n = 1;
    s.io[0] = FrameIO{I1, I2, D1, D2, bytes_per_line, device_io};
    int32_t rc = submit_group(c, s);
    if (rc) { cudaStreamSynchronize(s.stream); cudaStreamSynchronize(s.copy_stream); return rc; }
    int32_t status = 0;
    rc = finish_group(c, s, &status);
    return rc ? rc : status;
}

// Batch scheduler.  Nothing of a frame runs on the CPU (device mesh stage), so a worker only enqueues launch
// chains and collects results: worker w drives the groups w, w + W, ...; for each of them in turn it waits
// (blocking, no spinning) for the chain in flight, reports its frames, claims the next `cap` frames of the
// batch and enqueues their chain.  With two or more groups per worker the GPU always has chains queued.
void worker_main(elas_b200_ctx* c, int worker, int n_workers)
{
    cudaSetDevice(c->device);
    uint64_t seen = 0;
    const int n_groups = (int)c->groups.size();
    std::vector<int32_t> status(kMaxGroupFrames);
    for (;;) {
        elas_b200_ctx::Job* job = nullptr;
        {
            std::unique_lock<std::mutex> lk(c->mu);
            c->cv_work.wait(lk, [&] { return c->stopping || (c->job && c->job_seq != seen); });
            if (c->stopping) return;
            job = c->job;
            seen = c->job_seq;
            job->active++;
        }
        auto report = [&](int i, int32_t rc) {
            if (job->status) job->status[i] = rc;
            if (rc < 0) { int w = job->worst.load(); while (rc < w && !job->worst.compare_exchange_weak(w, rc)) {} }
            job->done.fetch_add(1);
        };
        bool in_flight_any = true, frames_left = true;
        while (in_flight_any || frames_left) {
            in_flight_any = false;
            for (int gi = worker; gi < n_groups; gi += n_workers) {
                Group& s = *c->groups[gi];
                if (s.first_frame >= 0) {
                    // collect the chain in flight
                    const int32_t rc = finish_group(c, s, status.data());
                    for (int f = 0; f < s.n; f++) report(s.first_frame + f, rc ? rc : status[f]);
                    s.first_frame = -1;
                }
                if (!frames_left) continue;
                const int i0 = job->next.fetch_add(s.cap);
                if (i0 >= job->n) { frames_left = false; continue; }
                s.n = std::min(s.cap, job->n - i0);
                for (int f = 0; f < s.n; f++)
                    s.io[f] = FrameIO{job->I1[i0 + f], job->I2[i0 + f], job->D1[i0 + f], job->D2[i0 + f], job->bpl, job->device_io};
                const int32_t rc = submit_group(c, s);
                if (rc) {
                    cudaStreamSynchronize(s.stream); cudaStreamSynchronize(s.copy_stream);
                    for (int f = 0; f < s.n; f++) report(i0 + f, rc);
                } else {
                    s.first_frame = i0;
                    in_flight_any = true;
                }
            }
        }
        {
            std::lock_guard<std::mutex> lk(c->mu);
            job->active--;
            c->cv_done.notify_all();
        }
    }
}

int32_t run_batch(elas_b200_ctx* c, int32_t n, const uint8_t* const* I1, const uint8_t* const* I2,
                  float* const* D1, float* const* D2, int32_t bpl, int32_t* status, bool device_io)
{
    if (!c || n < 0 || !I1 || !I2 || !D1 || !D2 || bpl < c->g.W) return ELAS_B200_E_BAD_ARG;
    if (n == 0) return ELAS_B200_OK;
    std::lock_guard<std::mutex> batch(c->batch_mu);
    elas_b200_ctx::Job job;
    job.n = n; job.I1 = I1; job.I2 = I2; job.D1 = D1; job.D2 = D2; job.bpl = bpl;
    job.device_io = device_io; job.status = status;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        c->job = &job;
        c->job_seq++;
    }
    c->cv_work.notify_all();
    {
        std::unique_lock<std::mutex> lk(c->mu);
        c->cv_done.wait(lk, [&] { return job.done.load() >= n && job.active == 0; });
        c->job = nullptr;
    }
    return job.worst.load();
}

// cache of single-slot contexts behind the synchronous drop-in call
struct CacheKey {
    int device, W, H;
    elas_b200_params p;
    bool operator<(const CacheKey& o) const { return std::memcmp(this, &o, sizeof *this) < 0; }
};
// at most kCacheMax contexts stay cached (each pins a frame group of device memory and a worker thread); the least
// recently used one is destroyed when another (device, size, parameter block) combination shows up
constexpr size_t kCacheMax = 4;
struct CacheEntry { elas_b200_ctx* ctx; uint64_t last_use; };
std::mutex g_cache_mu;
std::map<CacheKey, CacheEntry> g_cache;
uint64_t g_cache_tick = 0;

}  // namespace

// Frame groups run their launch chains on separate streams, and a chain holds kernels that run for a long time on
// one or two CTAs (the mesh stage).  With the default of 8 hardware work queues, streams share queues and a chain that
// waits for such a kernel blocks the streams queued behind it (4096x2160: 530 -> 910 pairs/s with 32 queues; 1242x375
// end to end: 12.1 k -> 13.1 k).  The variable is read when the CUDA context is created: the library sets it when it is
// loaded unless the host application has chosen a value; a host that initialises CUDA before loading the library
// should export CUDA_DEVICE_MAX_CONNECTIONS=32 itself (bench.py does).
__attribute__((constructor)) static void elas_b200_more_work_queues() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

extern "C" {

void elas_b200_default_params(elas_b200_params* p, int32_t setting)
{
    // Elas::parameters(setting), elas.h:88-147
    const bool mb = setting == ELAS_B200_MIDDLEBURY;
    p->disp_min = 0;                 p->disp_max = 255;
    p->support_threshold = mb ? 0.95f : 0.85f;
    p->support_texture = 10;         p->candidate_stepsize = 5;
    p->incon_window_size = 5;        p->incon_threshold = 5;
    p->incon_min_support = 5;        p->add_corners = mb ? 1 : 0;
    p->grid_size = 20;               p->beta = 0.02f;
    p->gamma = mb ? 5.f : 3.f;       p->sigma = 1.f;
    p->sradius = mb ? 3.f : 2.f;     p->match_texture = mb ? 0 : 1;
    p->lr_threshold = 2;             p->speckle_sim_threshold = 1.f;
    p->speckle_size = 200;           p->ipol_gap_width = mb ? 5000 : 3;
    p->filter_median = mb ? 1 : 0;   p->filter_adaptive_mean = mb ? 0 : 1;
    p->postprocess_only_left = mb ? 0 : 1;
    p->subsampling = 0;
}

void elas_b200_stereomapper_params(elas_b200_params* p)
{
    elas_b200_default_params(p, ELAS_B200_ROBOTICS);     // stereothread.cpp:76-80
    p->postprocess_only_left = 1;
    p->filter_adaptive_mean = 1;
    p->support_texture = 30;
}

int32_t elas_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* elas_b200_version(void) { return "elas_b200 0.1 (sm_100a; CUDA kernels only, no CPU fallback)"; }

int32_t elas_b200_create(elas_b200_ctx** out, int32_t device, const elas_b200_params* p,
                         int32_t width, int32_t height, int32_t n_slots)
{
    return elas_b200_create_ex(out, device, p, width, height, n_slots, 0);
}

int32_t elas_b200_create_ex(elas_b200_ctx** out, int32_t device, const elas_b200_params* p,
                            int32_t width, int32_t height, int32_t n_slots, int32_t n_workers)
{
    int frames = 1;
    if (const char* e = std::getenv("ELAS_B200_FRAMES_PER_GROUP")) frames = std::atoi(e);
    return elas_b200_create_grouped(out, device, p, width, height, n_slots, frames, n_workers);
}

int32_t elas_b200_create_grouped(elas_b200_ctx** out, int32_t device, const elas_b200_params* p,
                                 int32_t width, int32_t height, int32_t n_groups, int32_t frames_per_group,
                                 int32_t n_workers)
{
    if (!out || !p || width < 16 || n_groups < 1 || n_groups > 64 || frames_per_group < 0) return ELAS_B200_E_BAD_ARG;
    *out = nullptr;
    if (p->disp_max < 0 || p->disp_max > 4095 || p->disp_min > p->disp_max || p->candidate_stepsize < 1) return ELAS_B200_E_BAD_ARG;
    // the matching kernel divides by grid_size with a 32-bit reciprocal (exact for u, grid_size < 65536)
    if (p->grid_size < 2 || p->grid_size > 4096 || width >= 65536 || height >= 65536) return ELAS_B200_E_UNSUPPORTED;
    // createGrid's diffusion walks from grid row 2 (elas.cpp:732-748): fewer than 3 grid rows is UB there
    if ((int)std::ceil((float)height / (float)p->grid_size) < 3) return ELAS_B200_E_UNSUPPORTED;
    if (elas_b200_device_count() <= device || device < 0) return ELAS_B200_E_NO_DEVICE;
    CK(cudaSetDevice(device));
    std::unique_ptr<elas_b200_ctx> c(new elas_b200_ctx);
    c->device = device; c->p = *p;
    c->g = make_geom(*p, width, height);
    const FrameGeom& g = c->g;
    // frames per launch chain: small frames are batched so that every kernel spans several waves of CTAs
    // (default at most 8: longer chains -- up to kMaxGroupFrames on request -- make the kernels a little more efficient in
    // isolation, K7 0.60 -> 0.66 of its roofline at 16 frames, but the pipeline coarser: same device-resident rate, lower end-to-end rate)
    if (frames_per_group == 0) frames_per_group = (int)std::max<long long>(1, std::min<long long>(8, 8500000ll / ((long long)width * height)));
    frames_per_group = std::min(frames_per_group, (int)kMaxGroupFrames);
    c->support_cap = g.Wc * g.Hc + 6;
    c->tri_cap = 2 * c->support_cap + 8;
    while ((1 << c->map_tag_shift) < c->tri_cap) c->map_tag_shift++;
    c->map_tag_max = (1 << (30 - c->map_tag_shift)) - 1;
    if (c->map_tag_max < 1) return ELAS_B200_E_UNSUPPORTED;
    // scan-conversion work units per image: every triangle is at least one unit, large ones split into 32-column x
    // 32-row pieces of their bounding boxes.  Triangles that do not fit the list are scan-converted whole by one
    // warp each (FrameHeader::ovf_from), so this is a performance knob, not a limit.
    c->unit_cap = c->tri_cap + 4 * ((width + 31) / 32) * ((height + kRasterBandRows - 1) / kRasterBandRows) + 64;
    c->mesh_device = mesh_on_device(g, *p);
    // Lattices that do not fit one CTA's shared memory (beyond ~1500x800) are filtered and triangulated out of L2 by the
    // same single-CTA kernels: milliseconds instead of tens of microseconds per frame, but many frame groups run them
    // side by side (32 hardware queues, see the constructor below) and no host core is needed -- measured at 1920x1080:
    // 4.6 k pairs/s against 3.6 k with the host stage on 15 cores; at 4096x2160: 0.92 k against 0.95 k.
    // ELAS_B200_HOST_STAGE=1 forces the host stage, =0 the device mesh stage wherever the parameters allow it
    if (const char* e = std::getenv("ELAS_B200_HOST_STAGE")) c->mesh_device = std::atoi(e) ? false : mesh_on_device(g, *p);
    // both SAD kernels stage their descriptor strips (segment + disparity range) in shared memory: at most
    // 200 KB per CTA, i.e. disp_max up to ~700 for the support search
    if (matching_smem_bytes(g, *p) > 200 * 1024 || support_smem_bytes(g, *p) > 200 * 1024 ||
        g.plane_radius >= 16) return ELAS_B200_E_UNSUPPORTED;
    std::vector<int32_t> prior = make_prior(*p, g.dn);
    // the matching kernel packs (cost, evaluation order) into one 32-bit key: costs must stay below 2^15,
    // and relies on the prior never being positive (-log(1 + e/gamma)/beta <= 0 for gamma, beta > 0; k_matching.cu)
    for (int k = 0; k <= g.plane_radius && k < g.dn; k++)
        if (prior[k] > 0 || prior[k] < -5000) return ELAS_B200_E_UNSUPPORTED;
    CK(cudaMalloc(&c->d_prior, prior.size() * 4));
    CK(cudaMemcpy(c->d_prior, prior.data(), prior.size() * 4, cudaMemcpyHostToDevice));
    c->prior_host = prior;
    c->launches_at_create = launches_issued();
    if (const char* e = std::getenv("ELAS_B200_NARROW_D2")) c->narrow_d2 = std::max(0, std::min(2, std::atoi(e)));
    {
        GroupStrides& st = c->st;
        const size_t cells = (size_t)g.gw * g.gh;
        st.img = (size_t)g.bpl * g.H;                 // bytes
        st.desc = (size_t)g.W * g.H;                  // uint4
        st.dcan = ((size_t)g.Wc * g.Hc + 7) & ~(size_t)7;   // int16
        st.support = 3 * (size_t)c->support_cap;      // int32
        st.tri = 3 * (size_t)c->tri_cap;
        st.units = 2 * (size_t)c->unit_cap;
        st.traster = (size_t)c->tri_cap;              // TriRaster
        st.planes = 6 * (size_t)c->tri_cap;           // float
        st.scratch = 4 * cells * g.gwords;            // uint32: two buffers of [2][cells][gwords]
        st.grid = cells * g.gwords;
        st.lists = cells * kGridListStride;           // uint16
        st.map = (size_t)map_pitch(g) * g.H;          // int32
        st.D = (size_t)g.Dw * g.Dh;                   // float
        st.mesh_scratch = 22 * (size_t)c->support_cap + 8;
        st.lat_work = lattice_work_ints(g);
        st.seg_nodes = segment_node_ints(g);
    }
    for (int i = 0; i < n_groups; i++) {
        c->groups.emplace_back(new Group);
        if (int32_t rc = alloc_group(c.get(), *c->groups.back(), frames_per_group)) {
            for (auto& s : c->groups) free_group(*s);
            cudaFree(c->d_prior);
            return rc;
        }
    }
    {
        // workers only enqueue launch chains and collect results (device mesh stage): a few are enough; with the
        // host stage every worker also runs lattice filters and triangulations, one per core pays off
        int cores = (int)std::thread::hardware_concurrency();
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof set, &set) == 0) cores = CPU_COUNT(&set);
        if (n_workers <= 0) n_workers = c->mesh_device ? std::max(1, std::min(cores, (n_groups + 1) / 2)) : std::max(1, cores);
        if (const char* e = std::getenv("ELAS_B200_WORKERS")) n_workers = std::atoi(e);
        n_workers = std::max(1, std::min(n_workers, n_groups));
    }
    for (int i = 0; i < n_workers; i++) c->workers.emplace_back(worker_main, c.get(), i, n_workers);
    *out = c.release();
    return ELAS_B200_OK;
}

void elas_b200_destroy(elas_b200_ctx* c)
{
    if (!c) return;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        c->stopping = true;
    }
    c->cv_work.notify_all();
    for (auto& t : c->workers) t.join();
    cudaSetDevice(c->device);
    for (auto& s : c->groups) { cudaStreamSynchronize(s->stream); cudaStreamSynchronize(s->copy_stream); free_group(*s); }
    cudaFree(c->d_prior);
    cudaFree(c->d_flush);
    delete c;
}

int32_t elas_b200_frames_per_group(elas_b200_ctx* c) { return c && !c->groups.empty() ? c->groups[0]->cap : 0; }
int32_t elas_b200_mesh_on_device(elas_b200_ctx* c) { return c && c->mesh_device ? 1 : 0; }

int32_t elas_b200_process_ctx(elas_b200_ctx* c, int32_t slot, const uint8_t* I1, const uint8_t* I2,
                              float* D1, float* D2, int32_t bytes_per_line)
{
    if (!c || slot < 0 || slot >= (int)c->groups.size() || !I1 || !I2 || !D1 || !D2 || bytes_per_line < c->g.W)
        return ELAS_B200_E_BAD_ARG;
    std::lock_guard<std::mutex> batch(c->batch_mu);
    CK(cudaSetDevice(c->device));
    return run_frame(c, *c->groups[slot], I1, I2, D1, D2, bytes_per_line, false);
}

int32_t elas_b200_process_batch(elas_b200_ctx* c, int32_t n, const uint8_t* const* I1, const uint8_t* const* I2,
                                float* const* D1, float* const* D2, int32_t bytes_per_line, int32_t* status)
{
    return run_batch(c, n, I1, I2, D1, D2, bytes_per_line, status, false);
}

int32_t elas_b200_process_batch_device(elas_b200_ctx* c, int32_t n, const uint8_t* const* dI1,
                                       const uint8_t* const* dI2, float* const* dD1, float* const* dD2,
                                       int32_t bytes_per_line, int32_t* status)
{
    return run_batch(c, n, dI1, dI2, dD1, dD2, bytes_per_line, status, true);
}

// ---- several GPUs from one process: frames are sharded round-robin over per-device contexts -----------
struct elas_b200_multi {
    std::vector<elas_b200_ctx*> ctx;
};

int32_t elas_b200_multi_create(elas_b200_multi** out, const int32_t* devices, int32_t n_devices, const elas_b200_params* p,
                               int32_t width, int32_t height, int32_t n_groups, int32_t frames_per_group, int32_t n_workers)
{
    if (!out || n_devices < 1 || n_devices > 64) return ELAS_B200_E_BAD_ARG;
    *out = nullptr;
    std::unique_ptr<elas_b200_multi> m(new elas_b200_multi);
    for (int i = 0; i < n_devices; i++) {
        elas_b200_ctx* c = nullptr;
        // every context receives the same parameter block: the path's one "broadcast"
        const int32_t rc = elas_b200_create_grouped(&c, devices ? devices[i] : i, p, width, height, n_groups, frames_per_group, n_workers);
        if (rc) {
            for (elas_b200_ctx* x : m->ctx) elas_b200_destroy(x);
            return rc;
        }
        m->ctx.push_back(c);
    }
    *out = m.release();
    return ELAS_B200_OK;
}

void elas_b200_multi_destroy(elas_b200_multi* m)
{
    if (!m) return;
    for (elas_b200_ctx* c : m->ctx) elas_b200_destroy(c);
    delete m;
}

int32_t elas_b200_multi_device_count(elas_b200_multi* m) { return m ? (int32_t)m->ctx.size() : 0; }
elas_b200_ctx* elas_b200_multi_context(elas_b200_multi* m, int32_t i) { return m && i >= 0 && i < (int)m->ctx.size() ? m->ctx[i] : nullptr; }

int32_t elas_b200_multi_process_batch(elas_b200_multi* m, int32_t n, const uint8_t* const* I1, const uint8_t* const* I2,
                                      float* const* D1, float* const* D2, int32_t bytes_per_line, int32_t* status)
{
    if (!m || n < 0 || !I1 || !I2 || !D1 || !D2) return ELAS_B200_E_BAD_ARG;
    const int nd = (int)m->ctx.size();
    struct Shard {
        std::vector<const uint8_t*> I1, I2; std::vector<float*> D1, D2; std::vector<int32_t> status; int32_t rc = 0;
    };
    std::vector<Shard> shard(nd);
    for (int i = 0; i < n; i++) {                       // frame i -> device i mod nd
        Shard& s = shard[i % nd];
        s.I1.push_back(I1[i]); s.I2.push_back(I2[i]); s.D1.push_back(D1[i]); s.D2.push_back(D2[i]);
    }
    std::vector<std::thread> th;
    for (int k = 0; k < nd; k++) {
        Shard& s = shard[k];
        s.status.assign(s.I1.size(), 0);
        if (s.I1.empty()) continue;
        th.emplace_back([&s, k, m, bytes_per_line] {
            s.rc = run_batch(m->ctx[k], (int32_t)s.I1.size(), s.I1.data(), s.I2.data(), s.D1.data(), s.D2.data(),
                             bytes_per_line, s.status.data(), false);
        });
    }
    for (auto& t : th) t.join();
    int32_t worst = 0;
    for (int k = 0; k < nd; k++) worst = std::min(worst, shard[k].rc);
    if (status) for (int i = 0; i < n; i++) status[i] = shard[i % nd].status[i / nd];
    return worst;
}

int32_t elas_b200_process(const elas_b200_params* p, const uint8_t* I1, const uint8_t* I2,
                          float* D1, float* D2, const int32_t dims[3])
{
    if (!p || !I1 || !I2 || !D1 || !D2 || !dims) return ELAS_B200_E_BAD_ARG;
    int device = 0;
    if (elas_b200_device_count() < 1) return ELAS_B200_E_NO_DEVICE;
    if (const char* e = std::getenv("ELAS_B200_DEVICE")) device = std::atoi(e);
    CacheKey key;
    std::memset(&key, 0, sizeof key);
    key.device = device; key.W = dims[0]; key.H = dims[1]; key.p = *p;
    elas_b200_ctx* c = nullptr;
    std::lock_guard<std::mutex> lk(g_cache_mu);          // one call at a time, from any thread
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
        if (g_cache.size() >= kCacheMax) {
            auto victim = g_cache.begin();
            for (auto j = g_cache.begin(); j != g_cache.end(); ++j) if (j->second.last_use < victim->second.last_use) victim = j;
            elas_b200_destroy(victim->second.ctx);
            g_cache.erase(victim);
        }
        if (int32_t rc = elas_b200_create(&c, device, p, dims[0], dims[1], 1)) return rc;
        g_cache[key] = CacheEntry{c, ++g_cache_tick};
    } else {
        c = it->second.ctx;
        it->second.last_use = ++g_cache_tick;
    }
    return elas_b200_process_ctx(c, 0, I1, I2, D1, D2, dims[2]);
}

// ---- D1's consumers in StereoThread::run ---------------------------------------
static int32_t view_buffers(elas_b200_ctx* c, Group& s)
{
    if (!s.d_view) CK(cudaMalloc(&s.d_view, 5 * (size_t)c->g.W * c->g.H * sizeof(float)));
    return ELAS_B200_OK;
}

int32_t elas_b200_colormap(elas_b200_ctx* c, int32_t slot, const float* D1, float* color)
{
    if (!c || slot < 0 || slot >= (int)c->groups.size() || !color) return ELAS_B200_E_BAD_ARG;
    std::lock_guard<std::mutex> batch(c->batch_mu);
    CK(cudaSetDevice(c->device));
    Group& s = *c->groups[slot];
    if (int32_t rc = view_buffers(c, s)) return rc;
    const size_t nd = (size_t)c->g.Dw * c->g.Dh;
    const float* src = s.last_D1;
    if (D1) {
        // staged behind the colour planes; a device pointer is copied device->device
        CK(cudaMemcpyAsync(s.d_view + 3 * nd, D1, nd * 4, cudaMemcpyDefault, s.stream));
        src = s.d_view + 3 * nd;
    }
    if (!src) return ELAS_B200_E_BAD_ARG;                 // no frame has run through this slot yet
    launch_colormap((int)nd, src, s.d_view, s.stream);
    CK(cudaMemcpyAsync(color, s.d_view, 3 * nd * 4, cudaMemcpyDefault, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    CK(cudaGetLastError());
    return ELAS_B200_OK;
}

int32_t elas_b200_reproject(elas_b200_ctx* c, int32_t slot, const uint8_t* I1, int32_t bytes_per_line,
                            const float* D1, const elas_b200_view* view,
                            float* I, float* D, float* X, float* Y, float* Z)
{
    if (!c || slot < 0 || slot >= (int)c->groups.size() || !view || !I || !D || !X || !Y || !Z) return ELAS_B200_E_BAD_ARG;
    if (c->p.subsampling) return ELAS_B200_E_UNSUPPORTED;            // createCurrentMap reads D1 at full resolution
    if (I1 && bytes_per_line < c->g.W) return ELAS_B200_E_BAD_ARG;
    std::lock_guard<std::mutex> batch(c->batch_mu);
    CK(cudaSetDevice(c->device));
    Group& s = *c->groups[slot];
    if (int32_t rc = view_buffers(c, s)) return rc;
    const FrameGeom& g = c->g;
    const size_t n = (size_t)g.W * g.H;
    const float* src = s.last_D1;
    if (D1) {
        // the slot's scratch planes are free between frames
        CK(cudaMemcpyAsync(s.d_tmp, D1, n * 4, cudaMemcpyDefault, s.stream));
        src = s.d_tmp;
    }
    if (!src) return ELAS_B200_E_BAD_ARG;
    if (I1) { if (int32_t rc = copy_image_in(c, s.d_img[0], I1, bytes_per_line, s.stream, nullptr)) return rc; }
    float* out[5] = {s.d_view, s.d_view + n, s.d_view + 2 * n, s.d_view + 3 * n, s.d_view + 4 * n};
    launch_reproject(g.W, g.H, s.d_img[0], g.bpl, src, *view, out[0], out[1], out[2], out[3], out[4], s.stream);
    float* user[5] = {I, D, X, Y, Z};
    for (int k = 0; k < 5; k++) CK(cudaMemcpyAsync(user[k], out[k], n * 4, cudaMemcpyDefault, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    CK(cudaGetLastError());
    return ELAS_B200_OK;
}

int32_t elas_b200_fuse(elas_b200_ctx* c, int32_t slot, const elas_b200_view* view, const elas_b200_map3d* prev,
                       const elas_b200_map3d* cur, float* points_prev, int32_t* n_prev, float* points_curr, int32_t* n_curr)
{
    if (!c || slot < 0 || slot >= (int)c->groups.size() || !view || !cur || !points_curr || !n_curr) return ELAS_B200_E_BAD_ARG;
    if (!cur->I || !cur->D || !cur->X || !cur->Y || !cur->Z) return ELAS_B200_E_BAD_ARG;
    const bool has_prev = prev && prev->I && prev->D && prev->X && prev->Y && prev->Z;      // stereothread.cpp:296-300
    if (has_prev && (!points_prev || !n_prev)) return ELAS_B200_E_BAD_ARG;
    if (c->p.subsampling) return ELAS_B200_E_UNSUPPORTED;                 // the maps are full-resolution (createCurrentMap)
    std::lock_guard<std::mutex> batch(c->batch_mu);
    CK(cudaSetDevice(c->device));
    Group& s = *c->groups[slot];
    const int W = c->g.W, H = c->g.H;
    const size_t n = (size_t)W * H;
    // device layout: [prev I D X Y Z | cur I D X Y Z | points_prev 4n | points_curr 4n | work ints | 2 counts]; planes are
    // np = n rounded up to 4 floats apart so that the point lists (float4 stores) stay 16-byte aligned
    const size_t np = (n + 3) & ~(size_t)3;
    const size_t floats = 18 * np, ints = fuse_work_ints(W, H) + 2;
    if (!s.d_fuse) CK(cudaMalloc(&s.d_fuse, (floats + ints) * 4));
    float* pm[5]; float* cm[5];
    for (int k = 0; k < 5; k++) { pm[k] = s.d_fuse + (size_t)k * np; cm[k] = s.d_fuse + (size_t)(5 + k) * np; }
    float* d_pp = s.d_fuse + 10 * np; float* d_pc = s.d_fuse + 14 * np;
    int32_t* work = reinterpret_cast<int32_t*>(s.d_fuse + floats);
    int32_t* d_counts = work + fuse_work_ints(W, H);
    float* const user_prev[5] = {has_prev ? prev->I : nullptr, has_prev ? prev->D : nullptr, has_prev ? prev->X : nullptr,
                                 has_prev ? prev->Y : nullptr, has_prev ? prev->Z : nullptr};
    float* const user_cur[5] = {cur->I, cur->D, cur->X, cur->Y, cur->Z};
    cudaStream_t st = s.stream;
    for (int k = 0; k < 5; k++) {
        if (has_prev) CK(cudaMemcpyAsync(pm[k], user_prev[k], n * 4, cudaMemcpyDefault, st));
        CK(cudaMemcpyAsync(cm[k], user_cur[k], n * 4, cudaMemcpyDefault, st));
    }
    if (!launch_fuse(W, H, *view, has_prev ? pm : nullptr, cm, work, d_pp, d_pc, d_counts, st)) return ELAS_B200_E_BAD_ARG;   // singular pose
    CK(cudaGetLastError());
    int32_t counts[2] = {0, 0};
    CK(cudaMemcpyAsync(counts, d_counts, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int k = 0; k < 5; k++) CK(cudaMemcpyAsync(user_cur[k], cm[k], n * 4, cudaMemcpyDefault, st));
    if (has_prev) {
        CK(cudaMemcpyAsync(user_prev[1], pm[1], n * 4, cudaMemcpyDefault, st));
        if (counts[0]) CK(cudaMemcpyAsync(points_prev, d_pp, (size_t)counts[0] * 16, cudaMemcpyDefault, st));
        *n_prev = counts[0];
    } else if (n_prev) *n_prev = 0;
    if (counts[1]) CK(cudaMemcpyAsync(points_curr, d_pc, (size_t)counts[1] * 16, cudaMemcpyDefault, st));
    *n_curr = counts[1];
    CK(cudaStreamSynchronize(st));
    return ELAS_B200_OK;
}

// ---- the feature filters of libviso2's Matcher ------------------------------------------------
namespace {
// device scratch of elas_b200_matcher_filters, kept per device and grown on demand (the Matcher calls it per image)
struct FilterScratch { uint8_t* base = nullptr; size_t bytes = 0; cudaStream_t stream = nullptr; };
std::mutex g_filter_mu;
std::map<int, FilterScratch> g_filter_scratch;
}  // namespace

int32_t elas_b200_matcher_filters(int32_t device, const uint8_t* I, int32_t w, int32_t h, uint8_t* I_du, uint8_t* I_dv,
                                  int16_t* I_f1, int16_t* I_f2, int32_t iters, float* ms_per_pass)
{
    if (!I || !I_du || !I_dv || !I_f1 || !I_f2 || w < 16 || w % 16 || h < 6 || iters < 1) return ELAS_B200_E_BAD_ARG;   // filter.cpp:294 asserts w % 16
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) { cudaGetLastError(); return ELAS_B200_E_NO_DEVICE; }
    if (device < 0 || device >= n_dev) return ELAS_B200_E_BAD_ARG;
    std::lock_guard<std::mutex> lk(g_filter_mu);
    CK(cudaSetDevice(device));
    const size_t n = (size_t)w * h;
    // [in n | du n | dv n | f1 2n | f2 2n], every part 256-byte aligned
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t need = 3 * up(n) + 2 * up(2 * n);
    FilterScratch& fs = g_filter_scratch[device];
    if (fs.bytes < need) {
        if (fs.base) CK(cudaFree(fs.base));
        fs.base = nullptr; fs.bytes = 0;
        CK(cudaMalloc(&fs.base, need));
        fs.bytes = need;
    }
    if (!fs.stream) CK(cudaStreamCreateWithFlags(&fs.stream, cudaStreamNonBlocking));
    uint8_t* d_in = fs.base; uint8_t* d_du = d_in + up(n); uint8_t* d_dv = d_du + up(n);
    int16_t* d_f1 = reinterpret_cast<int16_t*>(d_dv + up(n)); int16_t* d_f2 = reinterpret_cast<int16_t*>(d_dv + up(n) + up(2 * n));
    // memory of this device is used in place (it must be 16-byte aligned: the kernels move words)
    auto on_device = [&](const void* ptr) { return classify(ptr, device).device; };
    const void* all[5] = {I, I_du, I_dv, I_f1, I_f2};
    for (const void* ptr : all) if (on_device(ptr) && (reinterpret_cast<uintptr_t>(ptr) & 15)) return ELAS_B200_E_BAD_ARG;
    const uint8_t* k_in = I;
    if (!on_device(I)) { CK(cudaMemcpyAsync(d_in, I, n, cudaMemcpyDefault, fs.stream)); k_in = d_in; }
    uint8_t* k_du = on_device(I_du) ? I_du : d_du; uint8_t* k_dv = on_device(I_dv) ? I_dv : d_dv;
    int16_t* k_f1 = on_device(I_f1) ? I_f1 : d_f1; int16_t* k_f2 = on_device(I_f2) ? I_f2 : d_f2;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms_per_pass) { CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventRecord(e0, fs.stream)); }
    for (int i = 0; i < iters; i++) launch_matcher_filters(k_in, w, h, k_du, k_dv, k_f1, k_f2, fs.stream);
    CK(cudaGetLastError());
    if (ms_per_pass) CK(cudaEventRecord(e1, fs.stream));
    if (k_du != I_du) CK(cudaMemcpyAsync(I_du, k_du, n, cudaMemcpyDefault, fs.stream));
    if (k_dv != I_dv) CK(cudaMemcpyAsync(I_dv, k_dv, n, cudaMemcpyDefault, fs.stream));
    if (k_f1 != I_f1) CK(cudaMemcpyAsync(I_f1, k_f1, 2 * n, cudaMemcpyDefault, fs.stream));
    if (k_f2 != I_f2) CK(cudaMemcpyAsync(I_f2, k_f2, 2 * n, cudaMemcpyDefault, fs.stream));
    CK(cudaStreamSynchronize(fs.stream));
    if (ms_per_pass) {
        CK(cudaEventElapsedTime(ms_per_pass, e0, e1));
        *ms_per_pass /= iters;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    return ELAS_B200_OK;
}

int32_t elas_b200_time_view(elas_b200_ctx* c, int32_t slot, int32_t iters, float ms_out[2])
{
    if (!c || slot < 0 || slot >= (int)c->groups.size() || iters < 1 || !ms_out || c->p.subsampling) return ELAS_B200_E_BAD_ARG;
    std::lock_guard<std::mutex> batch(c->batch_mu);
    CK(cudaSetDevice(c->device));
    Group& s = *c->groups[slot];
    if (!s.last_D1) return ELAS_B200_E_BAD_ARG;
    if (int32_t rc = view_buffers(c, s)) return rc;
    const FrameGeom& g = c->g;
    const size_t n = (size_t)g.W * g.H;
    elas_b200_view view{};
    view.f = 721.5377f; view.cu = g.W * 0.5f; view.cv = g.H * 0.5f; view.base = 0.54f; view.max_dist = 30.f; view.gain = 1.2f;
    view.H[0] = view.H[5] = view.H[10] = 1.0;
    cudaEvent_t e[3];
    for (auto& x : e) CK(cudaEventCreate(&x));
    CK(cudaEventRecord(e[0], s.stream));
    for (int i = 0; i < iters; i++) launch_colormap((int)n, s.last_D1, s.d_view, s.stream);
    CK(cudaEventRecord(e[1], s.stream));
    for (int i = 0; i < iters; i++)
        launch_reproject(g.W, g.H, s.d_img[0], g.bpl, s.last_D1, view, s.d_view, s.d_view + n, s.d_view + 2 * n,
                         s.d_view + 3 * n, s.d_view + 4 * n, s.stream);
    CK(cudaEventRecord(e[2], s.stream));
    CK(cudaStreamSynchronize(s.stream));
    CK(cudaEventElapsedTime(&ms_out[0], e[0], e[1]));
    CK(cudaEventElapsedTime(&ms_out[1], e[1], e[2]));
    ms_out[0] /= iters; ms_out[1] /= iters;
    for (auto& x : e) cudaEventDestroy(x);
    return ELAS_B200_OK;
}

int32_t elas_b200_stage_capture(elas_b200_ctx* c, int32_t slot, int32_t enable)
{
    if (!c || slot < 0 || slot >= (int)c->groups.size()) return ELAS_B200_E_BAD_ARG;
    c->groups[slot]->capture = enable != 0;
    if (!enable) c->groups[slot]->stages.clear();
    return ELAS_B200_OK;
}

static const std::vector<uint8_t>* find_stage(elas_b200_ctx* c, int32_t slot, const char* name, std::vector<uint8_t>& scratch)
{
    if (!c || !name || slot < 0 || slot >= (int)c->groups.size()) return nullptr;
    Group& s = *c->groups[slot];
    const std::string n(name);
    if (n == "grid1" || n == "grid2") {
        auto it = s.stages.find(n + "_bits");
        if (it == s.stages.end()) return nullptr;
        scratch = expand_grid(c->g, it->second);
        return &scratch;
    }
    auto it = s.stages.find(n);
    return it == s.stages.end() ? nullptr : &it->second;
}

int64_t elas_b200_stage_bytes(elas_b200_ctx* c, int32_t slot, const char* name)
{
    std::vector<uint8_t> scratch;
    const std::vector<uint8_t>* v = find_stage(c, slot, name, scratch);
    return v ? (int64_t)v->size() : -1;
}

int32_t elas_b200_stage_read(elas_b200_ctx* c, int32_t slot, const char* name, void* dst, int64_t cap)
{
    std::vector<uint8_t> scratch;
    const std::vector<uint8_t>* v = find_stage(c, slot, name, scratch);
    if (!v) return ELAS_B200_E_NO_STAGE;
    if ((int64_t)v->size() > cap || !dst) return ELAS_B200_E_BAD_ARG;
    std::memcpy(dst, v->data(), v->size());
    return ELAS_B200_OK;
}

int32_t elas_b200_host_stage(const elas_b200_params* p, int32_t width, int32_t height, int16_t* dcan,
                             int32_t* support, int32_t support_cap, int32_t* tri1, int32_t* tri2,
                             float* planes1, float* planes2, int32_t tri_cap, int32_t n_out[3])
{
    if (!p || !dcan || !support || !tri1 || !tri2 || !planes1 || !planes2 || !n_out) return ELAS_B200_E_BAD_ARG;
    const FrameGeom g = make_geom(*p, width, height);
    HostStage hs;
    const int n = hs.run(g, *p, dcan, false, true);
    n_out[0] = n; n_out[1] = (int)hs.tri[0].size() / 3; n_out[2] = (int)hs.tri[1].size() / 3;
    if (n > support_cap || n_out[1] > tri_cap || n_out[2] > tri_cap) return ELAS_B200_E_BAD_ARG;
    std::memcpy(support, hs.support.data(), hs.support.size() * 4);
    if (n < 3) return ELAS_B200_E_FEW_SUPPORT;
    std::memcpy(tri1, hs.tri[0].data(), hs.tri[0].size() * 4);
    std::memcpy(tri2, hs.tri[1].data(), hs.tri[1].size() * 4);
    std::memcpy(planes1, hs.planes[0].data(), hs.planes[0].size() * 4);
    std::memcpy(planes2, hs.planes[1].data(), hs.planes[1].size() * 4);
    return ELAS_B200_OK;
}

int64_t elas_b200_launch_count(elas_b200_ctx* c)
{
    return c ? launches_issued() - c->launches_at_create : launches_issued();
}

int32_t elas_b200_host_times(elas_b200_ctx* c, double ms_out[5], int64_t* frames, int32_t reset)
{
    if (!c || !ms_out) return ELAS_B200_E_BAD_ARG;
    ms_out[0] = c->ns_submit * 1e-6; ms_out[1] = c->ns_wait * 1e-6; ms_out[2] = c->ns_host * 1e-6;
    ms_out[3] = c->ns_finish * 1e-6; ms_out[4] = 0.0;
    if (frames) *frames = c->frames;
    if (reset) { c->ns_submit = 0; c->ns_wait = 0; c->ns_host = 0; c->ns_finish = 0; c->frames = 0; }
    return ELAS_B200_OK;
}

int32_t elas_b200_stage_timing(elas_b200_ctx* c, int32_t enable)
{
    if (!c) return ELAS_B200_E_BAD_ARG;
    c->timing = enable != 0;
    return ELAS_B200_OK;
}

int32_t elas_b200_stage_times(elas_b200_ctx* c, int32_t slot, const char** names_out, float* ms_out, int32_t cap)
{
    if (!c || slot < 0 || slot >= (int)c->groups.size()) return ELAS_B200_E_BAD_ARG;
    Group& s = *c->groups[slot];
    int n = 0;
    for (auto& e : s.timer.last) {
        if (n >= cap) break;
        // names point into the slot's mark table, stable for the life of the context
        for (auto& m : s.timer.marks) if (m.first == e.first) names_out[n] = m.first.c_str();
        ms_out[n] = e.second;
        n++;
    }
    return n;
}

float elas_b200_time_matching(elas_b200_ctx* c, int32_t slot, int32_t iters, int32_t flush_l2)
{
    return elas_b200_time_matching_ex(c, slot, iters, flush_l2, nullptr);
}

float elas_b200_time_matching_ex(elas_b200_ctx* c, int32_t slot, int32_t iters, int32_t flush_l2, int32_t* frames_per_launch)
{
    if (!c || slot < 0 || slot >= (int)c->groups.size() || iters < 1) return -1.f;
    Group& s = *c->groups[slot];
    if (!s.tables_valid || s.last_n < 1) return -1.f;
    std::lock_guard<std::mutex> batch(c->batch_mu);
    if (cudaSetDevice(c->device) != cudaSuccess) return -1.f;
    if (flush_l2 && !c->d_flush) {
        c->flush_bytes = (size_t)512 << 20;       // 4x the 126 MB L2
        if (cudaMalloc(&c->d_flush, c->flush_bytes) != cudaSuccess) return -1.f;
    }
    if (frames_per_launch) *frames_per_launch = s.last_n;
    // K7 writes the raw maps; plane 0 doubles as the finished left map of the group's last chain, which
    // elas_b200_colormap / _reproject may no longer use afterwards
    s.last_D1 = nullptr;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double total = 0;
    for (int i = 0; i < iters; i++) {
        if (flush_l2) cudaMemsetAsync(c->d_flush, i & 0xff, c->flush_bytes, s.stream);
        cudaEventRecord(e0, s.stream);
        launch_matching(c->g, c->p, match_buffers(c, s), s.last_n, s.map_tag << c->map_tag_shift, c->map_tag_shift, s.stream);
        cudaEventRecord(e1, s.stream);
        if (cudaStreamSynchronize(s.stream) != cudaSuccess) { total = -1; break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        total += ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return total < 0 ? -1.f : (float)(total / iters);
}

}  // extern "C"
