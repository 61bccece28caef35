/*
 * elas_b200.h -- C ABI of the B200 (sm_100a) dense-stereo hot path.
 *
 * This is the drop-in boundary for the path that libelas's
 *     void Elas::process(uint8_t* I1, uint8_t* I2, float* D1, float* D2, const int32_t* dims)
 * (reference libelas/src/elas.h:165, implementation libelas/src/elas.cpp:32-170) runs per frame
 * inside stereomapper's StereoThread::run (reference stereomapper/stereothread.cpp:76-114).
 *
 * Plain C: pointers, sizes and PODs only.  No CUDA, torch or C++ types cross this boundary.
 * Every entry point returns 0 on success, ELAS_B200_E_FEW_SUPPORT (1) when fewer than three
 * support points were found (D1/D2 are then filled with -10, the reference leaves them
 * untouched: elas.cpp:69-75), and a negative ELAS_B200_E_* code on a runtime error.
 * There is no CPU fallback behind this ABI: without a usable CUDA device every compute entry
 * point fails with ELAS_B200_E_NO_DEVICE.
 */
#ifndef ELAS_B200_H_
#define ELAS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ELAS_B200_OK               0
#define ELAS_B200_E_FEW_SUPPORT    1   /* elas.cpp:69-75 */
#define ELAS_B200_E_NO_DEVICE     -1
#define ELAS_B200_E_CUDA          -2
#define ELAS_B200_E_BAD_ARG       -3
#define ELAS_B200_E_UNSUPPORTED   -4
#define ELAS_B200_E_NO_STAGE      -5

/* Presets of Elas::parameters (elas.h:56, :93-118, :121-146). */
#define ELAS_B200_ROBOTICS   0
#define ELAS_B200_MIDDLEBURY 1

/*
 * POD mirror of Elas::parameters (elas.h:59-85): same 23 fields in the same order, the four
 * C++ bools widened to int32 so the layout is the same from C, C++, Python ctypes and cgo.
 */
typedef struct elas_b200_params {
    int32_t disp_min;               /* elas.h:61 */
    int32_t disp_max;               /* elas.h:62 */
    float   support_threshold;      /* elas.h:63 */
    int32_t support_texture;        /* elas.h:64 */
    int32_t candidate_stepsize;     /* elas.h:65 */
    int32_t incon_window_size;      /* elas.h:66 */
    int32_t incon_threshold;        /* elas.h:67 */
    int32_t incon_min_support;      /* elas.h:68 */
    int32_t add_corners;            /* elas.h:69 (bool) */
    int32_t grid_size;              /* elas.h:70 */
    float   beta;                   /* elas.h:71 */
    float   gamma;                  /* elas.h:72 */
    float   sigma;                  /* elas.h:73 */
    float   sradius;                /* elas.h:74 */
    int32_t match_texture;          /* elas.h:75 */
    int32_t lr_threshold;           /* elas.h:76 */
    float   speckle_sim_threshold;  /* elas.h:77 */
    int32_t speckle_size;           /* elas.h:78 */
    int32_t ipol_gap_width;         /* elas.h:79 */
    int32_t filter_median;          /* elas.h:80 (bool) */
    int32_t filter_adaptive_mean;   /* elas.h:81 (bool) */
    int32_t postprocess_only_left;  /* elas.h:82 (bool) */
    int32_t subsampling;            /* elas.h:83 (bool) */
} elas_b200_params;

/* Fills *p with the ROBOTICS or MIDDLEBURY preset (elas.h:88-147). */
void elas_b200_default_params(elas_b200_params* p, int32_t setting);

/* The parameter set stereomapper's StereoThread::run uses (stereothread.cpp:76-80):
 * ROBOTICS + postprocess_only_left=1, filter_adaptive_mean=1, support_texture=30. */
void elas_b200_stereomapper_params(elas_b200_params* p);

/* ------------------------------------------------------------------------------------------
 * 1. The synchronous drop-in call: what a replacement Elas::process binds (elas.h:165).
 *    Host pointers; I1/I2 are uint8 rows of dims[2] bytes, dims = {width, height, bytes/line};
 *    D1/D2 are dense float maps of width x height (width/2 x height/2 with subsampling),
 *    both non-null.  Caller owns all buffers before and after; nothing is retained.
 *    Safe to call from any thread (stereomapper starts a new QThread per frame): the device
 *    context is cached per (device, width, height, parameter block) and guarded by a mutex.
 * ------------------------------------------------------------------------------------------ */
int32_t elas_b200_process(const elas_b200_params* p,
                          const uint8_t* I1, const uint8_t* I2,
                          float* D1, float* D2, const int32_t dims[3]);

/* ------------------------------------------------------------------------------------------
 * 2. Persistent context for throughput: frames are independent (a fresh Elas object per frame,
 *    stereothread.cpp:113), so a context owns n_slots frame slots, each with its own CUDA
 *    stream, device buffers, pinned staging and a host worker for the sequential middle stage
 *    (lattice filters + Delaunay + plane fit).  No per-frame allocation.
 * ------------------------------------------------------------------------------------------ */
typedef struct elas_b200_ctx elas_b200_ctx;

int32_t elas_b200_create(elas_b200_ctx** out, int32_t device, const elas_b200_params* p,
                         int32_t width, int32_t height, int32_t n_slots);
/* Same with an explicit number of host worker threads for the batch calls (0 = one per available
 * core, never more than slots).  Workers are not tied to slots: give a context more slots than workers
 * (e.g. 2x) and the GPU stays fed while every worker runs host stages. */
int32_t elas_b200_create_ex(elas_b200_ctx** out, int32_t device, const elas_b200_params* p,
                            int32_t width, int32_t height, int32_t n_slots, int32_t n_workers);
/* Frame groups: a context owns n_groups groups of frames_per_group frames each (at most 16; 0 = chosen from the
 * frame size: 8 500 000 / (width * height), between 1 and 8).  The frames of a group share one chain of kernel launches -- every kernel takes the frame as a
 * grid dimension -- so small frames still fill the GPU and the launch cost per frame shrinks.  The whole path of
 * a frame runs on the device (for the parameter sets the device mesh stage covers, which include both presets'
 * ROBOTICS family; others use the host stage of section 3), so n_workers host threads only enqueue chains and
 * collect results.  elas_b200_create[_ex] = groups of one frame ("slots"). */
int32_t elas_b200_create_grouped(elas_b200_ctx** out, int32_t device, const elas_b200_params* p,
                                 int32_t width, int32_t height, int32_t n_groups, int32_t frames_per_group,
                                 int32_t n_workers);
int32_t elas_b200_frames_per_group(elas_b200_ctx* ctx);
/* 1 when lattice filters and Delaunay triangulation run on the GPU for this context, 0 when they use the host stage */
int32_t elas_b200_mesh_on_device(elas_b200_ctx* ctx);
void    elas_b200_destroy(elas_b200_ctx* ctx);

/* One frame through one slot, synchronous, host buffers (same contract as elas_b200_process). */
int32_t elas_b200_process_ctx(elas_b200_ctx* ctx, int32_t slot,
                              const uint8_t* I1, const uint8_t* I2,
                              float* D1, float* D2, int32_t bytes_per_line);

/* n frames pipelined over all slots.  Host buffers; images are tightly strided by
 * bytes_per_line; I1[i], I2[i], D1[i], D2[i] address frame i.  Host<->device copies are part
 * of the call.  status[i] (optional) receives the per-frame return code.  D2[i] may be NULL: the right map of
 * that frame is then not returned (it is still computed, the left/right check needs it) -- stereomapper reads only
 * D1 (stereothread.cpp:116-147), and the return path is what bounds the end-to-end rate (2/3 of the bytes). */
int32_t elas_b200_process_batch(elas_b200_ctx* ctx, int32_t n,
                                const uint8_t* const* I1, const uint8_t* const* I2,
                                float* const* D1, float* const* D2,
                                int32_t bytes_per_line, int32_t* status);

/* Same, but all four buffers per frame are DEVICE pointers on the context's device (images with
 * pitch bytes_per_line, disparity maps dense): nothing but the 37 KB candidate lattice and the
 * support/triangle tables crosses PCIe.  Used by the HBM-resident throughput measurement. */
int32_t elas_b200_process_batch_device(elas_b200_ctx* ctx, int32_t n,
                                       const uint8_t* const* dI1, const uint8_t* const* dI2,
                                       float* const* dD1, float* const* dD2,
                                       int32_t bytes_per_line, int32_t* status);

/* ------------------------------------------------------------------------------------------
 * 2a. Several GPUs from ONE process (a C++ host like stereomapper needs no launcher): one context per device,
 *     every context gets the same parameter block (the path's only "broadcast"), frame i of a batch goes to
 *     device i mod n (stereo pairs are independent: stereothread.cpp:113 builds a fresh Elas per frame), the
 *     per-device batches run concurrently.  devices == NULL: devices 0..n_devices-1.  Host buffers.
 * ------------------------------------------------------------------------------------------ */
typedef struct elas_b200_multi elas_b200_multi;
int32_t elas_b200_multi_create(elas_b200_multi** out, const int32_t* devices, int32_t n_devices,
                               const elas_b200_params* p, int32_t width, int32_t height,
                               int32_t n_groups, int32_t frames_per_group, int32_t n_workers);
void    elas_b200_multi_destroy(elas_b200_multi* m);
int32_t elas_b200_multi_device_count(elas_b200_multi* m);
elas_b200_ctx* elas_b200_multi_context(elas_b200_multi* m, int32_t i);      /* for the introspection entries */
int32_t elas_b200_multi_process_batch(elas_b200_multi* m, int32_t n,
                                      const uint8_t* const* I1, const uint8_t* const* I2,
                                      float* const* D1, float* const* D2,
                                      int32_t bytes_per_line, int32_t* status);

/* ------------------------------------------------------------------------------------------
 * 2b. The consumers of D1 inside StereoThread::run, computed where D1 already is (HBM):
 *     the HSV colour map shown by View2D (stereothread.cpp:116-147) and the back-projected
 *     map StereoThread::createCurrentMap builds for the 3-D reconstruction (:180-255).
 *     D1 == NULL uses the left disparity map the slot's last frame left on the device (no
 *     device->host->device round trip); otherwise D1 is a host or device pointer to a dense
 *     map.  Likewise I1 == NULL uses the slot's device copy of the last left image.  Output
 *     pointers may be host or device memory.  Synchronous.
 * ------------------------------------------------------------------------------------------ */
typedef struct elas_b200_view {
    float  f, cu, cv, base;   /* StereoThread::getIntrinsics, stereothread.cpp:441-447 */
    float  max_dist;          /* stereothread.h:196 */
    float  gain;              /* stereothread.h:181; 0 = no gain ramp (stereothread.cpp:234) */
    double H[12];             /* rows 0..2 of the accumulated pose _H_total (libviso2 Matrix, double) */
} elas_b200_view;

/* color: 3 floats (r,g,b) per pixel of the disparity map (width x height, halved with subsampling). */
int32_t elas_b200_colormap(elas_b200_ctx* ctx, int32_t slot, const float* D1, float* color);

/* I, D, X, Y, Z: width x height floats each.  D = D1 with -1 where z = f*base/d falls outside
 * (0.1, max_dist); X/Y/Z are 0 where the reference writes nothing (d <= 0 or z out of range).
 * Only without subsampling (createCurrentMap indexes D1 at full resolution). */
int32_t elas_b200_reproject(elas_b200_ctx* ctx, int32_t slot, const uint8_t* I1, int32_t bytes_per_line,
                            const float* D1, const elas_b200_view* view,
                            float* I, float* D, float* X, float* Y, float* Z);

/* Fusion of the current map with the previous one: StereoThread::addDisparityMapToReconstruction
 * (stereothread.cpp:290-437).  cur = the map elas_b200_reproject produced for this frame (I, D, X, Y, Z, width x
 * height floats each; host or device memory), fused in place: previous points that project onto a current point
 * closer than 0.2 (L1) are averaged into it, points that land where the current map has none are created there
 * (D = 1).  prev = the fused map the PREVIOUS call returned in `cur` (the reference hands its own previous map over
 * through freed memory, :433-434; this is the evident intent), or NULL for the first frame; prev->D comes back with
 * the merged points invalidated.  view->H = the CURRENT pose.  points_prev / points_curr (capacity width*height x 4
 * floats each) receive (x, y, z, intensity) of the previous points that were kept and of all valid current points,
 * both in the reference's push_back order (u outer, v inner); these are the two lists handed to View3D::addPoints. */
typedef struct elas_b200_map3d { float* I; float* D; float* X; float* Y; float* Z; } elas_b200_map3d;
int32_t elas_b200_fuse(elas_b200_ctx* ctx, int32_t slot, const elas_b200_view* view,
                       const elas_b200_map3d* prev, const elas_b200_map3d* cur,
                       float* points_prev, int32_t* n_prev, float* points_curr, int32_t* n_curr);

/* ------------------------------------------------------------------------------------------
 * 3. Introspection for parity tests and the bench (not used by stereomapper).
 * ------------------------------------------------------------------------------------------ */

/* When enabled, each stage's output of the NEXT frame run through `slot` is kept on the host and
 * can be read back by name.  Names and element types:
 *   "desc1","desc2"      u8  [H][W][16]   descriptor.cpp:88-120 (border = 0)
 *   "dcan_raw"           i16 [Hc][Wc]     elas.cpp:471-493 (before the lattice filters)
 *   "dcan"               i16 [Hc][Wc]     elas.cpp:496-502
 *   "support"            i32 [n][3]       elas.cpp:505-517 (u,v,d)
 *   "tri1","tri2"        i32 [t][3]       elas.cpp:534-600
 *   "planes1","planes2"  f32 [t][6]       elas.cpp:605-680 (t1a,t1b,t1c,t2a,t2b,t2c)
 *   "grid1","grid2"      i32 [gh][gw][dmax+2]  elas.cpp:684-780 (expanded to the reference layout)
 *   "D1_raw","D2_raw"    f32 [H][W]       elas.cpp:960-1118
 *   "D1_lr","D2_lr"      f32              elas.cpp:1122-1204
 *   "D1_seg","D2_seg"    f32              elas.cpp:1208-1326
 *   "D1_gap","D2_gap"    f32              elas.cpp:1330-1530
 *   "D1_mean","D2_mean"  f32              elas.cpp:1535-1754
 */
int32_t elas_b200_stage_capture(elas_b200_ctx* ctx, int32_t slot, int32_t enable);
int64_t elas_b200_stage_bytes(elas_b200_ctx* ctx, int32_t slot, const char* name);
int32_t elas_b200_stage_read(elas_b200_ctx* ctx, int32_t slot, const char* name,
                             void* dst, int64_t capacity_bytes);

/* The host middle stage on its own (no device needed): in-place lattice filters
 * (elas.cpp:174-279), support list (:505-523), Triangle-compatible Delaunay of both point sets
 * (:534-600) and disparity planes (:605-680).  dcan is the candidate lattice [Hc][Wc] as K2
 * produces it (filtered in place).  Outputs hold up to support_cap / tri_cap entries; counts are
 * returned through n_out = {n_support, n_tri1, n_tri2}.  Returns 0, or ELAS_B200_E_FEW_SUPPORT. */
int32_t elas_b200_host_stage(const elas_b200_params* p, int32_t width, int32_t height, int16_t* dcan,
                             int32_t* support, int32_t support_cap,
                             int32_t* tri1, int32_t* tri2, float* planes1, float* planes2,
                             int32_t tri_cap, int32_t n_out[3]);

/* Number of CUDA kernels launched by this context since creation (bench: gpu_launches). */
int64_t elas_b200_launch_count(elas_b200_ctx* ctx);

/* Device time of the last frame run through `slot`, per stage, in milliseconds (CUDA events on
 * the slot's stream).  names_out/ms_out hold up to `cap` entries; returns the entry count.
 * Timing must have been switched on with elas_b200_stage_timing(ctx, 1). */
int32_t elas_b200_stage_timing(elas_b200_ctx* ctx, int32_t enable);
int32_t elas_b200_stage_times(elas_b200_ctx* ctx, int32_t slot,
                              const char** names_out, float* ms_out, int32_t cap);

/* Host-side wall time spent per frame phase, summed over all frames since the last reset, in
 * milliseconds: {enqueueing launch chains, waiting for results, host stage (contexts that use it), collecting
 * results incl. widening D2, 0}. */
int32_t elas_b200_host_times(elas_b200_ctx* ctx, double ms_out[5], int64_t* frames, int32_t reset);

/* Bench hook: runs only the dense matching kernel (left+right, elas.cpp:960-1118) `iters` times
 * on the tables left in `slot` by the last frame and returns the mean milliseconds per launch
 * (CUDA events on the slot's stream; flush_l2 != 0 rewrites a >L2-sized buffer between
 * launches, outside the timed span).  Negative on error. */
float   elas_b200_time_matching(elas_b200_ctx* ctx, int32_t slot, int32_t iters, int32_t flush_l2);
/* Same; *frames_per_launch receives the number of frames one launch processes (the group's last chain). */
float   elas_b200_time_matching_ex(elas_b200_ctx* ctx, int32_t slot, int32_t iters, int32_t flush_l2,
                                   int32_t* frames_per_launch);

/* Bench hook for the consumers of D1: runs k_colormap and k_reproject `iters` times each on the left map
 * and image the slot's last frame left on the device (nothing crosses PCIe) and returns the mean
 * milliseconds per launch in ms_out = {colour map, back-projection}. */
int32_t elas_b200_time_view(elas_b200_ctx* ctx, int32_t slot, int32_t iters, float ms_out[2]);

/* ------------------------------------------------------------------------------------------
 * 4. The feature filters of libviso2's Matcher.
 * ------------------------------------------------------------------------------------------ */

/* Replaces the three calls of Matcher::computeFeatures (libviso2/src/matcher.cpp:799-801):
 *     filter::sobel5x5(I, I_du, I_dv, dims[2], dims[1]);  filter::blob5x5(I, I_f1, dims[2], dims[1]);
 *     filter::checkerboard5x5(I, I_f2, dims[2], dims[1]);                     (libviso2/src/filter.cpp:474-530)
 * I = bytes_per_line x height bytes (bytes_per_line a multiple of 16 as the reference asserts, height >= 6); the four
 * maps have the same extent (I_du/I_dv uint8, I_f1/I_f2 int16).  Pointers may be host memory or 16-byte-aligned
 * memory of `device`.  Bit-identical to the reference wherever the reference's result is defined; elements it never
 * writes are 0, and I_du/I_dv[n-2], [n-1] (for which the reference reads beyond its temporaries) take those as 0.
 * iters > 1 repeats the kernels (bench); *ms_per_pass, if not NULL, receives the mean device time of one pass. */
int32_t elas_b200_matcher_filters(int32_t device, const uint8_t* I, int32_t bytes_per_line, int32_t height,
                                  uint8_t* I_du, uint8_t* I_dv, int16_t* I_f1, int16_t* I_f2,
                                  int32_t iters, float* ms_per_pass);

/* Library / device description (static string, never freed). */
const char* elas_b200_version(void);
int32_t     elas_b200_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ELAS_B200_H_ */
