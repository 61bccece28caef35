// Drop-in replacement for libelas/src/elas.cpp: Elas::process forwards to the B200 library.
//
// The library is bound at run time (dlopen) so that stereomapper.pro, which lists this file in
// SOURCES (stereomapper.pro:25) and links libviso2's matrix/triangle/filter objects, needs no edit:
//   ELAS_B200_LIB=/path/to/libelas_b200.so   (default: libelas_b200.so on the loader path)
#include "elas.h"

#include <dlfcn.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

namespace {

// mirror of elas_b200_params (include/elas_b200.h): the reference's fields with bools widened to int32
struct abi_params {
    int32_t disp_min, disp_max; float support_threshold; int32_t support_texture, candidate_stepsize,
    incon_window_size, incon_threshold, incon_min_support, add_corners, grid_size; float beta, gamma, sigma,
    sradius; int32_t match_texture, lr_threshold; float speckle_sim_threshold; int32_t speckle_size,
    ipol_gap_width, filter_median, filter_adaptive_mean, postprocess_only_left, subsampling;
};

typedef void (*default_params_fn)(abi_params*, int32_t);
typedef int32_t (*process_fn)(const abi_params*, const uint8_t*, const uint8_t*, float*, float*, const int32_t*);

struct Binding {
    void* handle = nullptr;
    default_params_fn default_params = nullptr;
    process_fn process = nullptr;
    bool tried = false;
};

Binding& binding()
{
    static Binding b;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!b.tried) {
        b.tried = true;
        const char* path = getenv("ELAS_B200_LIB");
        b.handle = dlopen(path ? path : "libelas_b200.so", RTLD_NOW | RTLD_LOCAL);
        if (!b.handle) {
            fprintf(stderr, "ERROR: cannot load the B200 stereo library: %s\n", dlerror());
        } else {
            b.default_params = (default_params_fn)dlsym(b.handle, "elas_b200_default_params");
            b.process = (process_fn)dlsym(b.handle, "elas_b200_process");
        }
    }
    return b;
}

abi_params to_abi(const Elas::parameters& p)
{
    abi_params a;
    a.disp_min = p.disp_min; a.disp_max = p.disp_max; a.support_threshold = p.support_threshold;
    a.support_texture = p.support_texture; a.candidate_stepsize = p.candidate_stepsize;
    a.incon_window_size = p.incon_window_size; a.incon_threshold = p.incon_threshold;
    a.incon_min_support = p.incon_min_support; a.add_corners = p.add_corners; a.grid_size = p.grid_size;
    a.beta = p.beta; a.gamma = p.gamma; a.sigma = p.sigma; a.sradius = p.sradius;
    a.match_texture = p.match_texture; a.lr_threshold = p.lr_threshold;
    a.speckle_sim_threshold = p.speckle_sim_threshold; a.speckle_size = p.speckle_size;
    a.ipol_gap_width = p.ipol_gap_width; a.filter_median = p.filter_median;
    a.filter_adaptive_mean = p.filter_adaptive_mean; a.postprocess_only_left = p.postprocess_only_left;
    a.subsampling = p.subsampling;
    return a;
}

}  // namespace

Elas::parameters::parameters(setting s)
{
    // values of the reference presets (elas.h:93-118 ROBOTICS, :121-146 MIDDLEBURY); kept here so the
    // constructor works even before the library is loaded
    const bool mb = s == MIDDLEBURY;
    disp_min = 0;                  disp_max = 255;
    support_threshold = mb ? 0.95f : 0.85f;
    support_texture = 10;          candidate_stepsize = 5;
    incon_window_size = 5;         incon_threshold = 5;
    incon_min_support = 5;         add_corners = mb;
    grid_size = 20;                beta = 0.02f;
    gamma = mb ? 5.f : 3.f;        sigma = 1.f;
    sradius = mb ? 3.f : 2.f;      match_texture = mb ? 0 : 1;
    lr_threshold = 2;              speckle_sim_threshold = 1.f;
    speckle_size = 200;            ipol_gap_width = mb ? 5000 : 3;
    filter_median = mb;            filter_adaptive_mean = !mb;
    postprocess_only_left = !mb;   subsampling = false;
}

void Elas::process(uint8_t* I1, uint8_t* I2, float* D1, float* D2, const int32_t* dims)
{
    Binding& b = binding();
    int32_t rc = -1;
    if (b.process) {
        const abi_params a = to_abi(_param);
        rc = b.process(&a, I1, I2, D1, D2, dims);
    }
    if (rc == 1) {
        // same message as the reference (elas.cpp:71); D1/D2 have been filled with -10 by the library
        printf("ERROR: Need at least 3 support points!\n");
    } else if (rc != 0) {
        fprintf(stderr, "ERROR: elas_b200_process failed (%d); disparity maps set to invalid\n", rc);
        const int32_t w = _param.subsampling ? dims[0] / 2 : dims[0], h = _param.subsampling ? dims[1] / 2 : dims[1];
        for (int64_t i = 0; i < (int64_t)w * h; i++) { D1[i] = -10; D2[i] = -10; }
    }
}
