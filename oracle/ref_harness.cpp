// TEST INFRASTRUCTURE -- not part of the product.
//
// Step-wise driver around the UNMODIFIED reference libelas.  It is compiled only by oracle/Makefile,
// which takes the reference sources from where they lie (REF=/root/reference/libelas/src) and writes
// nothing but oracle/_ref/libelas_ref.so.  No reference source is copied into this repository: the
// reference translation unit is pulled in with #include at build time so that its private/inline
// stages (elas.h:196-304) can be called one by one and every intermediate can be dumped.
//
// ref_process()      = Elas::process (elas.cpp:32-170) through its public API, nothing else.
// ref_run_stages()   = the same stages in the same order (elas.cpp:61-159), keeping every
//                      intermediate; tests check that both give bit-identical D1/D2.
//
// Determinism: Descriptor's _mm_malloc'ed border is never written by the reference
// (descriptor.cpp:30,91,99).  The Makefile links this .so with -Bsymbolic so
// the posix_memalign below (zero-filled) is what _mm_malloc reaches: border = 0 by construction.

// Built with -fno-access-control so Elas's private stages, members and PODs are reachable.
#include "elas.cpp"               // resolved through -I$(REF); also brings descriptor.h/matrix.h/triangle.h

#include <map>
#include <string>
#include <sys/time.h>
#include "../include/elas_b200.h" // the POD parameter block shared with the product ABI

extern "C" int posix_memalign(void** memptr, size_t alignment, size_t size) noexcept
{
    size_t rounded = (size + alignment - 1) / alignment * alignment;
    if (rounded == 0) rounded = alignment;
    void* p = aligned_alloc(alignment, rounded);
    if (!p) return 12; // ENOMEM
    memset(p, 0, rounded);
    *memptr = p;
    return 0;
}

namespace {

Elas::parameters to_ref(const elas_b200_params* p)
{
    Elas::parameters q(Elas::ROBOTICS);
    q.disp_min = p->disp_min;                     q.disp_max = p->disp_max;
    q.support_threshold = p->support_threshold;   q.support_texture = p->support_texture;
    q.candidate_stepsize = p->candidate_stepsize; q.incon_window_size = p->incon_window_size;
    q.incon_threshold = p->incon_threshold;       q.incon_min_support = p->incon_min_support;
    q.add_corners = p->add_corners != 0;          q.grid_size = p->grid_size;
    q.beta = p->beta; q.gamma = p->gamma; q.sigma = p->sigma; q.sradius = p->sradius;
    q.match_texture = p->match_texture;           q.lr_threshold = p->lr_threshold;
    q.speckle_sim_threshold = p->speckle_sim_threshold;
    q.speckle_size = p->speckle_size;             q.ipol_gap_width = p->ipol_gap_width;
    q.filter_median = p->filter_median != 0;      q.filter_adaptive_mean = p->filter_adaptive_mean != 0;
    q.postprocess_only_left = p->postprocess_only_left != 0;
    q.subsampling = p->subsampling != 0;
    return q;
}

std::map<std::string, std::string> g_stage;   // name -> raw bytes of the last ref_run_stages()

template <class T> void keep(const char* name, const T* data, size_t count)
{
    g_stage[name] = std::string(reinterpret_cast<const char*>(data), count * sizeof(T));
}

double now_ms()
{
    timeval tv; gettimeofday(&tv, nullptr);
    return tv.tv_sec * 1e3 + tv.tv_usec * 1e-3;
}

} // namespace

extern "C" {

void ref_default_params(elas_b200_params* p, int32_t setting)
{
    Elas::parameters q(setting == ELAS_B200_MIDDLEBURY ? Elas::MIDDLEBURY : Elas::ROBOTICS);
    p->disp_min = q.disp_min; p->disp_max = q.disp_max;
    p->support_threshold = q.support_threshold; p->support_texture = q.support_texture;
    p->candidate_stepsize = q.candidate_stepsize; p->incon_window_size = q.incon_window_size;
    p->incon_threshold = q.incon_threshold; p->incon_min_support = q.incon_min_support;
    p->add_corners = q.add_corners; p->grid_size = q.grid_size;
    p->beta = q.beta; p->gamma = q.gamma; p->sigma = q.sigma; p->sradius = q.sradius;
    p->match_texture = q.match_texture; p->lr_threshold = q.lr_threshold;
    p->speckle_sim_threshold = q.speckle_sim_threshold; p->speckle_size = q.speckle_size;
    p->ipol_gap_width = q.ipol_gap_width; p->filter_median = q.filter_median;
    p->filter_adaptive_mean = q.filter_adaptive_mean;
    p->postprocess_only_left = q.postprocess_only_left; p->subsampling = q.subsampling;
}

// The reference through its public API only.  Always returns 0: Elas::process is void, its "fewer than
// 3 support points" early return (elas.cpp:69-75) is visible to the caller only through D1/D2 keeping
// whatever they held before the call (checkers.py pre-fills them with a sentinel).
int32_t ref_process(const elas_b200_params* p, const uint8_t* I1, const uint8_t* I2,
                    float* D1, float* D2, const int32_t* dims)
{
    Elas elas(to_ref(p));
    elas.process(const_cast<uint8_t*>(I1), const_cast<uint8_t*>(I2), D1, D2, dims);
    return 0;
}

// Times `reps` calls of Elas::process; returns the best wall-clock milliseconds, mean in *mean_ms.
double ref_time_process(const elas_b200_params* p, const uint8_t* I1, const uint8_t* I2,
                        float* D1, float* D2, const int32_t* dims, int32_t reps, double* mean_ms)
{
    double best = 1e300, sum = 0;
    for (int32_t r = 0; r < reps; r++) {
        double t0 = now_ms();
        Elas elas(to_ref(p));
        elas.process(const_cast<uint8_t*>(I1), const_cast<uint8_t*>(I2), D1, D2, dims);
        double dt = now_ms() - t0;
        best = dt < best ? dt : best;
        sum += dt;
    }
    if (mean_ms) *mean_ms = sum / (reps > 0 ? reps : 1);
    return best;
}

// Elas::process re-enacted stage by stage (same calls, same order as elas.cpp:35-159).
int32_t ref_run_stages(const elas_b200_params* p, const uint8_t* I1_, const uint8_t* I2_,
                       float* D1, float* D2, const int32_t* dims)
{
    g_stage.clear();
    Elas e(to_ref(p));
    const Elas::parameters& prm = e._param;

    // elas.cpp:35-56
    e._width = dims[0]; e._height = dims[1];
    e._bpl = e._width + 15 - (e._width - 1) % 16;
    const int W = e._width, H = e._height, bpl = e._bpl;
    e._I1 = (uint8_t*)_mm_malloc(bpl * H, 16);
    e._I2 = (uint8_t*)_mm_malloc(bpl * H, 16);
    memset(e._I1, 0, bpl * H); memset(e._I2, 0, bpl * H);
    if (bpl == dims[2]) { memcpy(e._I1, I1_, bpl * H); memcpy(e._I2, I2_, bpl * H); }
    else for (int v = 0; v < H; v++) {
        memcpy(e._I1 + v * bpl, I1_ + v * dims[2], W);
        memcpy(e._I2 + v * bpl, I2_ + v * dims[2], W);
    }

    // elas.cpp:61-62
    Descriptor desc1(e._I1, W, H, bpl, prm.subsampling);
    Descriptor desc2(e._I2, W, H, bpl, prm.subsampling);
    keep("desc1", desc1._I_desc, (size_t)16 * W * H);
    keep("desc2", desc2._I_desc, (size_t)16 * W * H);

    // elas.cpp:449-493 re-enacted with the reference's own computeMatchingDisparity, to expose the
    // candidate lattice before the in-place filters.
    {
        int step = prm.candidate_stepsize;
        if (prm.subsampling) step += step % 2;
        int Wc = 0, Hc = 0;
        for (int u = 0; u < W; u += step) Wc++;
        for (int v = 0; v < H; v += step) Hc++;
        std::vector<int16_t> dcan((size_t)Wc * Hc, 0);
        for (int uc = 1; uc < Wc; uc++) for (int vc = 1; vc < Hc; vc++) {
            int u = uc * step, v = vc * step;
            dcan[vc * Wc + uc] = -1;
            int16_t d = e.computeMatchingDisparity(u, v, desc1._I_desc, desc2._I_desc, false);
            if (d >= 0) {
                int16_t d2 = e.computeMatchingDisparity(u - d, v, desc1._I_desc, desc2._I_desc, true);
                if (d2 >= 0 && abs(d - d2) <= prm.lr_threshold) dcan[vc * Wc + uc] = d;
            }
        }
        keep("dcan_raw", dcan.data(), dcan.size());
        e.removeInconsistentSupportPoints(dcan.data(), Wc, Hc);
        keep("dcan_incon", dcan.data(), dcan.size());
        e.removeRedundantSupportPoints(dcan.data(), Wc, Hc, 5, 1, true);
        e.removeRedundantSupportPoints(dcan.data(), Wc, Hc, 5, 1, false);
        keep("dcan", dcan.data(), dcan.size());
        int32_t lat[2] = {Wc, Hc};
        keep("lattice_dims", lat, 2);
    }

    // elas.cpp:66
    std::vector<Elas::support_pt> sup = e.computeSupportMatches(desc1._I_desc, desc2._I_desc);
    keep("support", reinterpret_cast<const int32_t*>(sup.data()), sup.size() * 3);
    if (sup.size() < 3) { _mm_free(e._I1); _mm_free(e._I2); return 1; }   // elas.cpp:69-75

    // elas.cpp:80-88
    std::vector<Elas::triangle> tri1 = e.computeDelaunayTriangulation(sup, 0);
    std::vector<Elas::triangle> tri2 = e.computeDelaunayTriangulation(sup, 1);
    e.computeDisparityPlanes(sup, tri1);
    e.computeDisparityPlanes(sup, tri2);
    for (int k = 0; k < 2; k++) {
        const std::vector<Elas::triangle>& t = k ? tri2 : tri1;
        std::vector<int32_t> idx(t.size() * 3);
        std::vector<float> pl(t.size() * 6);
        for (size_t i = 0; i < t.size(); i++) {
            idx[3*i] = t[i].c1; idx[3*i+1] = t[i].c2; idx[3*i+2] = t[i].c3;
            pl[6*i] = t[i].t1a; pl[6*i+1] = t[i].t1b; pl[6*i+2] = t[i].t1c;
            pl[6*i+3] = t[i].t2a; pl[6*i+4] = t[i].t2b; pl[6*i+5] = t[i].t2c;
        }
        keep(k ? "tri2" : "tri1", idx.data(), idx.size());
        keep(k ? "planes2" : "planes1", pl.data(), pl.size());
    }

    // elas.cpp:98-105
    int32_t gw = (int32_t)ceil((float)W / (float)prm.grid_size);
    int32_t gh = (int32_t)ceil((float)H / (float)prm.grid_size);
    int32_t grid_dims[3] = {prm.disp_max + 2, gw, gh};
    size_t gsz = (size_t)(prm.disp_max + 2) * gh * gw;
    int32_t* grid1 = (int32_t*)calloc(gsz, sizeof(int32_t));
    int32_t* grid2 = (int32_t*)calloc(gsz, sizeof(int32_t));
    e.createGrid(sup, grid1, grid_dims, 0);
    e.createGrid(sup, grid2, grid_dims, 1);
    keep("grid1", grid1, gsz); keep("grid2", grid2, gsz);
    keep("grid_dims", grid_dims, 3);

    size_t nd = prm.subsampling ? (size_t)(W / 2) * (H / 2) : (size_t)W * H;

    // elas.cpp:110-111
    e.computeDisparity(sup, tri1, grid1, grid_dims, desc1._I_desc, desc2._I_desc, 0, D1);
    e.computeDisparity(sup, tri2, grid2, grid_dims, desc1._I_desc, desc2._I_desc, 1, D2);
    keep("D1_raw", D1, nd); keep("D2_raw", D2, nd);

    // elas.cpp:116
    e.leftRightConsistencyCheck(D1, D2);
    keep("D1_lr", D1, nd); keep("D2_lr", D2, nd);

    // elas.cpp:121-125
    e.removeSmallSegments(D1);
    if (!prm.postprocess_only_left) e.removeSmallSegments(D2);
    keep("D1_seg", D1, nd); keep("D2_seg", D2, nd);

    // elas.cpp:130-134
    e.gapInterpolation(D1);
    if (!prm.postprocess_only_left) e.gapInterpolation(D2);
    keep("D1_gap", D1, nd); keep("D2_gap", D2, nd);

    // elas.cpp:136-146
    if (prm.filter_adaptive_mean) {
        e.adaptiveMean(D1);
        if (!prm.postprocess_only_left) e.adaptiveMean(D2);
    }
    keep("D1_mean", D1, nd); keep("D2_mean", D2, nd);

    // elas.cpp:148-159
    if (prm.filter_median) {
        e.median(D1);
        if (!prm.postprocess_only_left) e.median(D2);
    }
    keep("D1", D1, nd); keep("D2", D2, nd);

    free(grid1); free(grid2);
    _mm_free(e._I1); _mm_free(e._I2);
    return 0;
}

int64_t ref_stage_bytes(const char* name)
{
    auto it = g_stage.find(name);
    return it == g_stage.end() ? -1 : (int64_t)it->second.size();
}

int32_t ref_stage_read(const char* name, void* dst, int64_t cap)
{
    auto it = g_stage.find(name);
    if (it == g_stage.end()) return ELAS_B200_E_NO_STAGE;
    if ((int64_t)it->second.size() > cap) return ELAS_B200_E_BAD_ARG;
    memcpy(dst, it->second.data(), it->second.size());
    return 0;
}

// Single reference stages on caller-provided tables, used to pin the restatement stage by stage.
// Delaunay of the support points (elas.cpp:534-600 -> triangle.cpp:8499); returns the triangle
// count and writes up to cap triangles (3 indices each).
int32_t ref_delaunay(const int32_t* sup, int32_t n, int32_t right_image, int32_t* tri_out, int32_t cap)
{
    if (n < 3) return 0;     // Triangle exit()s below three vertices; Elas::process never gets there (elas.cpp:69-75)
    elas_b200_params dp; ref_default_params(&dp, 0);
    Elas e(to_ref(&dp));
    std::vector<Elas::support_pt> s;
    for (int i = 0; i < n; i++) s.push_back(Elas::support_pt(sup[3*i], sup[3*i+1], sup[3*i+2]));
    std::vector<Elas::triangle> t = e.computeDelaunayTriangulation(s, right_image);
    for (size_t i = 0; i < t.size() && (int32_t)i < cap; i++) {
        tri_out[3*i] = t[i].c1; tri_out[3*i+1] = t[i].c2; tri_out[3*i+2] = t[i].c3;
    }
    return (int32_t)t.size();
}

} // extern "C"
