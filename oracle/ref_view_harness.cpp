// TEST INFRASTRUCTURE -- not part of the product.
//
// Pins the oracle's restatement of the two per-pixel loops that consume D1 inside stereomapper's
// StereoThread:
//   * the HSV colour map                StereoThread::run              stereothread.cpp:116-147
//   * back-projection + intensity gain  StereoThread::createCurrentMap stereothread.cpp:180-255
//   * fusion with the previous map      StereoThread::addDisparityMapToReconstruction  stereothread.cpp:290-437
// stereothread.cpp itself needs Qt and OpenCV and cannot be compiled here.  oracle/Makefile therefore
// cuts exactly those line ranges out of the reference file WHERE IT LIES into oracle/_ref/gen/*.inc
// (git-ignored build products, never committed) and this harness #includes them between minimal
// stand-ins for the members they touch.  The statements that run are the reference's own text,
// compiled with the reference's flags for this file (stereomapper.pro:145-150: -O0 -msse3), together
// with libviso2's Matrix (the class StereoThread uses, stereothread.h:8).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>
#include "matrix.h"              // libviso2/src/matrix.h through -I$(REFVISO)

using namespace std;             // stereothread.cpp:4

extern "C" void ref_colormap(const float* D1, int32_t d_width, int32_t d_height, float* out)
{
    struct { float* D1; } simg_obj = {const_cast<float*>(D1)};
    auto* _simg = &simg_obj;
    float* _D_color = 0;
#include "gen/colormap.inc"      // stereothread.cpp:116-147
    memcpy(out, _D_color, 3 * (size_t)d_width * d_height * sizeof(float));
    free(_D_color);
}

namespace {
struct View3D {                  // view3d.h:15-20
    struct point_3d {
        float x, y, z;
        float val;
        point_3d(float x, float y, float z, float val) : x(x), y(y), z(z), val(val) {}
    };
};
struct StereoThread {
    struct simage { unsigned char* I1; float* D1; int width, height, step; };
    struct map3d {               // stereothread.h:51-112 without the freeing destructor (see releaseMap below)
        float *I, *D, *X, *Y, *Z;
        Matrix H;
        int32_t width, height, idx;
        map3d() : I(0), D(0), X(0), Y(0), Z(0), width(0), height(0), idx(0) {}
    };
    simage* _simg;
    Matrix _H_total, _K;
    float _gain, _f, _cu, _cv, _base, _max_dist;
    map3d _previous_map3d;
    std::vector<std::vector<View3D::point_3d> > _points;
    map3d createCurrentMap();
    // The reference ends the fusion with `_previous_map3d = current_map3d; releaseMap(current_map3d);`
    // (stereothread.cpp:433-434): a shallow copy followed by freeing the copied buffers, so its next call reads
    // freed memory.  The defined semantics used throughout this repository is what the author evidently meant:
    // the previous map of a call IS the fused current map of the call before.  Here: releaseMap does nothing.
    void releaseMap(map3d&) {}
    void addDisparityMapToReconstruction();
};
#include "gen/create_current_map.inc"   // stereothread.cpp:180-255 (the whole member function)
#include "gen/fuse.inc"                 // stereothread.cpp:290-437 (the whole member function)
}  // namespace

// view = {f, cu, cv, base, max_dist, gain}; H = 3x4 row-major (rows 0..2 of the 4x4 pose).
// X/Y/Z are pre-filled with 0 where the reference leaves its malloc'ed arrays untouched.
extern "C" void ref_reproject(const uint8_t* I1, const float* D1, int32_t width, int32_t height, int32_t step,
                              const float* view, const double* H, float* I, float* D, float* X, float* Y, float* Z)
{
    StereoThread t;
    StereoThread::simage img = {const_cast<unsigned char*>(I1), const_cast<float*>(D1), width, height, step};
    t._simg = &img;
    t._H_total = Matrix::eye(4);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) t._H_total._val[r][c] = H[4 * r + c];
    t._f = view[0]; t._cu = view[1]; t._cv = view[2]; t._base = view[3]; t._max_dist = view[4]; t._gain = view[5];
    // the reference mallocs X/Y/Z and writes only reconstructable pixels: make "untouched" observable as 0
    StereoThread::map3d m = t.createCurrentMap();
    const size_t n = (size_t)width * height;
    memcpy(I, m.I, n * 4); memcpy(D, m.D, n * 4);
    for (size_t i = 0; i < n; i++) {
        const bool written = m.D[i] > 0;          // d>0 and z in range (out-of-range pixels were set to -1)
        X[i] = written ? m.X[i] : 0.f; Y[i] = written ? m.Y[i] : 0.f; Z[i] = written ? m.Z[i] : 0.f;
    }
    free(m.I); free(m.D); free(m.X); free(m.Y); free(m.Z);
}

// prev[5] = I, D, X, Y, Z of the previous map (the fused map of the call before; all null: there is none); prev D
// comes back with the merged points invalidated.  cur[5] receives the fused current map (X/Y/Z are meaningful where
// D > 0 only).  points_prev / points_curr: (x, y, z, val) quadruples in push_back order, capacity width*height each.
extern "C" void ref_fuse(const uint8_t* I1, const float* D1, int32_t width, int32_t height, int32_t step,
                         const float* view, const double* H, float* const* prev, float* const* cur,
                         float* points_prev, int32_t* n_prev, float* points_curr, int32_t* n_curr)
{
    StereoThread t;
    StereoThread::simage img = {const_cast<unsigned char*>(I1), const_cast<float*>(D1), width, height, step};
    t._simg = &img;
    t._H_total = Matrix::eye(4);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) t._H_total._val[r][c] = H[4 * r + c];
    t._f = view[0]; t._cu = view[1]; t._cv = view[2]; t._base = view[3]; t._max_dist = view[4]; t._gain = view[5];
    t._K = Matrix(3, 3);                      // StereoThread::getIntrinsics, stereothread.cpp:450-455
    t._K._val[0][0] = t._f; t._K._val[1][1] = t._f; t._K._val[0][2] = t._cu; t._K._val[1][2] = t._cv; t._K._val[2][2] = 1;
    const size_t n = (size_t)width * height;
    float* mine[5] = {0, 0, 0, 0, 0};
    if (prev && prev[0]) {
        for (int k = 0; k < 5; k++) { mine[k] = (float*)malloc(n * 4); memcpy(mine[k], prev[k], n * 4); }
        t._previous_map3d.I = mine[0]; t._previous_map3d.D = mine[1]; t._previous_map3d.X = mine[2];
        t._previous_map3d.Y = mine[3]; t._previous_map3d.Z = mine[4];
        t._previous_map3d.width = width; t._previous_map3d.height = height;
        t._points.push_back(std::vector<View3D::point_3d>());     // the previous call's current points
    }
    t.addDisparityMapToReconstruction();
    if (mine[0]) memcpy(prev[1], mine[1], n * 4);
    for (int k = 0; k < 5; k++) free(mine[k]);
    const StereoThread::map3d& m = t._previous_map3d;             // = the fused current map (releaseMap is a no-op)
    float* src[5] = {m.I, m.D, m.X, m.Y, m.Z};
    for (int k = 0; k < 5; k++) { memcpy(cur[k], src[k], n * 4); free(src[k]); }
    *n_prev = 0;
    if (t._points.size() == 2) {
        *n_prev = (int32_t)t._points[0].size();
        memcpy(points_prev, t._points[0].data(), t._points[0].size() * 16);
    }
    *n_curr = (int32_t)t._points.back().size();
    memcpy(points_curr, t._points.back().data(), t._points.back().size() * 16);
}
