// TEST INFRASTRUCTURE -- not part of the product.
//
// Pins the oracle's restatement of the two per-pixel loops that consume D1 inside stereomapper's
// StereoThread:
//   * the HSV colour map                StereoThread::run              stereothread.cpp:116-147
//   * back-projection + intensity gain  StereoThread::createCurrentMap stereothread.cpp:180-255
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
struct StereoThread {
    struct simage { unsigned char* I1; float* D1; int width, height, step; };
    struct map3d {
        float *I, *D, *X, *Y, *Z;
        Matrix H;
        int32_t width, height, idx;
    };
    simage* _simg;
    Matrix _H_total;
    float _gain, _f, _cu, _cv, _base, _max_dist;
    map3d createCurrentMap();
};
#include "gen/create_current_map.inc"   // stereothread.cpp:180-255 (the whole member function)
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
