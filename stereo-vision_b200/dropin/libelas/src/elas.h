// Drop-in replacement for libelas/src/elas.h of willSapgreen/stereo-vision.
//
// Same public interface as the reference header (class Elas, enum setting, struct parameters with the
// same 23 fields, defaults and presets, Elas(parameters), process(I1,I2,D1,D2,dims)) so that
//   stereomapper/stereothread.h:12    #include "../libelas/src/elas.h"
//   stereomapper/stereothread.cpp:76-114, libelas/src/main.cpp:61-64, libelas/matlab/elasMex.cpp:55-108
// compile unchanged.  The work is done by the B200 library libelas_b200.so through its C ABI
// (include/elas_b200.h); this class holds nothing but the parameter block.
//
// Copy this directory over libelas/src (elas.h, elas.cpp, descriptor.h, descriptor.cpp); see
// INTEGRATION.md.  There is no CPU fallback: process() reports the error and leaves D1/D2 filled
// with -10 when the CUDA library or device is unavailable.
#ifndef __ELAS_H__
#define __ELAS_H__

#include <stdint.h>

#ifdef PROFILE
#include "timer.h"   // the reference adds a Timer member under -DPROFILE (elas.h:48-50,317-319); kept for layout parity
#endif

class Elas {
public:
    enum setting { ROBOTICS, MIDDLEBURY };

    struct parameters {
        int32_t disp_min;               // min disparity
        int32_t disp_max;               // max disparity
        float   support_threshold;      // max. uniqueness ratio (best vs. second best support match)
        int32_t support_texture;        // min texture for support points
        int32_t candidate_stepsize;     // step size of regular grid on which support points are matched
        int32_t incon_window_size;      // window size of inconsistent support point check
        int32_t incon_threshold;        // disparity similarity threshold for support point to be considered consistent
        int32_t incon_min_support;      // minimum number of consistent support points
        bool    add_corners;            // add support points at image corners with nearest neighbor disparities
        int32_t grid_size;              // size of neighborhood for additional support point extrapolation
        float   beta;                   // image likelihood parameter
        float   gamma;                  // prior constant
        float   sigma;                  // prior sigma
        float   sradius;                // prior sigma radius
        int32_t match_texture;          // min texture for dense matching
        int32_t lr_threshold;           // disparity threshold for left/right consistency check
        float   speckle_sim_threshold;  // similarity threshold for speckle segmentation
        int32_t speckle_size;           // maximal size of a speckle (small speckles get removed)
        int32_t ipol_gap_width;         // interpolate small gaps (left<->right, top<->bottom)
        bool    filter_median;          // optional median filter (approximated)
        bool    filter_adaptive_mean;   // optional adaptive mean filter (approximated)
        bool    postprocess_only_left;  // saves time by not postprocessing the right image
        bool    subsampling;            // only compute disparities for each 2nd pixel; D1/D2 are then
                                        // width/2 x height/2 (rounded towards zero)

        // presets are filled in by the library (elas_b200_default_params), values as elas.h:93-146
        parameters(setting s = ROBOTICS);
    };

    Elas(parameters param) : _param(param) {}
    ~Elas() {}

    // inputs:  I1, I2  left / right intensity image (uint8), dims[2] bytes per line
    // outputs: D1, D2  left / right disparity image (float, bytes per line = width), caller-allocated
    //          dims[0] = width, dims[1] = height, dims[2] = bytes per line of I1 and I2
    void process(uint8_t* I1, uint8_t* I2, float* D1, float* D2, const int32_t* dims);

private:
    parameters _param;
#ifdef PROFILE
    Timer timer;
#endif
};

#endif
