// CPU test harness of stereo-vision_b200/csrc/host_widen.h: the widening of the narrowed right map on the host.
#include "../../stereo-vision_b200/csrc/host_widen.h"

extern "C" void test_widen_i16(const int16_t* src, float* dst, size_t n) { elasb::widen_i16_to_f32(src, dst, n); }
extern "C" void test_widen_u8_mask(const uint8_t* vals, const uint32_t* mask, int words_per_row, float* dst, int Dw, int Dh)
{
    elasb::widen_u8_mask_to_f32(vals, mask, words_per_row, dst, Dw, Dh);
}
