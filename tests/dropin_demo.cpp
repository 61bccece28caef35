// Headless stand-in for the reference's two call sites, compiled against the DROP-IN elas.h:
//   mode "stereomapper": StereoThread::run's parameter set and call (stereothread.cpp:76-114)
//   mode "demo":         libelas/src/main.cpp:61-64
// Reads two raw uint8 images (width x height, tightly packed), writes D1 and D2 as raw float32.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "elas.h"

static std::vector<uint8_t> slurp(const char* path, size_t n)
{
    std::vector<uint8_t> v(n);
    FILE* f = fopen(path, "rb");
    if (!f || fread(v.data(), 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return v;
}

int main(int argc, char** argv)
{
    if (argc != 9) { fprintf(stderr, "usage: %s mode W H dmax left.raw right.raw D1.out D2.out\n", argv[0]); return 2; }
    const bool stereomapper = !strcmp(argv[1], "stereomapper");
    const int32_t width = atoi(argv[2]), height = atoi(argv[3]);
    std::vector<uint8_t> I1 = slurp(argv[5], (size_t)width * height), I2 = slurp(argv[6], (size_t)width * height);

    Elas::parameters param(Elas::ROBOTICS);
    if (stereomapper) {                       // stereothread.cpp:76-80
        param.postprocess_only_left = true;
        param.filter_adaptive_mean = true;
        param.support_texture = 30;
    } else {                                  // main.cpp:61-62
        param.postprocess_only_left = false;
    }
    param.disp_max = atoi(argv[4]);

    std::vector<float> D1((size_t)width * height, -77.f), D2((size_t)width * height, -77.f);
    const int32_t dims[3] = {width, height, width};
    Elas elas(param);
    elas.process(I1.data(), I2.data(), D1.data(), D2.data(), dims);

    FILE* f = fopen(argv[7], "wb"); fwrite(D1.data(), 4, D1.size(), f); fclose(f);
    f = fopen(argv[8], "wb"); fwrite(D2.data(), 4, D2.size(), f); fclose(f);
    return 0;
}
