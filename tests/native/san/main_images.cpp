// ASan / UBSan driver for the host front end's file readers (csrc/host/envmap.cpp ReadImagePFM, tga.cpp, exr.cpp, measured.cpp) and the
// InfiniteAreaLight constructor behind them: reads every file named on the command line; a reader may refuse a file, it may
// not touch memory it does not own.  scripts/sanitize_images.py feeds it valid files and thousands of mutations of them.
#include "../../../pbrt_v2_b200/csrc/host/scene.h"

#include <cstdio>

int main(int argc, char **argv) {
    int ok = 0, refused = 0;
    for (int i = 1; i < argc; ++i) {
        int w = 0, h = 0;
        std::vector<float> rgb;
        std::string err;
        const std::string name = argv[i];
        if (name.size() > 5 && name.substr(name.size() - 5) == ".brdf") {   // measured BRDF data (materials/measured.cpp:85-135)
            pb::MeasuredData md;
            if (pb::LoadIrregIsotropicBRDF(name, &md, &err)) ++ok; else ++refused;
            continue;
        }
        if (!pb::ReadImage(argv[i], &w, &h, &rgb, &err)) { ++refused; continue; }
        if ((size_t)w * h * 3 != rgb.size()) { fprintf(stderr, "%s: size mismatch\n", argv[i]); return 2; }
        if ((long long)w * h <= 4096) {   // the constructor as well (resampling, pyramid, distribution)
            pb::EnvMap em;
            const float L[3] = {1.f, .5f, 2.f};
            pb::BuildEnvMap(L, w, h, rgb.data(), &em);
        }
        ++ok;
    }
    printf("read %d, refused %d\n", ok, refused);
    return 0;
}
