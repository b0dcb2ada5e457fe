// main.cpp -- `pbrt_b200 [--ncores n] [--outfile f] [--quiet] scene.pbrt`:
// command-line twin of the reference's src/main/pbrt.cpp:41-82 for scenes whose
// Accelerator / SurfaceIntegrator / Renderer select the GPU hot path.
#include "scene.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

int main(int argc, char *argv[]) {
    pb::Options opt;
    std::vector<std::string> files;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--ncores")) opt.nCores = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--outfile")) opt.imageFile = argv[++i];
        else if (!strcmp(argv[i], "--quiet")) opt.quiet = true;
        else if (!strcmp(argv[i], "--verbose")) {}
        else if (!strcmp(argv[i], "--help") || !strcmp(argv[i], "-h")) {
            printf("usage: pbrt_b200 [--ncores n] [--outfile filename] [--quiet] <filename.pbrt> ...\n");
            return 0;
        } else files.push_back(argv[i]);
    }
    if (files.empty()) { fprintf(stderr, "pbrt_b200: no scene file given\n"); return 1; }
    for (const std::string &f : files) {
        pb::LoadedScene ls;
        if (!pb::LoadSceneFile(f, opt, &ls)) {
            fprintf(stderr, "Error: %s\n", ls.error.c_str());
            return 1;
        }
        if (!ls.renderer->Render(ls.scene.get())) return 1;
        if (!opt.quiet) {
            const PbrtGpuStats &s = ls.renderer->lastStats;
            printf("Rendering: %.3f s on device (%.2f Msamples/s), %llu camera samples, %llu+%llu rays\n",
                   s.renderMs * 1e-3, s.cameraSamples / (s.renderMs * 1e3), (unsigned long long)s.cameraSamples,
                   (unsigned long long)s.closestRays, (unsigned long long)s.shadowRays);
        }
    }
    return 0;
}
