// ASan / UBSan driver for the scene front end (csrc/host: parser, shapes, Loop subdivision, BVH build, flattening): loads every
// scene file named on the command line.  The front end may refuse a file; it may not touch memory it does not own, overflow
// an integer or run away.  The device library is replaced by stubs that fail: nothing here needs a GPU.
// scripts/sanitize_parser.py feeds it random in-scope scenes (scripts/fuzz_scenes.py's generator) and mutations of them.
#include "../../../pbrt_v2_b200/csrc/host/scene.h"

#include <cstdio>

extern "C" {
const char *pbrtgpu_last_error(void) { return "stub"; }
int pbrtgpu_scene_create(const PbrtGpuSceneDesc *, PbrtGpuScene **) { return PBRTGPU_EUNSUPPORTED; }
void pbrtgpu_scene_destroy(PbrtGpuScene *) {}
int pbrtgpu_render(PbrtGpuScene *, const PbrtGpuRenderDesc *, PbrtGpuPixel *, PbrtGpuStats *) { return PBRTGPU_EUNSUPPORTED; }
int pbrtgpu_render_sharded(PbrtGpuScene *, const PbrtGpuRenderDesc *, PbrtGpuComm *, PbrtGpuPixel *, PbrtGpuStats *) { return PBRTGPU_EUNSUPPORTED; }
int pbrtgpu_bvh_build(const float *, uint32_t, uint32_t, int, PbrtGpuBVHNode *, uint32_t *, uint32_t *, PbrtGpuBuildStats *) { return PBRTGPU_EUNSUPPORTED; }
int pbrtgpu_loopsubdiv_counts(const PbrtGpuSubdivMesh *, int, uint32_t *, uint32_t *) { return PBRTGPU_EUNSUPPORTED; }
int pbrtgpu_loopsubdiv_refine(const PbrtGpuSubdivMesh *, int, float *, float *, int32_t *) { return PBRTGPU_EUNSUPPORTED; }
}

int main(int argc, char **argv) {
    int ok = 0, refused = 0;
    for (int i = 1; i < argc; ++i) {
        pb::Options opt;
        opt.nCores = 2;
        opt.quiet = true;
        pb::LoadedScene ls;
        if (pb::LoadSceneFile(argv[i], opt, &ls)) {
            PbrtGpuRenderDesc rd;
            ls.renderer->FillRenderDesc(&rd);
            ++ok;
        } else
            ++refused;
    }
    printf("loaded %d, refused %d\n", ok, refused);
    return 0;
}
