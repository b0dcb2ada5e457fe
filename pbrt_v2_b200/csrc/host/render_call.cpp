// render_call.cpp -- GpuSamplerRenderer::Render: the ONE call that crosses the
// host<->device boundary (replaces SamplerRenderer::Render,
// renderers/samplerrenderer.cpp:188-222).  There is no CPU fallback: if the CUDA
// library reports an error the render fails loudly.
#include "scene.h"

#include <cstdio>

namespace pb {

bool GpuSamplerRenderer::Render(Scene *scene) {
    PbrtGpuScene *gs = nullptr;
    int rc = pbrtgpu_scene_create(&scene->desc, &gs);
    if (rc != PBRTGPU_OK) {
        fprintf(stderr, "Error: pbrtgpu_scene_create failed (%d): %s\n", rc, pbrtgpu_last_error());
        return false;
    }
    PbrtGpuRenderDesc rd;
    FillRenderDesc(&rd);
    // several GPUs: every rank renders its share of the task list, the films are summed onto rank 0 by the
    // library's NCCL reduce, rank 0 owns the image
    if (comm) rc = pbrtgpu_render_sharded(gs, &rd, comm, film->pixels.data(), &lastStats);
    else rc = pbrtgpu_render(gs, &rd, film->pixels.data(), &lastStats);
    pbrtgpu_scene_destroy(gs);
    if (rc != PBRTGPU_OK) {
        fprintf(stderr, "Error: pbrtgpu_render failed (%d): %s\n", rc, pbrtgpu_last_error());
        return false;
    }
    if (writeImage && (!comm || commRank == 0)) return film->WriteImage();
    return true;
}

}  // namespace pb
