// dpath.cuh -- what the path state machine (path_body.inc) shares with everything that runs it: the kernels of
// pbrtgpu.cu and the host-compiled test harness (tests/native/device_on_host.cpp).
#pragma once
#include "dmath.cuh"
#include "drng.cuh"

// pixel filters other than the 0.5 box (filters/*.cpp through ImageFilm's 16x16 weight table).  Off in
// the shipped kernels (the code below never ran on a GPU: the round's GPU budget ended first); the host-compiled harness
// builds it with -DPBRT_FILTER=1 and compares whole films with the oracle.
#ifndef PBRT_FILTER
#define PBRT_FILTER 0
#endif

namespace dev {

struct RenderParams {
    float rasterToCamera[16];
    float cameraToWorld[16];
    float lensRadius, focalDistance;
    float shutterOpen, shutterClose;
    float filterXWidth, filterYWidth;
    int xPixelStart, yPixelStart, xPixelCount, yPixelCount;
    int spp, maxDepth;
    int integrator, lightStrategy;   // PBRTGPU_INTEGRATOR_*, PBRTGPU_STRATEGY_*
    SampleLayoutDev lay;
#if PBRT_FILTER
    const float *filterTable;   // ImageFilm::filterTable (16 x 16 weights) in device memory; NULL: the 0.5 box (FilmAdd)
#endif
};

enum PathState {
    S_NEW_SAMPLE = 0, S_HIT_RESULT, S_SHADOW_RESULT, S_AFTER_LIGHT, S_MIS_RESULT, S_AFTER_DIRECT, S_END_SAMPLE, S_DONE,
    S_E_LIGHT, S_E_MIS, S_E_PATH,   // waiting for the BSDF value at the evaluation site
    S_DL_NEXT,                      // DirectLightingIntegrator: next (light, sample) of the light loop
    S_SPEC_NEXT                     // ... next of SpecularReflect / SpecularTransmit at this node (PBRT_SPECULAR builds)
};
// PBRT_SPECULAR builds: one frame per ancestor of the node being shaded in DirectLightingIntegrator's recursion
// (integrator.cpp:177-258): the parent's radiance so far and what the child's result is multiplied with
constexpr int DL_MAX_DEPTH = 8;
struct DlFrame {
    Spec L, f;
    float absDot, pdf;
    int stage;        // 0: the SpecularReflect child is in flight, 1: the SpecularTransmit child
    V3 p, nn, ng, sn, tn, wo;
    int material, areaLight;
    float rayEpsilon;
};
// ls[s].w bit field
enum { LS_TAIL = 1, LS_LEFT = 2, LS_RIGHT = 4, LS_UP = 8, LS_DOWN = 16,
       LS_HIT = 32,     // directlighting: the camera ray hit something (6 RNG draws follow, see k_film_direct)
       LS_SPEC = 64 };  // PBRT_SPECULAR builds: the parked vertex was reached by a specular bounce (path.cpp:62-63)

// Which pixels does ImageFilm::AddSample (film/image.cpp:77-96) touch for this sample?
// With the 0.5 box filter: the sample's own pixel, plus a neighbour when the
// sample sits exactly on a pixel edge or xPos + u rounded up to the next integer.
DEV int FilmFootprint(const RenderParams &rp, int ownX, int ownY, float imageX, float imageY) {
    float dimageX = imageX - 0.5f;
    float dimageY = imageY - 0.5f;
    int x0 = Ceil2Int(dimageX - rp.filterXWidth);
    int x1 = Floor2Int(dimageX + rp.filterXWidth);
    int y0 = Ceil2Int(dimageY - rp.filterYWidth);
    int y1 = Floor2Int(dimageY + rp.filterYWidth);
    x0 = maxi(x0, rp.xPixelStart);
    x1 = mini(x1, rp.xPixelStart + rp.xPixelCount - 1);
    y0 = maxi(y0, rp.yPixelStart);
    y1 = mini(y1, rp.yPixelStart + rp.yPixelCount - 1);
    int code = 0;
    if (x0 < ownX) code |= LS_LEFT;
    if (x1 > ownX) code |= LS_RIGHT;
    if (y0 < ownY) code |= LS_UP;
    if (y1 > ownY) code |= LS_DOWN;
    return code;
}

#if PBRT_FILTER
// ImageFilm::AddSample (film/image.cpp:77-137) for any filter table and width.  Every pixel of the footprint is another
// task's pixel as likely as not: all adds are atomic, and their order is as undefined as in the reference, which takes its
// AtomicAdd branch (image.cpp:128-134) for exactly these filters.  In one thread (the host harness, tasks in order) the
// order is the reference's --ncores 1 order and the film is bit-identical.
DEV void FilmAddFiltered(const RenderParams &rp, PbrtGpuPixel *film, float imageX, float imageY, const Spec &L) {
    const float dimageX = imageX - 0.5f;
    const float dimageY = imageY - 0.5f;
    int x0 = Ceil2Int(dimageX - rp.filterXWidth);
    int x1 = Floor2Int(dimageX + rp.filterXWidth);
    int y0 = Ceil2Int(dimageY - rp.filterYWidth);
    int y1 = Floor2Int(dimageY + rp.filterYWidth);
    x0 = maxi(x0, rp.xPixelStart);
    x1 = mini(x1, rp.xPixelStart + rp.xPixelCount - 1);
    y0 = maxi(y0, rp.yPixelStart);
    y1 = mini(y1, rp.yPixelStart + rp.yPixelCount - 1);
    if ((x1 - x0) < 0 || (y1 - y0) < 0) return;
    if (x1 - x0 >= 64 || y1 - y0 >= 64) return;   // the oracle's limit (its offset arrays hold 64 entries)
    float xyz[3];   // RGBToXYZ, core/spectrum.h:58-62
    xyz[0] = 0.412453f * L.r + 0.357580f * L.g + 0.180423f * L.b;
    xyz[1] = 0.212671f * L.r + 0.715160f * L.g + 0.072169f * L.b;
    xyz[2] = 0.019334f * L.r + 0.119193f * L.g + 0.950227f * L.b;
    const float invXWidth = 1.f / rp.filterXWidth, invYWidth = 1.f / rp.filterYWidth;   // filter.h:47-49
    for (int y = y0; y <= y1; ++y) {
        const float fy = fabsf((y - dimageY) * invYWidth * 16);
        const int ify = mini(Floor2Int(fy), 16 - 1);
        for (int x = x0; x <= x1; ++x) {
            const float fx = fabsf((x - dimageX) * invXWidth * 16);
            const int ifx = mini(Floor2Int(fx), 16 - 1);
            const float filterWt = rp.filterTable[ify * 16 + ifx];
            PbrtGpuPixel *px = film + (size_t)(y - rp.yPixelStart) * rp.xPixelCount + (x - rp.xPixelStart);
            atomicAdd(&px->Lxyz[0], filterWt * xyz[0]);
            atomicAdd(&px->Lxyz[1], filterWt * xyz[1]);
            atomicAdd(&px->Lxyz[2], filterWt * xyz[2]);
            atomicAdd(&px->weightSum, filterWt);
        }
    }
}
// The same, for sample `row` of a task-pixel whose radiance waited in ls[]: the image position again from the pixel's
// LD tables (LDPixelSample's first two dimensions, as CameraRay computes them)
DEV void FilmAddFilteredRow(const RenderParams &rp, PbrtGpuPixel *film, const uint16_t *row, const uint32_t *scr, int xPos,
                            int yPos, const Spec &L) {
    const uint32_t pi = row[0];
    FilmAddFiltered(rp, film, xPos + VanDerCorput(pi, scr[0]), yPos + Sobol2(pi, scr[1]), L);
}
// u16 entries per table row of the render's integrator (PathIntegrator: fixed layout; directlighting: per scene)
DEV int FilmRowStride(const RenderParams &rp) {
    return rp.integrator == PBRTGPU_INTEGRATOR_DIRECTLIGHTING ? rp.lay.rowStride : PERM_STRIDE;
}
#endif

// samplerrenderer.cpp:119-133: NaN / negative / infinite luminance -> black
DEV Spec SanitizeRadiance(const Spec &L) {
    if (L.HasNaNs()) return Spec(0.f);
    if ((double)L.y() < -1e-5) return Spec(0.f);
    if (isinf(L.y())) return Spec(0.f);
    return L;
}

}  // namespace dev
