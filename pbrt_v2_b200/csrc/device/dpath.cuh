// dpath.cuh -- what the path state machine (path_body.inc) shares with everything that runs it: the kernels of
// pbrtgpu.cu and the host-compiled test harness (tests/native/device_on_host.cpp).
#pragma once
#include "dmath.cuh"
#include "drng.cuh"

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

// samplerrenderer.cpp:119-133: NaN / negative / infinite luminance -> black
DEV Spec SanitizeRadiance(const Spec &L) {
    if (L.HasNaNs()) return Spec(0.f);
    if ((double)L.y() < -1e-5) return Spec(0.f);
    if (isinf(L.y())) return Spec(0.f);
    return L;
}

}  // namespace dev
