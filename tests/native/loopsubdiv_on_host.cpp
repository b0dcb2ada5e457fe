// loopsubdiv_on_host.cpp -- TEST INFRASTRUCTURE.  LoopSubdiv::Refine's device version (device/dloopsubdiv.cuh: every kernel
// and the level loop) executed on the CPU: launches as loops over the grid's threads, device memory as host memory.  The
// CPU test stage compares the refined mesh with the host front end's LoopSubdiv::Refine (libpbrthost:
// pbrthost_loopsubdiv_refine), vertex for vertex.  Never linked into the product.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define DOH_EXPORT extern "C" __attribute__((visibility("default")))
#define LSD_HOST_EMULATION 1
#define PBRT_NI_LEVEL 0
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
using std::isnan;
using std::isinf;

static uint3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }

static unsigned g_maxBlocks = 148 * 8;
static bool g_reverse = false;
static inline bool EmuAlloc(void **pp, size_t bytes) { *pp = malloc(bytes); return *pp != nullptr; }
static inline bool EmuCopy(void *dst, const void *src, size_t bytes) { if (bytes) memcpy(dst, src, bytes); return true; }
template <typename K, typename... A> static void EmuLaunch(K kernel, size_t work, A... args) {
    size_t blocks = (work + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > g_maxBlocks) blocks = g_maxBlocks;
    blockDim = dim3(256, 1, 1);
    gridDim = dim3((unsigned)blocks, 1, 1);
    const size_t total = blocks * 256;
    for (size_t k = 0; k < total; ++k) {
        const size_t t = g_reverse ? total - 1 - k : k;
        blockIdx.x = (unsigned)(t / 256);
        threadIdx.x = (unsigned)(t % 256);
        kernel(args...);
    }
}
#define LSD_ALLOC(pp, bytes) EmuAlloc((pp), (bytes))
#define LSD_FREE(p) free(p)
#define LSD_H2D(dst, src, bytes, stream) EmuCopy((dst), (src), (bytes))
#define LSD_D2H(dst, src, bytes, stream) EmuCopy((dst), (src), (bytes))
#define LSD_LAUNCH(kernel, work, stream, ...) EmuLaunch(kernel, (work), __VA_ARGS__)

#include "../../pbrt_v2_b200/csrc/device/dloopsubdiv.cuh"

// pbrtgpu_loopsubdiv_refine's arguments (the caller sizes the outputs with pbrtgpu_loopsubdiv_counts, which needs no
// device); maxBlocks / reverse choose the emulated schedule.  Returns lsd::Refine's code.
DOH_EXPORT int doh_loopsubdiv_refine(const PbrtGpuSubdivMesh *mesh, int nLevels, float *P, float *N, int32_t *indices, unsigned maxBlocks,
                                     int reverse) {
    g_maxBlocks = maxBlocks ? maxBlocks : 148 * 8;
    g_reverse = reverse != 0;
    return lsd::Refine(*mesh, nLevels, P, N, indices, nullptr);
}
