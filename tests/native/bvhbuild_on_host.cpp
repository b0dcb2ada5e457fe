// bvhbuild_on_host.cpp -- TEST INFRASTRUCTURE.  The device BVH construction (device/dbvhbuild.cuh: every kernel and the
// level loop that drives them) executed on the CPU: a launch becomes a loop over the grid's threads, device memory is
// host memory, the integer atomics are plain read-modify-writes (the threads of a launch run one after the other here,
// and the kernels are written so that their results do not depend on the order in which a launch's threads run).
// The CPU test stage compares the result with the host front end's recursive build (libpbrthost: pbrthost_bvh_build),
// node for node.  Never linked into the product.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define DOH_EXPORT extern "C" __attribute__((visibility("default")))
#define BVHB_HOST_EMULATION 1

static uint3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { const uint32_t o = *p; *p = o + v; return o; }
static inline uint32_t atomicMin(uint32_t *p, uint32_t v) { const uint32_t o = *p; if (v < o) *p = v; return o; }
static inline uint32_t atomicMax(uint32_t *p, uint32_t v) { const uint32_t o = *p; if (v > o) *p = v; return o; }

// the emulated grid: DOH_GRID_THREADS threads of 256 per block (the product uses up to 148 x 8 blocks; the results do not
// depend on the grid size, which the tests check by building with several)
static unsigned g_maxBlocks = 148 * 8;
static bool g_reverse = false;   // run the threads of every launch in descending order instead
static inline bool EmuAlloc(void **pp, size_t bytes) { *pp = malloc(bytes); return *pp != nullptr; }
static inline bool EmuCopy(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); return true; }
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
#define BVHB_ALLOC(pp, bytes) EmuAlloc((pp), (bytes))
#define BVHB_FREE(p) free(p)
#define BVHB_H2D(dst, src, bytes, stream) EmuCopy((dst), (src), (bytes))
#define BVHB_D2H(dst, src, bytes, stream) EmuCopy((dst), (src), (bytes))
#define BVHB_LAUNCH(kernel, work, stream, ...) EmuLaunch(kernel, (work), __VA_ARGS__)

#include "../../pbrt_v2_b200/csrc/device/dbvhbuild.cuh"

// pbrtgpu_bvh_build's arguments; maxBlocks / reverse choose the emulated schedule
DOH_EXPORT int doh_bvh_build(const float *bounds, uint32_t n, uint32_t maxPrimsInNode, PbrtGpuBVHNode *nodes, uint32_t *nNodes,
                             uint32_t *order, uint32_t *nLevels, unsigned maxBlocks, int reverse) {
    g_maxBlocks = maxBlocks ? maxBlocks : 148 * 8;
    g_reverse = reverse != 0;
    return bvhb::Build(bounds, n, maxPrimsInNode, nodes, nNodes, order, nLevels, nullptr);
}
