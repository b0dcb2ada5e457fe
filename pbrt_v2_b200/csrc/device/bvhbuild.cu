// bvhbuild.cu -- C ABI of the device BVH construction (include/pbrtgpu.h: pbrtgpu_bvh_build); the kernels and the
// level loop are in dbvhbuild.cuh.  Like the rest of libpbrtgpu there is no CPU fallback: without a CUDA device the
// call fails with PBRTGPU_ENODEVICE.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <string>

namespace {
thread_local cudaError_t g_bvhbErr = cudaSuccess;
inline bool BvhbCheck(cudaError_t e) {
    if (e != cudaSuccess && g_bvhbErr == cudaSuccess) g_bvhbErr = e;
    return e == cudaSuccess;
}
inline bool BvhbCopy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return BvhbCheck(cudaMemcpyAsync(dst, src, bytes, kind, st)) && BvhbCheck(cudaStreamSynchronize(st));
}
inline unsigned BvhbBlocks(size_t work) {   // grid-stride kernels: enough blocks for the work, at most 8 per SM of a B200
    const size_t want = (work + 255) / 256, cap = 148 * 8;
    return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
}
}  // namespace

#define BVHB_ALLOC(pp, bytes) BvhbCheck(cudaMalloc((pp), (bytes)))
#define BVHB_FREE(p) cudaFree(p)
#define BVHB_H2D(dst, src, bytes, stream) BvhbCopy((dst), (src), (bytes), cudaMemcpyHostToDevice, (stream))
#define BVHB_D2H(dst, src, bytes, stream) BvhbCopy((dst), (src), (bytes), cudaMemcpyDeviceToHost, (stream))
#define BVHB_LAUNCH(kernel, work, stream, ...) \
    do { kernel<<<BvhbBlocks(work), 256, 0, static_cast<cudaStream_t>(stream)>>>(__VA_ARGS__); BvhbCheck(cudaGetLastError()); } while (0)

#include "dbvhbuild.cuh"

extern "C" int pbrtgpu_set_error_text(int code, const char *text);   // pbrtgpu.cu

extern "C" int pbrtgpu_bvh_build(const float *bounds, uint32_t nPrimitives, uint32_t maxPrimsInNode, int splitMethod,
                                 PbrtGpuBVHNode *nodes, uint32_t *nNodes, uint32_t *order, PbrtGpuBuildStats *stats) {
    if (!nNodes || (nPrimitives && (!bounds || !nodes || !order)))
        return pbrtgpu_set_error_text(PBRTGPU_EINVAL, "pbrtgpu_bvh_build: null argument");
    *nNodes = 0;
    if (stats) memset(stats, 0, sizeof(*stats));
    if (splitMethod != PBRTGPU_SPLIT_SAH)
        return pbrtgpu_set_error_text(PBRTGPU_EUNSUPPORTED,
                                      "pbrtgpu_bvh_build: only splitmethod \"sah\" is built on the device (\"middle\" and \"equal\" "
                                      "run std::nth_element over whole ranges)");
    if (maxPrimsInNode < 1 || maxPrimsInNode > 255)
        return pbrtgpu_set_error_text(PBRTGPU_EINVAL, "pbrtgpu_bvh_build: maxPrimsInNode must be 1..255 (BVHAccel clamps to 255, bvh.cpp:155)");
    if (nPrimitives > 300000000u)   // 32-bit indexing of the per-node key arrays (6 keys x 2n nodes)
        return pbrtgpu_set_error_text(PBRTGPU_EINVAL, "pbrtgpu_bvh_build: more than 3e8 primitives");
    for (size_t i = 0; i < (size_t)nPrimitives * 6; ++i)
        if (!std::isfinite(bounds[i]))   // (an infinite extent makes the reference's bucket index int(NaN): undefined there too)
            return pbrtgpu_set_error_text(PBRTGPU_EINVAL, "pbrtgpu_bvh_build: a primitive's bounds are not finite");
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
        return pbrtgpu_set_error_text(PBRTGPU_ENODEVICE, "no CUDA device visible; libpbrtgpu has no CPU fallback");
    g_bvhbErr = cudaSuccess;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (stats) { BvhbCheck(cudaEventCreate(&e0)); BvhbCheck(cudaEventCreate(&e1)); BvhbCheck(cudaEventRecord(e0, nullptr)); }
    uint32_t nLevels = 0;
    const int rc = bvhb::Build(bounds, nPrimitives, maxPrimsInNode, nodes, nNodes, order, &nLevels, nullptr);
    if (stats && g_bvhbErr == cudaSuccess) {
        float ms = 0.f;
        BvhbCheck(cudaEventRecord(e1, nullptr));
        BvhbCheck(cudaEventSynchronize(e1));
        BvhbCheck(cudaEventElapsedTime(&ms, e0, e1));
        stats->buildMs = ms;
        stats->nLevels = nLevels;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    BvhbCheck(cudaDeviceSynchronize());
    if (rc != 0 || g_bvhbErr != cudaSuccess) {
        *nNodes = 0;
        const std::string msg = std::string("pbrtgpu_bvh_build: ") + cudaGetErrorString(g_bvhbErr == cudaSuccess ? cudaErrorUnknown : g_bvhbErr);
        return pbrtgpu_set_error_text(PBRTGPU_ECUDA, msg.c_str());
    }
    return PBRTGPU_OK;
}
