// loopsubdiv.cu -- C ABI of LoopSubdiv::Refine on the device (include/pbrtgpu.h: pbrtgpu_loopsubdiv_*); the kernels and
// the level loop are in dloopsubdiv.cuh.  No CPU fallback: without a CUDA device the call fails with PBRTGPU_ENODEVICE.
#define PBRT_NI_LEVEL 0   // dmath.cuh: every helper inline here (its out-of-line ones belong to pbrtgpu.cu's translation unit)
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <string>

namespace {
thread_local cudaError_t g_lsdErr = cudaSuccess;
inline bool LsdCheck(cudaError_t e) {
    if (e != cudaSuccess && g_lsdErr == cudaSuccess) g_lsdErr = e;
    return e == cudaSuccess;
}
inline bool LsdCopy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, void *stream) {
    if (bytes == 0) return true;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return LsdCheck(cudaMemcpyAsync(dst, src, bytes, kind, st)) && LsdCheck(cudaStreamSynchronize(st));
}
inline unsigned LsdBlocks(size_t work) {
    const size_t want = (work + 255) / 256, cap = 148 * 8;
    return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
}
}  // namespace

#define LSD_ALLOC(pp, bytes) LsdCheck(cudaMalloc((pp), (bytes)))
#define LSD_FREE(p) cudaFree(p)
#define LSD_H2D(dst, src, bytes, stream) LsdCopy((dst), (src), (bytes), cudaMemcpyHostToDevice, (stream))
#define LSD_D2H(dst, src, bytes, stream) LsdCopy((dst), (src), (bytes), cudaMemcpyDeviceToHost, (stream))
#define LSD_LAUNCH(kernel, work, stream, ...) \
    do { kernel<<<LsdBlocks(work), 256, 0, static_cast<cudaStream_t>(stream)>>>(__VA_ARGS__); LsdCheck(cudaGetLastError()); } while (0)

#include "dloopsubdiv.cuh"

extern "C" int pbrtgpu_set_error_text(int code, const char *text);   // pbrtgpu.cu

namespace {
// what the kernels rely on: indices in range, reciprocal neighbour links across a shared edge, no edge in three faces
const char *ValidateSubdivMesh(const PbrtGpuSubdivMesh &m, int nLevels) {
    if (nLevels < 0 || nLevels > 12) return "nLevels out of range";
    if (m.nFaces == 0 || m.nVertices == 0) return "empty control mesh";
    if (!m.P || !m.startFace || !m.vflags || !m.faceV || !m.faceF) return "null array";
    const int64_t nv = m.nVertices, nf = m.nFaces;
    for (int64_t i = 0; i < 3 * nv; ++i) if (!std::isfinite(m.P[i])) return "a control vertex is not finite";
    std::vector<uint64_t> keys;
    keys.reserve(3 * (size_t)nf);
    for (int64_t j = 0; j < nf; ++j)
        for (int k = 0; k < 3; ++k) {
            const int64_t a = m.faceV[3 * j + k], b = m.faceV[3 * j + (k + 1) % 3], f2 = m.faceF[3 * j + k];
            if (a < 0 || a >= nv || b < 0 || b >= nv || a == b) return "a face refers to a bad vertex index or repeats one";
            if (f2 < -1 || f2 >= nf || f2 == j) return "bad neighbour link";
            keys.push_back(((uint64_t)std::min(a, b) << 32) | (uint64_t)std::max(a, b));
            if (f2 >= 0) {
                bool back = false;
                for (int e = 0; e < 3; ++e) {
                    const int64_t c = m.faceV[3 * f2 + e], d = m.faceV[3 * f2 + (e + 1) % 3];
                    if (((c == a && d == b) || (c == b && d == a)) && m.faceF[3 * f2 + e] == j) back = true;
                }
                if (!back) return "neighbour links are not reciprocal";
            }
        }
    std::sort(keys.begin(), keys.end());
    for (size_t i = 2; i < keys.size(); ++i) if (keys[i] == keys[i - 2]) return "an edge is shared by more than two faces";
    for (int64_t i = 0; i < nv; ++i) {
        const int64_t f = m.startFace[i];
        if (f < -1 || f >= nf) return "bad startFace";
        if (f >= 0 && m.faceV[3 * f] != i && m.faceV[3 * f + 1] != i && m.faceV[3 * f + 2] != i) return "startFace does not contain its vertex";
    }
    return nullptr;
}
bool FinalCounts(const PbrtGpuSubdivMesh &m, int nLevels, uint64_t *nv, uint64_t *nf) {
    uint32_t nB = 0;
    for (size_t i = 0; i < 3 * (size_t)m.nFaces; ++i) nB += m.faceF[i] < 0;
    return lsd::Counts(m.nVertices, m.nFaces, nB, nLevels, nv, nf) && *nv < 0x7fffffffull && 3 * *nf < 0x7fffffffull;
}
}  // namespace

extern "C" int pbrtgpu_loopsubdiv_counts(const PbrtGpuSubdivMesh *mesh, int nLevels, uint32_t *nVerticesOut, uint32_t *nTrianglesOut) {
    if (!mesh || !nVerticesOut || !nTrianglesOut) return pbrtgpu_set_error_text(PBRTGPU_EINVAL, "pbrtgpu_loopsubdiv_counts: null argument");
    if (const char *why = ValidateSubdivMesh(*mesh, nLevels))
        return pbrtgpu_set_error_text(PBRTGPU_EINVAL, (std::string("pbrtgpu_loopsubdiv: ") + why).c_str());
    uint64_t nv = 0, nf = 0;
    if (!FinalCounts(*mesh, nLevels, &nv, &nf)) return pbrtgpu_set_error_text(PBRTGPU_EINVAL, "pbrtgpu_loopsubdiv: the neighbour links do not describe a manifold, or the refined mesh exceeds 2^31 indices");
    *nVerticesOut = (uint32_t)nv;
    *nTrianglesOut = (uint32_t)nf;
    return PBRTGPU_OK;
}

extern "C" int pbrtgpu_loopsubdiv_refine(const PbrtGpuSubdivMesh *mesh, int nLevels, float *P, float *N, int32_t *vertexIndices) {
    if (!mesh || !P || !N || !vertexIndices) return pbrtgpu_set_error_text(PBRTGPU_EINVAL, "pbrtgpu_loopsubdiv_refine: null argument");
    if (const char *why = ValidateSubdivMesh(*mesh, nLevels))
        return pbrtgpu_set_error_text(PBRTGPU_EINVAL, (std::string("pbrtgpu_loopsubdiv: ") + why).c_str());
    uint64_t nv = 0, nf = 0;
    if (!FinalCounts(*mesh, nLevels, &nv, &nf)) return pbrtgpu_set_error_text(PBRTGPU_EINVAL, "pbrtgpu_loopsubdiv: the neighbour links do not describe a manifold, or the refined mesh exceeds 2^31 indices");
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
        return pbrtgpu_set_error_text(PBRTGPU_ENODEVICE, "no CUDA device visible; libpbrtgpu has no CPU fallback");
    g_lsdErr = cudaSuccess;
    const int rc = lsd::Refine(*mesh, nLevels, P, N, vertexIndices, nullptr);
    LsdCheck(cudaDeviceSynchronize());
    if (rc == 2) return pbrtgpu_set_error_text(PBRTGPU_EINVAL, "pbrtgpu_loopsubdiv: the neighbour links do not describe a manifold");
    if (rc != 0 || g_lsdErr != cudaSuccess) {
        const std::string msg = std::string("pbrtgpu_loopsubdiv_refine: ") + cudaGetErrorString(g_lsdErr == cudaSuccess ? cudaErrorUnknown : g_lsdErr);
        return pbrtgpu_set_error_text(PBRTGPU_ECUDA, msg.c_str());
    }
    return PBRTGPU_OK;
}
