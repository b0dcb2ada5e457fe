// device_on_host.cpp -- TEST INFRASTRUCTURE.  The device headers' traversal and hit-geometry code (dscene.cuh, danim.cuh,
// dmath.cuh, refmath.h) compiled for the HOST by g++ (no FMA contraction, like the device build's -fmad=false), so that
// the CPU test stage can compare the product's own arithmetic with the oracle, bit for bit, without a GPU: a guard for
// edits made while no GPU is at hand.  It is never linked into the product; only tests/ load it.
//
// Build with -fvisibility=hidden -Wl,-Bsymbolic (tests/test_device_on_host.py does): libpbrtgpu.so exports host stubs of
// the same dev:: functions (nvcc emits them for __device__ code), and a process that has the product library loaded
// must not bind this harness's calls to those stubs.
//
// The shims below give the handful of CUDA intrinsics the headers use their plain C++ meaning.
#define DOH_EXPORT extern "C" __attribute__((visibility("default")))
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>   // float4 / uint4 / make_float4; __device__ etc. expand to nothing for a host compiler

#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
using std::isnan;
using std::isinf;
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __brev(unsigned x) {
    x = (x << 16) | (x >> 16);
    x = ((x & 0x00ff00ffu) << 8) | ((x & 0xff00ff00u) >> 8);
    x = ((x & 0x0f0f0f0fu) << 4) | ((x & 0xf0f0f0f0u) >> 4);
    x = ((x & 0x33333333u) << 2) | ((x & 0xccccccccu) >> 2);
    x = ((x & 0x55555555u) << 1) | ((x & 0xaaaaaaaau) >> 1);
    return x;
}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
// ---- the vertex-wavefront kernels (kwave.cuh) run with launches as loops, one "thread" at a time: a thread is alone in its
// warp (every warp-wide primitive sees just the calling lane), which is a schedule the kernels must be correct under
#undef __launch_bounds__
#define __launch_bounds__(...)
static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;
static inline unsigned doh_lane_bit() { return 1u << (threadIdx.x & 31); }
static inline unsigned __activemask() { return doh_lane_bit(); }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? doh_lane_bit() : 0u; }
static inline int __any_sync(unsigned, int p) { return p != 0; }
template <typename T> static inline unsigned __match_any_sync(unsigned, T) { return doh_lane_bit(); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int) { return T(0) * v; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline float atomicAdd(float *p, float v) { float o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned atomicOr(unsigned *p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
template <typename T> static inline void __stcg(T *p, T v) { *p = v; }
using std::min;
using std::max;

#include "../../pbrt_v2_b200/csrc/device/dscene.cuh"
#include "../../pbrt_v2_b200/csrc/device/dshade.cuh"
#include "../../pbrt_v2_b200/csrc/device/dpath.cuh"
#include "../../pbrt_v2_b200/csrc/device/dvertex.cuh"
#include "../../pbrt_v2_b200/csrc/device/kwave.cuh"

using namespace dev;

namespace {
int g_vertexForm = 0;
int g_waveSort = 1;   // the vertex wavefront's phase 1 with its slots sorted by next action (WvArgs::sortList)
struct HostScene {
    std::vector<float4> nodes, prims;
#if PBRT_PAIR_NODES
    std::vector<float4> pairs;
    std::vector<uint32_t> pairOfNode;
#endif
#if PBRT_KD_INTERLEAVED
    std::vector<float4> kdRec;
#endif
    std::vector<uint4> tris;
    std::vector<float4> top;   // heap-ordered top of the scene BVH (what the kernels stage in shared memory)
    DScene d;
};
// -DPBRT_TOP_LEVELS=n builds: the walks read the first n levels from the heap-ordered table, as the kernels do from shared memory
#define DOH_TOP(sc) (PBRT_TOP_LEVELS ? (sc).topNodes : nullptr)
// the same packing as pbrtgpu_scene_create (device/pbrtgpu.cu): nodes -> 2 x float4, primitives -> 3 x float4
void Pack(const PbrtGpuSceneDesc *desc, HostScene *h) {
    h->nodes.resize(2 * (size_t)desc->nNodes);
    for (uint32_t i = 0; i < desc->nNodes; ++i) {
        const PbrtGpuBVHNode &n = desc->nodes[i];
        uint32_t meta = (uint32_t)n.nPrimitives | ((uint32_t)n.axis << 8);
        h->nodes[2 * i] = make_float4(n.bmin[0], n.bmin[1], n.bmin[2], n.bmax[0]);
        h->nodes[2 * i + 1] = make_float4(n.bmax[1], n.bmax[2], __uint_as_float(n.offset), __uint_as_float(meta));
    }
    std::vector<int> owner(desc->nPrimitives, -1);
    for (uint32_t k = 0; k < desc->nAnimated; ++k)
        for (uint32_t i = 0; i < desc->animated[k].nPrimitives; ++i) owner[desc->animated[k].firstPrimitive + i] = (int)k;
    h->prims.resize(3 * (size_t)desc->nPrimitives);
    for (uint32_t i = 0; i < desc->nPrimitives; ++i) {
        const PbrtGpuPrimitive &p = desc->primitives[i];
        float v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        uint32_t w0 = p.type, w1 = p.index, w2 = 0;
        if (p.type != PBRTGPU_PRIM_ANIMATED) {
            if (p.type == PBRTGPU_PRIM_TRIANGLE) {
                const PbrtGpuTriangle &t = desc->triangles[p.index];
                for (int k = 0; k < 3; ++k) memcpy(v + 3 * k, desc->P + 3 * (size_t)t.v[k], 12);
            }
            w0 = p.type | ((uint32_t)(owner[i] + 1) << 8);
            w2 = (uint32_t)p.material | ((uint32_t)(p.areaLight + 1) << 16);
        }
        h->prims[3 * i] = make_float4(v[0], v[1], v[2], __uint_as_float(w0));
        h->prims[3 * i + 1] = make_float4(v[3], v[4], v[5], __uint_as_float(w1));
        h->prims[3 * i + 2] = make_float4(v[6], v[7], v[8], __uint_as_float(w2));
    }
    h->tris.resize(desc->nTriangles);
    for (uint32_t i = 0; i < desc->nTriangles; ++i)
        h->tris[i] = make_uint4(desc->triangles[i].v[0], desc->triangles[i].v[1], desc->triangles[i].v[2], desc->triangles[i].mesh);
    DScene &d = h->d;
    memset(&d, 0, sizeof(d));
    d.nodes = h->nodes.data(); d.prims = h->prims.data(); d.tris = h->tris.data(); d.meshes = desc->meshes;
    BuildTopNodes(desc, PBRT_TOP_LEVELS, &h->top);
    d.topNodes = h->top.data();
#if PBRT_KD_INTERLEAVED
    BuildKdRecords(desc, &h->kdRec);
    d.kdRec = h->kdRec.data();
#endif
#if PBRT_PAIR_NODES
    BuildNodePairs(desc, &h->pairs, &h->pairOfNode);
    d.pairs = h->pairs.data(); d.pairOfNode = h->pairOfNode.data();
#endif
    d.N = desc->N; d.uv = desc->uv; d.quadrics = desc->quadrics; d.xforms = desc->xforms; d.materials = desc->materials;
    d.measured = desc->measured; d.kdNodes = desc->kdNodes; d.kdData = desc->kdData; d.lights = desc->lights;
    d.lightShapes = desc->lightShapes; d.animated = desc->animated;
#if PBRT_INFINITE
    d.envMaps = desc->envMaps; d.envData = desc->envData;
#endif
    d.nNodes = desc->nNodes; d.nPrims = desc->nPrimitives; d.nLights = desc->nLights; d.nAnimated = desc->nAnimated;
}
}  // namespace

namespace {

// the task's MT19937 stream (core/rng.cpp), tempered, addressable ahead of the read position like the device's
// two-block window: draws at [pos, pos + k) can be looked at before `pos` moves
struct Stream {
    uint32_t mt[MT_N];
    std::vector<std::vector<uint32_t>> blocks;   // block b: the raw state words the draws [b * 624, (b + 1) * 624) are tempered from
    explicit Stream(uint32_t seed) { SeedMT(mt, seed); }
    void Regen() {
        for (int k = 0; k < MT_N; ++k) {
            uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % MT_N] & 0x7fffffffu);
            mt[k] = mt[(k + MT_M) % MT_N] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
    }
    const std::vector<uint32_t> &Block(size_t b) {
        while (blocks.size() <= b) {
            Regen();
            blocks.emplace_back(mt, mt + MT_N);
        }
        return blocks[b];
    }
    uint32_t At(size_t i) { return Temper(Block(i / MT_N)[i % MT_N]); }
};

// Sampler::ComputeSubWindow / ImageFilm extent, as the host code of pbrtgpu.cu computes them
struct Extent { int xs, ys, xc, yc, sx0, sx1, sy0, sy1; };
Extent ComputeExtentH(const PbrtGpuRenderDesc &rd) {
    Extent e;
    e.xs = (int)ceilf(rd.xResolution * rd.cropWindow[0]);
    e.xc = std::max(1, (int)ceilf(rd.xResolution * rd.cropWindow[1]) - e.xs);
    e.ys = (int)ceilf(rd.yResolution * rd.cropWindow[2]);
    e.yc = std::max(1, (int)ceilf(rd.yResolution * rd.cropWindow[3]) - e.ys);
    e.sx0 = (int)floorf(e.xs + 0.5f - rd.filterXWidth);
    e.sx1 = (int)ceilf(e.xs - 0.5f + e.xc + rd.filterXWidth);
    e.sy0 = (int)floorf(e.ys + 0.5f - rd.filterYWidth);
    e.sy1 = (int)ceilf(e.ys - 0.5f + e.yc + rd.filterYWidth);
    return e;
}
void SubWindow(const Extent &e, int num, int count, int *x0, int *x1, int *y0, int *y1) {
    int dx = e.sx1 - e.sx0, dy = e.sy1 - e.sy0;
    int nx = count, ny = 1;
    while ((nx & 0x1) == 0 && 2 * dx * ny < dy * nx) { nx >>= 1; ny <<= 1; }
    int xo = num % nx, yo = num / nx;
    float tx0 = float(xo) / float(nx), tx1 = float(xo + 1) / float(nx);
    float ty0 = float(yo) / float(ny), ty1 = float(yo + 1) / float(ny);
    auto lerp = [](float t, float a, float b) { return (1.f - t) * a + t * b; };
    *x0 = (int)floorf(lerp(tx0, e.sx0, e.sx1)); *x1 = (int)floorf(lerp(tx1, e.sx0, e.sx1));
    *y0 = (int)floorf(lerp(ty0, e.sy0, e.sy1)); *y1 = (int)floorf(lerp(ty1, e.sy0, e.sy1));
}

// the Sample layout -> row / scramble positions, as BuildLayout in pbrtgpu.cu lays them out
void LayoutH(const PbrtGpuRenderDesc &rd, SampleLayoutDev *L) {
    memset(L, 0, sizeof(*L));
    L->n1D = rd.n1D; L->n2D = rd.n2D;
    int row = 3, scr = 5;
    for (int j = 0; j < rd.n1D; ++j) { L->size1D[j] = rd.size1D[j]; L->row1D[j] = (uint16_t)row; L->scr1D[j] = (uint16_t)scr; row += rd.size1D[j]; scr += 1; }
    for (int j = 0; j < rd.n2D; ++j) { L->size2D[j] = rd.size2D[j]; L->row2D[j] = (uint16_t)row; L->scr2D[j] = (uint16_t)scr; row += rd.size2D[j]; scr += 2; }
    L->rowStride = (row + 7) / 8 * 8;
}

// LDPixelSample (montecarlo.cpp:200-252) on sequence INDICES: what k_tables / k_tables_general leave in a task's rows
// ImageFilm::AddSample: through the footprint bits for the 0.5 box (FilmAdd in pbrtgpu.cu, without atomics), through the
// weight table otherwise (dpath.cuh FilmAddFilteredRow, PBRT_FILTER builds: the device code itself)
void FilmAddH(const RenderParams &rp, PbrtGpuPixel *film, int xPos, int yPos, int code, const uint16_t *row, const uint32_t *scr,
              const Spec &Ls) {
#if PBRT_FILTER
    if (rp.filterTable) { FilmAddFilteredRow(rp, film, row, scr, xPos, yPos, Ls); return; }
#endif
    (void)row; (void)scr;
    float xyz[3];
    xyz[0] = 0.412453f * Ls.r + 0.357580f * Ls.g + 0.180423f * Ls.b;
    xyz[1] = 0.212671f * Ls.r + 0.715160f * Ls.g + 0.072169f * Ls.b;
    xyz[2] = 0.019334f * Ls.r + 0.119193f * Ls.g + 0.950227f * Ls.b;
    const int fy0 = yPos - ((code & LS_UP) ? 1 : 0), fy1 = yPos + ((code & LS_DOWN) ? 1 : 0);
    const int fx0 = xPos - ((code & LS_LEFT) ? 1 : 0), fx1 = xPos + ((code & LS_RIGHT) ? 1 : 0);
    for (int y = fy0; y <= fy1; ++y)
        for (int x = fx0; x <= fx1; ++x) {
            PbrtGpuPixel &px = film[(size_t)(y - rp.yPixelStart) * rp.xPixelCount + (x - rp.xPixelStart)];
            px.Lxyz[0] += xyz[0]; px.Lxyz[1] += xyz[1]; px.Lxyz[2] += xyz[2]; px.weightSum += 1.f;
        }
}

void TablesH(const SampleLayoutDev &L, int spp, Stream &rng, int *pos, std::vector<uint16_t> *rows, uint32_t *scr) {
    rows->assign((size_t)spp * L.rowStride, 0);
    auto array = [&](int n, int dims, int rowAt, int scrAt) {
        for (int k = 0; k < dims; ++k) scr[scrAt + k] = rng.At((*pos)++);
        std::vector<uint32_t> idx((size_t)spp * n);
        for (size_t i = 0; i < idx.size(); ++i) idx[i] = (uint32_t)i;
        for (int i = 0; i < spp; ++i)   // Shuffle(samples + i * n, n, dims, rng)
            for (int k = 0; k < n; ++k) {
                uint32_t other = k + rng.At((*pos)++) % (uint32_t)(n - k);
                std::swap(idx[(size_t)i * n + k], idx[(size_t)i * n + other]);
            }
        for (int i = 0; i < spp; ++i) {   // Shuffle(samples, nPixel, dims * n, rng)
            uint32_t other = i + rng.At((*pos)++) % (uint32_t)(spp - i);
            for (int k = 0; k < n; ++k) std::swap(idx[(size_t)i * n + k], idx[(size_t)other * n + k]);
        }
        for (int i = 0; i < spp; ++i)
            for (int k = 0; k < n; ++k) (*rows)[(size_t)i * L.rowStride + rowAt + k] = (uint16_t)idx[(size_t)i * n + k];
    };
    array(1, 2, 0, 0); array(1, 2, 1, 2); array(1, 1, 2, 4);
    for (int j = 0; j < L.n1D; ++j) array(L.size1D[j], 1, L.row1D[j], L.scr1D[j]);
    for (int j = 0; j < L.n2D; ++j) array(L.size2D[j], 2, L.row2D[j], L.scr2D[j]);
}

// One task, pixel by pixel, sample by sample: the path state machine (path_body.inc) run to the end of every
// camera path by a single host thread -- the tables for bounces < 3, the task's stream after that, exactly what the
// kernels do in two phases.
template <int INTEG>
void RunTaskH(const DScene &sc, const RenderParams &rp, const PbrtGpuRenderDesc &rd, int taskNum, const Extent &ext,
              PbrtGpuPixel *film) {
    constexpr int PHASE = 2;
    constexpr bool STATS = false;
    int wx0, wx1, wy0, wy1;
    SubWindow(ext, taskNum, rd.nTasks, &wx0, &wx1, &wy0, &wy1);
    if (wx0 == wx1 || wy0 == wy1) return;
    Stream rng((uint32_t)taskNum);
    int mti = 0;   // position in the task's stream
    const int spp = rp.spp;
    const int rowStride = rp.lay.rowStride;
    std::vector<uint16_t> rows;
    uint32_t scrBuf[SCR_STRIDE];
    float4 *lsBase = nullptr, *recBase = nullptr;   // (phase-1 outputs: unused, PHASE == 2)
    auto rngF = [&](int k) -> float { return UIntToFloat(rng.At((size_t)(mti + k))); };
    for (int yPos = wy0; yPos < wy1; ++yPos)
        for (int xPos = wx0; xPos < wx1; ++xPos) {
            TablesH(rp.lay, spp, rng, &mti, &rows, scrBuf);
            const uint16_t *permBase = rows.data();
            const uint32_t *scr = scrBuf;
            for (int s = 0; s < spp; ++s) {
                int state = S_NEW_SAMPLE, bounces = 0, code = 0, lightNum = 0;
                Ray ray;
                ray.o = V3(0, 0, 0); ray.d = V3(0, 0, 1); ray.mint = 0.f; ray.maxt = 0.f; ray.time = 0.f;
                Hit hit;
                hit.prim = -1; hit.t = 0.f; hit.b1 = hit.b2 = 0.f; hit.maxtIn = 0.f;
                bool found = false;
                Spec T(1.f), L(0.f), Ld(0.f), pend(0.f);
                ShadeCtx ctx;
                ctx.p = ctx.nn = ctx.ng = ctx.sn = ctx.tn = V3(0, 0, 1);
                ctx.material = 0; ctx.areaLight = -1; ctx.rayEpsilon = 0.f;
                V3 wo(0, 0, 1), wiW(0, 0, 1);
                float bsU0 = 0.f, bsU1 = 0.f, bsUc = 0.f, sPdf = 0.f, visMaxt = 0.f;
                unsigned long long nShadedN = 0;
                int dlI = 0, dlJ = 0;
                Spec dlSum(0.f), dlAll(0.f);
                bool ctxDirty = false;
                bool specularBounce = false;   // (PBRT_SPECULAR builds: path.cpp:53)
                int dlDepth = 0, dlStage = 0, dlNodes = 0;   // (PBRT_SPECULAR builds: the recursion of DirectLightingIntegrator)
                bool dlReturn = false;
                DlFrame dlFrames[DL_MAX_DEPTH];
                (void)ctxDirty; (void)nShadedN; (void)specularBounce; (void)dlDepth; (void)dlStage; (void)dlNodes; (void)dlReturn;
                (void)dlFrames;
                TraceStats st;
                st.nodes = st.prims = 0;
                bool first = true;
                for (;;) {
#define PATH_FEVAL_IN Spec(0.f)
#define PATH_EVALWO_IN V3(0, 0, 1)
#define PATH_EVALWI_IN V3(0, 0, 1)
#define PATH_DEFER_MEASURED false
#define PATH_ANIM true
#define PATH_INTEG INTEG
#define PATH_COUNT_TAIL() ((void)0)
#define PATH_DL_FRAME(i) dlFrames[i]
#include "../../pbrt_v2_b200/csrc/device/path_body.inc"
#undef PATH_DL_FRAME
#undef PATH_COUNT_TAIL
#undef PATH_FEVAL_IN
#undef PATH_EVALWO_IN
#undef PATH_EVALWI_IN
#undef PATH_DEFER_MEASURED
#undef PATH_ANIM
#undef PATH_INTEG
                    (void)deferEval;
                    if (state == S_END_SAMPLE) break;
                    if (state == S_NEW_SAMPLE && first) {
                        // the camera ray, as in k_logic / k_path (perspective.cpp:81-138)
                        first = false;
                        const uint16_t *row = permBase + (size_t)s * rowStride;
                        const uint32_t pi = row[0];
                        const float imageX = xPos + VanDerCorput(pi, scr[0]);
                        const float imageY = yPos + Sobol2(pi, scr[1]);
                        code = FilmFootprint(rp, xPos, yPos, imageX, imageY);
                        V3 Pcamera = XfPoint(rp.rasterToCamera, V3(imageX, imageY, 0));
                        V3 dir = Normalize(V3(Pcamera.x, Pcamera.y, Pcamera.z));
                        ray.o = V3(0, 0, 0); ray.d = dir; ray.mint = 0.f; ray.maxt = INF_F;
                        ray.time = Lerp(VanDerCorput(row[2], scr[4]), rp.shutterOpen, rp.shutterClose);
                        if (rp.lensRadius > 0.f) {
                            const uint32_t pl = row[1];
                            float lensU, lensV;
                            ConcentricSampleDisk(VanDerCorput(pl, scr[2]), Sobol2(pl, scr[3]), &lensU, &lensV);
                            lensU *= rp.lensRadius;
                            lensV *= rp.lensRadius;
                            float ft = rp.focalDistance / ray.d.z;
                            V3 Pfocus = RayAt(ray, ft);
                            ray.o = V3(lensU, lensV, 0.f);
                            ray.d = Normalize(Pfocus - ray.o);
                        }
                        ray.o = XfPoint(rp.cameraToWorld, ray.o);
                        ray.d = XfVector(rp.cameraToWorld, ray.d);
                        T = Spec(1.f); L = Spec(0.f);
                        bounces = 0;
                        wantTrace = true; anyHit = false;
                        state = S_HIT_RESULT;
                    }
                    if (wantTrace) found = Traverse<false, true>(sc, ray, anyHit, &hit, &st, DOH_TOP(sc));
                }
                // DirectLightingIntegrator: SpecularReflect/Transmit's BSDFSample(rng) draws (what k_film_direct accounts for)
#if PBRT_SPECULAR
                if (INTEG == PBRTGPU_INTEGRATOR_DIRECTLIGHTING) mti += 6 * dlNodes;   // every node that made the two specular calls
#else
                if (INTEG == PBRTGPU_INTEGRATOR_DIRECTLIGHTING && (code & LS_HIT) && 0 + 1 < rp.maxDepth) mti += 6;
#endif
                // ImageFilm::AddSample through the footprint bits (FilmAdd in pbrtgpu.cu, without atomics)
                const Spec Ls = SanitizeRadiance(L);
                FilmAddH(rp, film, xPos, yPos, code, permBase + (size_t)s * rowStride, scr, Ls);
            }
        }
}

// The same task through the VERTEX form of PathIntegrator::Li (dvertex.cuh: VertexA -> the three BSDF values -> VertexB ->
// the three rays back to back), which is what the vertex-synchronous kernels (k_vertex) execute per lane.
void RunTaskVertexH(const DScene &sc, const RenderParams &rp, const PbrtGpuRenderDesc &rd, int taskNum, const Extent &ext,
                    PbrtGpuPixel *film) {
    int wx0, wx1, wy0, wy1;
    SubWindow(ext, taskNum, rd.nTasks, &wx0, &wx1, &wy0, &wy1);
    if (wx0 == wx1 || wy0 == wy1) return;
    Stream rng((uint32_t)taskNum);
    int mti = 0;
    const int spp = rp.spp;
    std::vector<uint16_t> rows;
    uint32_t scrBuf[SCR_STRIDE];
    auto rngF = [&](int k) -> float { return UIntToFloat(rng.At((size_t)(mti + k))); };
    for (int yPos = wy0; yPos < wy1; ++yPos)
        for (int xPos = wx0; xPos < wx1; ++xPos) {
            TablesH(rp.lay, spp, rng, &mti, &rows, scrBuf);
            for (int s = 0; s < spp; ++s) {
                const uint16_t *row = rows.data() + (size_t)s * rp.lay.rowStride;
                Ray ray;
                const int code = CameraRay(rp, row, scrBuf, xPos, yPos, &ray);
                Spec T(1.f), L(0.f);
                int bounces = 0;
                bool afterSpecular = false;
                Hit hit;
                TraceStats st;
                st.nodes = st.prims = 0;
                bool found = Traverse<false, true>(sc, ray, false, &hit, &st, DOH_TOP(sc));
                if (!found) EscapedRadiance(sc, ray.d, true, false, T, L);
                while (found) {
                    VertexRandoms u;
                    const int nLights = (int)sc.nLights;
                    int k0 = 0;
                    if (bounces < 3) VertexRandomsFromTables(row, scrBuf, bounces, &u);
                    else {
                        if (nLights > 0) {
                            u.lnU = rngF(0); u.lsU0 = rngF(1); u.lsU1 = rngF(2);
                            u.bsU0 = rngF(4); u.bsU1 = rngF(5); u.bsUc = rngF(6);
                            k0 = 7;
                        } else u.lnU = u.lsU0 = u.lsU1 = u.bsU0 = u.bsU1 = u.bsUc = 0.f;
                        u.pU0 = rngF(k0); u.pU1 = rngF(k0 + 1); u.pUc = rngF(k0 + 2);
                    }
                    ShadeCtx ctx;
                    Vertex v;
                    VertexA<true>(sc, ray, hit, bounces, afterSpecular, u, T, L, ctx, &v, nullptr);
                    Spec f[3] = {Spec(0.f), Spec(0.f), Spec(0.f)};
                    for (int e = 0; e < 3; ++e) {
#if PBRT_SPECULAR
                        if (e == VX_PATH && v.specSampled) continue;
#endif
                        if (v.need[e]) f[e] = VertexEval(sc, v.material, v.evalWo, v.evalWi[e], v.evalFlags[e]);
                    }
                    bool rrDrawn = false;
                    const float rrU = bounces > 3 ? rngF(k0 + 3) : 0.f;
                    const bool cont = VertexB(f, bounces, rp.maxDepth, T, rrU, &v, &rrDrawn);
                    if (bounces >= 3) mti += k0 + 3 + (rrDrawn ? 1 : 0);
                    Spec Ld(0.f);
                    if (v.ray[VX_LIGHT]) {
                        Ray r; r.o = v.p; r.d = v.wiW[VX_LIGHT]; r.mint = v.eps; r.maxt = v.visMaxt; r.time = ray.time;
                        Hit h;
                        if (!Traverse<false, true>(sc, r, true, &h, &st, DOH_TOP(sc))) Ld = Ld + v.pendS;
                    }
                    if (v.ray[VX_MIS]) {
                        Ray r; r.o = v.p; r.d = v.wiW[VX_MIS]; r.mint = v.eps; r.maxt = INF_F; r.time = ray.time;
                        Hit h;
                        const bool fnd = Traverse<false, true>(sc, r, false, &h, &st, DOH_TOP(sc));
                        if (MisRayContributes<true>(sc, r, h, fnd, v.lightNum)) Ld = Ld + v.pendM;
                    }
                    L = L + T * ((float)nLights * Ld);
                    T = v.Tnew;
                    if (!cont) break;
#if PBRT_SPECULAR
                    afterSpecular = v.specSampled;
#endif
                    ++bounces;
                    ray.o = v.p; ray.d = v.wiW[VX_PATH]; ray.mint = v.eps; ray.maxt = INF_F;
                    found = Traverse<false, true>(sc, ray, false, &hit, &st, DOH_TOP(sc));
                    if (!found) EscapedRadiance(sc, ray.d, false, afterSpecular, T, L);
                }
                const Spec Ls = SanitizeRadiance(L);
                FilmAddH(rp, film, xPos, yPos, code, row, scrBuf, Ls);
            }
        }
}

// A batch of tasks through the VERTEX WAVEFRONT kernels themselves (kwave.cuh: k_wshade / k_weval / k_wtrace, the product's
// phase 1 of PathIntegrator), launches as loops over a pool of `P` path slots; then the tails (bounce 3 on, the task's RNG
// stream) and the film in sample order, as k_vertex<2> + k_film do.  Pixel by pixel ("rounds") so that a task's stream
// position is handed from one pixel's tails to the next pixel's tables.
// `leadersOnly`: kernels whose warps elect lane 0 (k_wtrace's chunk reservation) run as warps of one lane, lane 0
template <typename K> void LaunchH(unsigned threads, K kernel, bool leadersOnly = false) {
    blockDim = dim3(128, 1, 1);
    gridDim = dim3((threads + 127) / 128, 1, 1);
    for (unsigned i = 0; i < gridDim.x * 128u; ++i) {
        if (leadersOnly && (i & 31u)) continue;
        blockIdx.x = i / 128; threadIdx.x = i % 128;
        kernel();
    }
    threadIdx.x = blockIdx.x = 0;
}
void RunWaveH(const DScene &sc, const RenderParams &rp, const PbrtGpuRenderDesc &rd, const Extent &ext, PbrtGpuPixel *film,
              uint32_t P) {
    std::vector<TaskRec> tasks;
    int rounds = 0;
    for (int t = rd.taskBegin; t < rd.taskEnd; ++t) {
        TaskRec tr;
        tr.taskNum = t;
        SubWindow(ext, t, rd.nTasks, &tr.x0, &tr.x1, &tr.y0, &tr.y1);
        if (tr.x0 == tr.x1 || tr.y0 == tr.y1) continue;
        tasks.push_back(tr);
        rounds = std::max(rounds, (tr.x1 - tr.x0) * (tr.y1 - tr.y0));
    }
    const size_t nb = tasks.size();
    if (!nb) return;
    const int spp = rp.spp;
    const int maskWords = std::max(1, spp / 32);
    std::vector<Stream> rng;
    for (size_t i = 0; i < nb; ++i) rng.emplace_back((uint32_t)tasks[i].taskNum);
    std::vector<int> mti(nb, 0);
    std::vector<uint16_t> perm(nb * spp * PERM_STRIDE);
    std::vector<uint32_t> scr(nb * SCR_STRIDE);
    std::vector<float4> ls(nb * spp), rec(nb * spp * 4), pool((size_t)W_COUNT_ANIM * P);
    std::vector<unsigned> tailCount(nb), counters(WC_COUNT);
    std::vector<uint32_t> tailMask(nb * maskWords), qRay(3 * (size_t)P), qEval(3 * (size_t)P);
    const bool hasMeasured = sc.measured != nullptr;
    for (int r = 0; r < rounds; ++r) {
        for (size_t i = 0; i < nb; ++i) {   // k_tables
            const TaskRec &tr = tasks[i];
            if (r >= (tr.x1 - tr.x0) * (tr.y1 - tr.y0)) continue;
            std::vector<uint16_t> rows;
            TablesH(rp.lay, spp, rng[i], &mti[i], &rows, scr.data() + i * SCR_STRIDE);
            for (int s = 0; s < spp; ++s)
                memcpy(perm.data() + (i * spp + s) * PERM_STRIDE, rows.data() + (size_t)s * rp.lay.rowStride,
                       sizeof(uint16_t) * std::min<int>(PERM_STRIDE, rp.lay.rowStride));
        }
        std::fill(tailCount.begin(), tailCount.end(), 0u);
        std::fill(tailMask.begin(), tailMask.end(), 0u);
        WvArgs wa;
        memset(&wa, 0, sizeof(wa));
        wa.sc = sc; wa.rp = rp;
        wa.b.tasks = tasks.data(); wa.b.nBatch = (int)nb; wa.b.round = r;
        wa.b.perm = perm.data(); wa.b.scr = scr.data(); wa.b.ls = ls.data(); wa.b.rec = rec.data();
        wa.b.tailCount = tailCount.data(); wa.b.tailMask = tailMask.data(); wa.b.maskWords = maskWords;
        wa.pool = pool.data(); wa.P = P; wa.qRay = qRay.data(); wa.qEval = qEval.data(); wa.counters = counters.data();
        wa.totalSamples = (unsigned)(nb * (size_t)spp); wa.stats = nullptr;
        std::fill(counters.begin(), counters.end(), 0u);
        std::vector<uint32_t> sortList(2 * (size_t)P);
        std::vector<uint8_t> sortClass(((size_t)P + 1023) / 1024 * 1024, 0);
        wa.sortList = g_waveSort ? sortList.data() : nullptr;
        wa.sortClass = sortClass.data();
        wa.liveList = nullptr;
        for (int iter = 0;; ++iter) {
            wa.iter = iter;
            for (int k = 0; k < WC_NEXT_SAMPLE; ++k) counters[k] = 0;
            LaunchH(P, [&]() { k_wshade<1, false, true>(wa); });
            if (hasMeasured) LaunchH(1, [&]() { k_weval(wa); }, PBRT_KD_INTERLEAVED != 0);   // (the persistent form elects lane 0)
            LaunchH(1, [&]() { k_wtrace<false, true>(wa); }, true);
            if (g_waveSort) {
                counters[WC_SORT] = counters[WC_SORT + 1] = 0;
                LaunchH(1, [&]() { k_wcompact(wa); });
            }
            if (counters[WC_LIVE] == 0) break;
        }
        // ---- phase 2: the task chains on the same kernels (k_wshade<2>), each chain's RNG window in the kernels' format
        std::vector<uint32_t> cur(nb * MT_N), next(nb * MT_N);
        std::vector<size_t> baseBlock(nb);
        std::vector<int> mtiDev(nb);
        for (size_t i = 0; i < nb; ++i) {
            baseBlock[i] = (size_t)mti[i] / MT_N;
            mtiDev[i] = mti[i] % MT_N;
            memcpy(cur.data() + i * MT_N, rng[i].Block(baseBlock[i]).data(), MT_N * 4);
            memcpy(next.data() + i * MT_N, rng[i].Block(baseBlock[i] + 1).data(), MT_N * 4);
        }
        wa.b.rngCur = cur.data(); wa.b.rngNext = next.data(); wa.b.mti = mtiDev.data(); wa.b.order = nullptr;
        wa.P = std::max<uint32_t>(P, (uint32_t)nb);
        std::vector<float4> pool2((size_t)W_COUNT_ANIM * wa.P);
        std::vector<uint32_t> qRay2(3 * (size_t)wa.P), qEval2(3 * (size_t)wa.P);
        wa.pool = pool2.data(); wa.qRay = qRay2.data(); wa.qEval = qEval2.data();
        std::vector<uint32_t> liveList(2 * nb);
        wa.liveList = liveList.data(); wa.sortList = nullptr; wa.sortClass = nullptr;
        for (int iter = 0;; ++iter) {
            wa.iter = iter;
            for (int k = 0; k < WC_LIVE; ++k) counters[k] = 0;
            counters[WvLiveCounter((iter + 1) & 1)] = 0;
            LaunchH((unsigned)nb, [&]() { k_wshade<2, false, true>(wa); });
            if (hasMeasured) LaunchH(1, [&]() { k_weval(wa); }, PBRT_KD_INTERLEAVED != 0);
            LaunchH(1, [&]() { k_wtrace<false, true>(wa); }, true);
            if (counters[WvLiveCounter((iter + 1) & 1)] == 0) break;
        }
        for (size_t i = 0; i < nb; ++i) {   // where the chain left its task's stream: the window may have moved on by whole blocks
            size_t moved = 0;
            while (memcmp(cur.data() + i * MT_N, rng[i].Block(baseBlock[i] + moved).data(), MT_N * 4) != 0) ++moved;
            mti[i] = (int)((baseBlock[i] + moved) * MT_N) + mtiDev[i];
        }
        // ---- the film: every task's current pixel, samples in order (k_film)
        for (size_t i = 0; i < nb; ++i) {
            const TaskRec &tr = tasks[i];
            const int tw = tr.x1 - tr.x0;
            if (r >= tw * (tr.y1 - tr.y0)) continue;
            const int xPos = tr.x0 + r % tw, yPos = tr.y0 + r / tw;
            for (int s = 0; s < spp; ++s) {
                const float4 v = ls[i * spp + s];
                const int code = __float_as_int(v.w);
                FilmAddH(rp, film, xPos, yPos, code, perm.data() + (i * spp + s) * PERM_STRIDE, scr.data() + i * SCR_STRIDE,
                         Spec(v.x, v.y, v.z));
            }
        }
    }
}

}  // namespace

extern "C" {

// pbrtgpu_render's result computed by one host thread running the device state machine (see RunTaskH)
DOH_EXPORT int doh_render(const PbrtGpuSceneDesc *desc, const PbrtGpuRenderDesc *rd, PbrtGpuPixel *pixels) {
    HostScene hs;
    Pack(desc, &hs);
    const Extent ext = ComputeExtentH(*rd);
    RenderParams rp;
    memset(&rp, 0, sizeof(rp));
    memcpy(rp.rasterToCamera, rd->rasterToCamera, sizeof(rp.rasterToCamera));
    memcpy(rp.cameraToWorld, rd->cameraToWorld, sizeof(rp.cameraToWorld));
    rp.lensRadius = rd->lensRadius; rp.focalDistance = rd->focalDistance;
    rp.shutterOpen = rd->shutterOpen; rp.shutterClose = rd->shutterClose;
    rp.filterXWidth = rd->filterXWidth; rp.filterYWidth = rd->filterYWidth;
    rp.xPixelStart = ext.xs; rp.yPixelStart = ext.ys; rp.xPixelCount = ext.xc; rp.yPixelCount = ext.yc;
    rp.spp = rd->samplesPerPixel; rp.maxDepth = rd->maxDepth;
    rp.integrator = rd->integrator; rp.lightStrategy = rd->lightStrategy;
    LayoutH(*rd, &rp.lay);
#if PBRT_FILTER
    {   // as pbrtgpu_render_device: the 0.5 box keeps the footprint bits, everything else reads the table
        static float table[256];
        bool box = rd->filterXWidth == 0.5f && rd->filterYWidth == 0.5f;
        for (int i = 0; i < 256; ++i) {
            table[i] = rd->hasFilterTable ? rd->filterTable[i] : 1.f;
            if (table[i] != 1.f) box = false;
        }
        rp.filterTable = box ? nullptr : table;
    }
#endif
    memset(pixels, 0, sizeof(PbrtGpuPixel) * (size_t)ext.xc * ext.yc);
    if (g_vertexForm >= 2 && rd->integrator != PBRTGPU_INTEGRATOR_DIRECTLIGHTING) {
        RunWaveH(hs.d, rp, *rd, ext, pixels, (uint32_t)g_vertexForm);   // (the value is the pool size: small pools force slot reuse)
        return 0;
    }
    for (int t = rd->taskEnd - 1; t >= rd->taskBegin; --t) {   // the reference's one-thread order (samplerrenderer.cpp:212)
        if (rd->integrator == PBRTGPU_INTEGRATOR_DIRECTLIGHTING) RunTaskH<1>(hs.d, rp, *rd, t, ext, pixels);
        else if (g_vertexForm) RunTaskVertexH(hs.d, rp, *rd, t, ext, pixels);
        else RunTaskH<0>(hs.d, rp, *rd, t, ext, pixels);
    }
    return 0;
}

// drng.cuh's closed form of the Sobol' (0,2)-sequence word, for the test that compares it with montecarlo.h's loop
DOH_EXPORT uint32_t doh_sobol2_word(uint32_t n) { return Sobol2Word(n); }

// PathIntegrator renders through the vertex form (dvertex.cuh) instead of the state machine (path_body.inc); a value >= 2
// runs the vertex-wavefront kernels (kwave.cuh) over a pool of that many path slots
DOH_EXPORT void doh_set_vertex_form(int on) { g_vertexForm = on; }
DOH_EXPORT void doh_set_wave_sort(int on) { g_waveSort = on; }

// k_intersect<false>'s body (device/pbrtgpu.cu) on the host: BVHAccel::Intersect + the hit's differential geometry
DOH_EXPORT int doh_intersect(const PbrtGpuSceneDesc *desc, const PbrtGpuRay *rays, int n, PbrtGpuHit *hits, int stepwise) {
    HostScene hs;
    Pack(desc, &hs);
    const DScene &sc = hs.d;
    for (int i = 0; i < n; ++i) {
        Ray ray;
        ray.o = V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]);
        ray.d = V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        ray.mint = rays[i].mint; ray.maxt = rays[i].maxt; ray.time = rays[i].time;
        Hit hit;
        TraceStats st;
        st.nodes = st.prims = 0;
        bool found;
        if (stepwise) {   // the wavefront's step-wise walk (k_trace)
            Trav t;
            TravInit(t, ray, false);
#if PBRT_SMEM_STACK
            static uint32_t fakeShared[PBRT_SMEM_STACK * 128];
            t.sstk = fakeShared + 5;
#endif
            if (stepwise == 2) {   // k_trace's two-phase loop: bounded descents, then the leaf
                for (;;) {
                    const int r = TravDescend<false, true>(sc, t, &st, 3, DOH_TOP(sc));
                    if (r == 2 || (r == 1 && TravLeaf<false, true>(sc, t, &st))) break;
                }
            } else
            while (!TravStep<false, true>(sc, t, &st)) {}
            found = t.found; hit = t.hit; ray.maxt = t.ray.maxt;
        } else
            found = Traverse<false, true>(sc, ray, false, &hit, &st, DOH_TOP(sc));
        PbrtGpuHit h;
        memset(&h, 0, sizeof(h));
        h.t = ray.maxt;
        if (found) {
            DiffGeom dg;
            float eps;
            int mat, al;
            uint4 tri;
            uint32_t ptype;
            DXf w2p;
            HitGeometry<true>(sc, ray, hit, &dg, &eps, &mat, &al, &tri, &ptype, &w2p);
            h.hit = 1; h.primitive = (uint32_t)hit.prim; h.b1 = hit.b1; h.b2 = hit.b2;
            h.p[0] = dg.p.x; h.p[1] = dg.p.y; h.p[2] = dg.p.z;
            h.n[0] = dg.nn.x; h.n[1] = dg.nn.y; h.n[2] = dg.nn.z;
        }
        hits[i] = h;
    }
    return 0;
}

DOH_EXPORT int doh_intersect_p(const PbrtGpuSceneDesc *desc, const PbrtGpuRay *rays, int n, uint8_t *occ, int stepwise) {
    HostScene hs;
    Pack(desc, &hs);
    const DScene &sc = hs.d;
    for (int i = 0; i < n; ++i) {
        Ray ray;
        ray.o = V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]);
        ray.d = V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        ray.mint = rays[i].mint; ray.maxt = rays[i].maxt; ray.time = rays[i].time;
        Hit hit;
        TraceStats st;
        st.nodes = st.prims = 0;
        if (stepwise) {
            Trav t;
            TravInit(t, ray, true);
#if PBRT_SMEM_STACK
            static uint32_t fakeShared[PBRT_SMEM_STACK * 128];
            t.sstk = fakeShared + 7;
#endif
            if (stepwise == 2) {
                for (;;) {
                    const int r = TravDescend<false, true>(sc, t, &st, 3, DOH_TOP(sc));
                    if (r == 2 || (r == 1 && TravLeaf<false, true>(sc, t, &st))) break;
                }
            } else
            while (!TravStep<false, true>(sc, t, &st)) {}
            occ[i] = t.found ? 1 : 0;
        } else
            occ[i] = Traverse<false, true>(sc, ray, true, &hit, &st, DOH_TOP(sc)) ? 1 : 0;
    }
    return 0;
}

// the shading steps of path_body.inc at the first hit of each ray, laid out like oracle_shade_test's output
DOH_EXPORT int doh_shade_test(const PbrtGpuSceneDesc *desc, const PbrtGpuRay *rays, int n, const float *u, float *out) {
    HostScene hs;
    Pack(desc, &hs);
    const DScene &sc = hs.d;
    const int nLights = (int)sc.nLights;
    for (int i = 0; i < n; ++i) {
        float *o = out + 36 * (size_t)i;
        for (int k = 0; k < 36; ++k) o[k] = 0.f;
        Ray ray;
        ray.o = V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]);
        ray.d = V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        ray.mint = rays[i].mint; ray.maxt = rays[i].maxt; ray.time = rays[i].time;
        Hit hit;
        TraceStats st;
        st.nodes = st.prims = 0;
        if (!Traverse<false, true>(sc, ray, false, &hit, &st, DOH_TOP(sc))) continue;
        ShadeCtx ctx;
        BuildShadeCtx<true>(sc, ray, hit, &ctx, nullptr);
        const V3 wo = -ray.d;
        auto put = [&](int at, const V3 &v) { o[at] = v.x; o[at + 1] = v.y; o[at + 2] = v.z; };
        put(0, ctx.p); put(3, ctx.nn); put(6, ctx.sn); put(9, ctx.tn); put(12, ctx.ng);
        o[15] = ctx.rayEpsilon;
        if (nLights == 0) continue;
        const float *uu = u + 6 * (size_t)i;
        const int lightNum = mini(Floor2Int(uu[5] * nLights), nLights - 1);
        const PbrtGpuLight &light = sc.lights[lightNum];
        V3 wiW;
        float sPdf;
        Ray vis;
        Spec Li = LightSample_L(sc, light, ctx.p, ctx.rayEpsilon, uu[0], uu[1], &wiW, &sPdf, &vis);
        put(16, wiW); o[19] = sPdf; o[20] = Li.r; o[21] = Li.g; o[22] = Li.b; o[23] = vis.maxt;
        BxSet bx;
        MaterialBxDFs(sc, ctx.material, &bx);
        {   // BSDF::f (path_body.inc, the light-sampling half) and BSDF::Pdf
            V3 evalWi = WorldToLocal(ctx, wiW), evalWo = WorldToLocal(ctx, wo);
            int flags = BSDF_ALL & ~BSDF_SPECULAR;
            if (Dot(wiW, ctx.ng) * Dot(wo, ctx.ng) > 0) flags &= ~BSDF_TRANSMISSION;
            else flags &= ~BSDF_REFLECTION;
            Spec f(0.f);
            for (int k = 0; k < bx.n; ++k)
                if (Matches(bx.type(k), flags)) f = f + BxF(sc, bx, k, evalWo, evalWi);
            o[24] = f.r; o[25] = f.g; o[26] = f.b;
            o[27] = BSDF_PdfLocal(bx, evalWo, evalWi, BSDF_ALL & ~BSDF_SPECULAR);
        }
        {   // BSDF::Sample_f
            V3 evalWo, evalWi, wi2;
            float pdf2;
            int sampledType, evalFlags;
#if PBRT_SPECULAR
            Spec specF(0.f);
            if (BSDF_SampleDir(ctx, bx, wo, uu[2], uu[3], uu[4], BSDF_ALL, &evalWo, &evalWi, &wi2, &pdf2, &sampledType, &evalFlags, &specF)) {
#else
            if (BSDF_SampleDir(ctx, bx, wo, uu[2], uu[3], uu[4], BSDF_ALL, &evalWo, &evalWi, &wi2, &pdf2, &sampledType, &evalFlags)) {
#endif
                Spec f(0.f);
                for (int k = 0; k < bx.n; ++k)
                    if (Matches(bx.type(k), evalFlags)) f = f + BxF(sc, bx, k, evalWo, evalWi);
#if PBRT_SPECULAR
                if (sampledType & BSDF_SPECULAR) f = specF;
#endif
                o[28] = f.r; o[29] = f.g; o[30] = f.b;
                put(31, wi2); o[34] = pdf2;
                o[35] = LightPdf(sc, light, ctx.p, wi2);
            }
        }
    }
    return 0;
}

}  // extern "C"
