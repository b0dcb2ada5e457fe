// pbrtgpu.cu -- CUDA kernels and C ABI (include/pbrtgpu.h) of the B200 path-tracing
// hot path.  sm_100a only; no CPU fallback anywhere in this file.
//
// Work decomposition (DESIGN.md "Execution model"):
//   unit        = one reference render task (SamplerRendererTask, its own MT19937
//                 stream seeded with taskNum, samplerrenderer.cpp:72-73)
//   k_tables    = one WARP per task: generates the LD sample tables of the task's
//                 current pixel (LDPixelSample) and advances the task's RNG state
//   k_path      = one THREAD per task: runs the pixel's spp camera samples one after
//                 the other (the RNG offset of sample i+1 depends on how many
//                 draws the tail of sample i's path consumed), as a state machine
//                 with a single BVH-traversal site per loop iteration so that all
//                 lanes of a warp traverse together whatever stage they are in
//   rounds      = tasks with more than one pixel alternate k_tables / k_path once
//                 per pixel (pixel p+1's tables start where pixel p's paths stopped)
#include "dshade.cuh"
#include "drng.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace dev;

// ===================================================================== device
struct TaskRec {
    int taskNum;
    int x0, x1, y0, y1;   // Sampler::ComputeSubWindow tile
};

struct RenderParams {
    float rasterToCamera[16];
    float cameraToWorld[16];
    float lensRadius, focalDistance;
    float filterXWidth, filterYWidth;
    int xPixelStart, yPixelStart, xPixelCount, yPixelCount;
    int spp, maxDepth;
};

struct BatchArgs {
    const TaskRec *tasks;
    int nBatch;
    int round;
    uint16_t *perm;      // [nBatch][spp][PERM_STRIDE]
    uint32_t *scr;       // [nBatch][SCR_STRIDE]
    uint32_t *rngCur;    // [nBatch][624]
    uint32_t *rngNext;   // [nBatch][624]
    int *mti;            // [nBatch]
    float4 *ls;          // [nBatch][spp]     finished radiance (xyz) + footprint/TAIL bits (w)
    float4 *rec;         // [nBatch][spp][4]  parked bounce-3 vertices (phase 1 -> phase 2)
};

struct DeviceStats {
    unsigned long long closestRays, shadowRays, nodes, prims, shadedN, cameraSamples;
};

// --------------------------------------------------------------------- k_tables
// LDSampler::GetMoreSamples / LDPixelSample for one pixel of each task.
template <int WPB>
__global__ void __launch_bounds__(32 * WPB) k_tables(BatchArgs a, int spp, int G) {
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = spp + 2;   // +2 u16 (one bank) of padding per array
    const size_t warpWords = MT_N + (size_t)G * stride;   // mt + (T,P as u16 pairs -> G*stride words)
    uint32_t *mt = smem + warp * warpWords;
    uint16_t *T = reinterpret_cast<uint16_t *>(mt + MT_N);
    uint16_t *P = T + (size_t)G * stride;
    const int bi = blockIdx.x * WPB + warp;
    if (bi >= a.nBatch) return;
    const TaskRec tr = a.tasks[bi];
    const int npx = (tr.x1 - tr.x0) * (tr.y1 - tr.y0);
    if (a.round >= npx) return;

    uint32_t *cur = a.rngCur + (size_t)bi * MT_N, *next = a.rngNext + (size_t)bi * MT_N;
    int mti;
    if (a.round == 0) {   // RNG rng(taskNum), samplerrenderer.cpp:73
        if (lane == 0) SeedMT(mt, (uint32_t)tr.taskNum);
        mti = MT_N;
    } else {
        mti = a.mti[bi];
        const uint32_t *src = cur;
        if (mti >= MT_N) { src = next; mti -= MT_N; }
        for (int j = lane; j < MT_N; j += 32) mt[j] = __ldcg(src + j);
    }
    __syncwarp();

    uint32_t consumed = 0;
    uint16_t *permOut = a.perm + (size_t)bi * spp * PERM_STRIDE;
    uint32_t *scrOut = a.scr + (size_t)bi * SCR_STRIDE;
    for (int a0 = 0; a0 < N_ARRAYS; a0 += G) {
        const int a1 = min(N_ARRAYS, a0 + G);
        const uint32_t end = ArrayStart(a1, spp);
        // stage 1: this group's slice of the pixel's draw stream -> scrambles + shuffle targets
        while (consumed < end) {
            if (mti >= MT_N) { WarpRegenMT(mt, lane); mti = 0; }
            const uint32_t n = min((uint32_t)(MT_N - mti), end - consumed);
            for (uint32_t j = lane; j < n; j += 32) {
                const uint32_t y = Temper(mt[mti + j]);
                int arr, role;
                uint32_t i;
                ClassifyDraw(consumed + j, spp, &arr, &role, &i);
                if (role == 0) scrOut[ScrambleSlot(arr) + i] = y;
                else if (role == 2) T[(arr - a0) * stride + i] = (uint16_t)(i + y % (spp - i));   // montecarlo.h:177
            }
            __syncwarp();
            mti += n;
            consumed += n;
        }
        // stage 2: Fisher-Yates (Shuffle, montecarlo.h:174-181) of the identity, one array per lane
        const int ng = a1 - a0;
        for (int idx = lane; idx < ng * spp; idx += 32) {
            const int l = idx / spp, i = idx - l * spp;
            P[l * stride + i] = (uint16_t)i;
        }
        __syncwarp();
        if (lane < ng) {
            uint16_t *p = P + lane * stride;
            const uint16_t *t = T + lane * stride;
            for (int i = 0; i < spp; ++i) {
                const int j = t[i];
                const uint16_t tmp = p[i];
                p[i] = p[j];
                p[j] = tmp;
            }
        }
        __syncwarp();
        for (int s = lane; s < spp; s += 32)
            for (int l = 0; l < ng; ++l) permOut[(size_t)s * PERM_STRIDE + a0 + l] = P[l * stride + s];
        __syncwarp();
    }
    // save the stream position: cur = block in use, next = the one after it
    for (int j = lane; j < MT_N; j += 32) cur[j] = mt[j];
    __syncwarp();
    WarpRegenMT(mt, lane);
    for (int j = lane; j < MT_N; j += 32) next[j] = mt[j];
    if (lane == 0) a.mti[bi] = mti;
}

// expand one task's tables into the reference's Sample layout (test hook)
__global__ void k_expand_samples(BatchArgs a, int spp, float shutterOpen, float shutterClose, float *out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= spp) return;
    const TaskRec tr = a.tasks[0];
    const uint16_t *row = a.perm + (size_t)s * PERM_STRIDE;
    const uint32_t *scr = a.scr;
    float *o = out + (size_t)s * 37;
    o[0] = tr.x0 + VanDerCorput(row[0], scr[0]);
    o[1] = tr.y0 + Sobol2(row[0], scr[1]);
    o[2] = VanDerCorput(row[1], scr[2]);
    o[3] = Sobol2(row[1], scr[3]);
    o[4] = Lerp(VanDerCorput(row[2], scr[4]), shutterOpen, shutterClose);
    for (int j = 0; j < N_1D; ++j) o[5 + j] = VanDerCorput(row[3 + j], scr[5 + j]);
    for (int j = 0; j < N_2D; ++j) {
        o[5 + N_1D + 2 * j] = VanDerCorput(row[17 + j], scr[19 + 2 * j]);
        o[5 + N_1D + 2 * j + 1] = Sobol2(row[17 + j], scr[20 + 2 * j]);
    }
}

// ----------------------------------------------------------------------- k_path
// The path integrator runs in two phases per pixel (DESIGN.md "Execution model"):
//
//   phase 1  one WARP per task-pixel, lanes = camera samples of that pixel.  Bounces
//            0..2 draw every random number from the pixel's LD tables
//            (path.cpp:71-76,86-88), so the spp samples of a pixel are independent
//            there and run side by side: 32 nearly identical primary rays per warp.
//            A sample either finishes (radiance stored in ls[s]) or reaches its
//            bounce-3 vertex, where PathIntegrator::Li starts consuming the task's
//            sequential RNG stream; it is then parked as a 64-byte record.
//   phase 2  one THREAD per task-pixel walks ls[0..spp) in sample order: finished
//            samples are added to the film (ImageFilm::AddSample order), parked
//            ones are resumed with the RNG exactly where the previous sample's tail
//            left it (UniformSampleOneLight 7 draws, BSDFSample 3, roulette 1).
//
// Both phases are one state machine with a single BVH-traversal site per loop
// iteration, so the lanes of a warp traverse together whatever stage each is in.
enum PathState {
    S_NEW_SAMPLE = 0, S_HIT_RESULT, S_SHADOW_RESULT, S_AFTER_LIGHT, S_MIS_RESULT, S_AFTER_DIRECT, S_END_SAMPLE, S_DONE,
    S_E_LIGHT, S_E_MIS, S_E_PATH   // waiting for the BSDF value at the evaluation site
};
// ls[s].w bit field
enum { LS_TAIL = 1, LS_LEFT = 2, LS_RIGHT = 4, LS_UP = 8, LS_DOWN = 16 };

struct PathArgs {
    DScene sc;
    RenderParams rp;
    BatchArgs b;
    PbrtGpuPixel *film;
    DeviceStats *stats;
    int directFilm;   // 1: every AddSample goes straight to the film (multi-pixel tasks)
    int lockstep;     // phase 1: hand out camera samples 32 at a time (see k_path)
};

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

// ImageFilm::AddSample, film/image.cpp:97-137, box filter (all table weights are 1)
DEV void FilmAdd(const PathArgs &a, int ownX, int ownY, int code, const Spec &L, float acc[4], bool direct) {
    const RenderParams &rp = a.rp;
    float xyz[3];   // RGBToXYZ, core/spectrum.h:58-62
    xyz[0] = 0.412453f * L.r + 0.357580f * L.g + 0.180423f * L.b;
    xyz[1] = 0.212671f * L.r + 0.715160f * L.g + 0.072169f * L.b;
    xyz[2] = 0.019334f * L.r + 0.119193f * L.g + 0.950227f * L.b;
    const int y0 = ownY - ((code & LS_UP) ? 1 : 0), y1 = ownY + ((code & LS_DOWN) ? 1 : 0);
    const int x0 = ownX - ((code & LS_LEFT) ? 1 : 0), x1 = ownX + ((code & LS_RIGHT) ? 1 : 0);
    for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) {
            if (!direct && x == ownX && y == ownY) {
                acc[0] += xyz[0]; acc[1] += xyz[1]; acc[2] += xyz[2]; acc[3] += 1.f;
            } else {
                // a neighbour pixel: the reference does a plain += that can race with the
                // owning task; atomics keep the sum complete, the order is as undefined as there
                PbrtGpuPixel *px = a.film + (size_t)(y - rp.yPixelStart) * rp.xPixelCount + (x - rp.xPixelStart);
                atomicAdd(&px->Lxyz[0], xyz[0]);
                atomicAdd(&px->Lxyz[1], xyz[1]);
                atomicAdd(&px->Lxyz[2], xyz[2]);
                atomicAdd(&px->weightSum, 1.f);
            }
        }
}

// samplerrenderer.cpp:119-133: NaN / negative / infinite luminance -> black
DEV Spec SanitizeRadiance(const Spec &L) {
    if (L.HasNaNs()) return Spec(0.f);
    if ((double)L.y() < -1e-5) return Spec(0.f);
    if (isinf(L.y())) return Spec(0.f);
    return L;
}

template <int PHASE, bool STATS>
// 4 CTAs/SM (<= 128 registers): measured +9 % over the unconstrained 168-register build
__global__ void __launch_bounds__(128, 4) k_path(PathArgs a) {
    __shared__ uint32_t smt[4][MT_N];
    const DScene &sc = a.sc;
    const RenderParams &rp = a.rp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    const int bi = PHASE == 1 ? (int)(blockIdx.x * 4 + warp) : (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int spp = rp.spp;

    bool valid = bi < a.b.nBatch;
    TaskRec tr = valid ? a.b.tasks[bi] : TaskRec{0, 0, 0, 0, 0};
    const int tw = tr.x1 - tr.x0;
    if (valid && a.b.round >= tw * (tr.y1 - tr.y0)) valid = false;
    const int xPos = valid ? tr.x0 + a.b.round % tw : 0;
    const int yPos = valid ? tr.y0 + a.b.round / tw : 0;
    const size_t slot = valid ? (size_t)bi : 0;
    const uint16_t *permBase = a.b.perm + slot * spp * PERM_STRIDE;
    const uint32_t *scr = a.b.scr + slot * SCR_STRIDE;
    const uint32_t *rngCur = a.b.rngCur + slot * MT_N, *rngNext = a.b.rngNext + slot * MT_N;
    float4 *lsBase = a.b.ls + slot * spp;
    float4 *recBase = a.b.rec + slot * spp * 4;
    int mti = (PHASE == 2 && valid) ? a.b.mti[bi] : 0;
    const bool direct = a.directFilm != 0;

    // draw `k` positions past mti (RNG::RandomUInt order), without advancing
    auto rngU = [&](int k) -> uint32_t {
        const int j = mti + k;
        return Temper(j < MT_N ? __ldcg(rngCur + j) : __ldcg(rngNext + (j - MT_N)));
    };
    auto rngF = [&](int k) -> float { return UIntToFloat(rngU(k)); };

    int state = valid ? S_NEW_SAMPLE : S_DONE;
    int s = 0, bounces = 0, nextS = 0, code = 0;
    Ray ray;
    ray.o = V3(0, 0, 0); ray.d = V3(0, 0, 1); ray.mint = 0.f; ray.maxt = 0.f;
    Hit hit;
    hit.prim = -1; hit.t = 0.f; hit.b1 = hit.b2 = 0.f; hit.maxtIn = 0.f;
    bool found = false;
    Spec T(1.f), L(0.f), Ld(0.f), pend(0.f);
    ShadeCtx ctx;
    ctx.p = ctx.nn = ctx.ng = ctx.sn = ctx.tn = V3(0, 0, 1);
    ctx.material = 0; ctx.areaLight = -1; ctx.rayEpsilon = 0.f;
    V3 wo(0, 0, 1);
    int lightNum = 0;
    float bsU0 = 0.f, bsU1 = 0.f, bsUc = 0.f;   // USOL's BSDFSample (kept from S_HIT_RESULT to S_AFTER_LIGHT)
    V3 wiW(0, 0, 1);                            // sampled incident direction (world)
    float sPdf = 0.f, visMaxt = 0.f;            // its pdf; the shadow segment's maxt
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    TraceStats st;
    st.nodes = st.prims = 0;
    unsigned long long nClosest = 0, nShadow = 0, nShadedN = 0;

    for (;;) {
        bool fresh = false;   // phase 1: this lane was handed a new sample index this iteration
        if (PHASE == 1) {
            // ---- hand the pixel's next camera samples to the idle lanes (warp-uniform counter)
            // Samples are handed out 32 at a time, when every lane has finished its previous one:
            // the lanes then stay in step (all primary rays, then all shadow rays, ...), which keeps
            // the BVH walks of a warp the same length; lanes whose path died early idle meanwhile.
            unsigned m = __ballot_sync(FULL, state == S_NEW_SAMPLE);
            const unsigned busy = __ballot_sync(FULL, state != S_NEW_SAMPLE && state != S_DONE);
            if (a.lockstep && busy) m = 0;
            if (m && state == S_NEW_SAMPLE) {
                s = nextS + __popc(m & ((1u << lane) - 1u));
                if (s >= spp) state = S_DONE;
                else fresh = true;
            }
            nextS += __popc(m);
        } else {
            // ---- cooperative RNG top-up: make sure the <= 11 draws of a bounce are addressable
            const bool need = (state == S_HIT_RESULT && found && mti >= MT_N);
            unsigned m = __ballot_sync(FULL, need);
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const size_t sl = __shfl_sync(FULL, slot, l);
                uint32_t *c = a.b.rngCur + sl * MT_N, *n = a.b.rngNext + sl * MT_N;
                uint32_t *mt = smt[warp];
                for (int j = lane; j < MT_N; j += 32) {
                    const uint32_t v = __ldcg(n + j);
                    __stcg(c + j, v);
                    mt[j] = v;
                }
                __syncwarp();
                WarpRegenMT(mt, lane);
                for (int j = lane; j < MT_N; j += 32) __stcg(n + j, mt[j]);
                __syncwarp();
                if (lane == l) mti -= MT_N;
            }
        }
        if (!__any_sync(FULL, state != S_DONE)) break;

        bool wantTrace = false, anyHit = false;
        bool evalReq = false;
        V3 evalWo(0, 0, 1), evalWi(0, 0, 1);
        int evalFlags = 0;

        // ================= part A: consume the traversal result, run up to the next BSDF evaluation
        if (state == S_HIT_RESULT) {
            if (!found) {
                // SamplerRenderer::Li miss (bounce 0: sum of Light::Le == 0) / path.cpp:113-118
                state = S_END_SAMPLE;
            } else if (PHASE == 1 && bounces == 3) {
                // ---- park the sample: from here on the path draws from the task's RNG stream
                float4 *r = recBase + (size_t)s * 4;
                r[0] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint);
                r[1] = make_float4(ray.d.x, ray.d.y, ray.d.z, hit.t);
                r[2] = make_float4(T.r, T.g, T.b, __int_as_float(hit.prim));
                r[3] = make_float4(hit.b1, hit.b2, hit.maxtIn, 0.f);
                lsBase[s] = make_float4(L.r, L.g, L.b, __int_as_float(code | LS_TAIL));
                state = S_NEW_SAMPLE;
            } else {
                // ---- path.cpp:61-80: emitted light, BSDF, UniformSampleOneLight up to bsdf->f()
                BuildShadeCtx(sc, ray, hit, &ctx, STATS ? &nShadedN : nullptr);
                wo = -ray.d;
                if (bounces == 0) L = L + T * HitLe(sc, ctx.areaLight, ctx.ng, wo);
                Ld = Spec(0.f);
                const int nLights = (int)sc.nLights;
                state = S_AFTER_DIRECT;
                if (nLights > 0) {
                    float lnU, lsU0, lsU1;
                    if (bounces < 3) {   // path.cpp:71-76: table samples of depth `bounces`
                        const uint16_t *row = permBase + (size_t)s * PERM_STRIDE;
                        const int d = bounces;
                        lnU = VanDerCorput(row[3 + 4 * d + 1], scr[5 + 4 * d + 1]);
                        const uint32_t pl = row[17 + 3 * d + 0];
                        lsU0 = VanDerCorput(pl, scr[19 + 2 * (3 * d + 0)]);
                        lsU1 = Sobol2(pl, scr[20 + 2 * (3 * d + 0)]);
                        const uint32_t pb = row[17 + 3 * d + 1];
                        bsU0 = VanDerCorput(pb, scr[19 + 2 * (3 * d + 1)]);
                        bsU1 = Sobol2(pb, scr[20 + 2 * (3 * d + 1)]);
                        bsUc = VanDerCorput(row[3 + 4 * d + 2], scr[5 + 4 * d + 2]);
                        // (light uComponent, oneD[4d+0], selects a shape of the ShapeSet: one shape in scope)
                    } else {             // integrator.cpp:95,107-108: lightNum, LightSample(rng), BSDFSample(rng)
                        lnU = rngF(0);
                        lsU0 = rngF(1); lsU1 = rngF(2);
                        bsU0 = rngF(4); bsU1 = rngF(5); bsUc = rngF(6);
                        mti += 7;
                    }
                    lightNum = mini(Floor2Int(lnU * nLights), nLights - 1);
                    const PbrtGpuLight &light = sc.lights[lightNum];
                    // ---- EstimateDirect, light-sampling half (integrator.cpp:123-142)
                    Ray vis;
                    Spec Li = LightSample_L(sc, light, ctx.p, ctx.rayEpsilon, lsU0, lsU1, &wiW, &sPdf, &vis);
                    visMaxt = vis.maxt;   // vis.o == p, vis.mint == rayEpsilon, vis.d == wi bit for bit
                    state = S_AFTER_LIGHT;
                    if (sPdf > 0.f && !Li.IsBlack()) {
                        pend = Li;
                        // Spectrum f = bsdf->f(wo, wi, flags)  -- BSDF::f, reflection.cpp:612-626
                        evalWi = WorldToLocal(ctx, wiW); evalWo = WorldToLocal(ctx, wo);
                        evalFlags = BSDF_ALL & ~BSDF_SPECULAR;
                        if (Dot(wiW, ctx.ng) * Dot(wo, ctx.ng) > 0) evalFlags &= ~BSDF_TRANSMISSION;
                        else evalFlags &= ~BSDF_REFLECTION;
                        evalReq = true;
                        state = S_E_LIGHT;
                    }
                }
            }
        } else if (state == S_SHADOW_RESULT) {
            if (!found) Ld = Ld + pend;   // visibility.Unoccluded(scene), Transmittance == 1
            state = S_AFTER_LIGHT;
        } else if (state == S_MIS_RESULT) {
            if (found) {
                DiffGeom dgl;
                float eps;
                int mat, al;
                uint4 tri;
                uint32_t ptype;
                HitGeometry(sc, ray, hit, &dgl, &eps, &mat, &al, &tri, &ptype);
                // lightIsect.primitive->GetAreaLight() == light  &&  L(p, n, -wi) not black
                const PbrtGpuLight &light = sc.lights[lightNum];
                const bool emits = !Spec(light.L[0], light.L[1], light.L[2]).IsBlack();
                if (al == lightNum && Dot(dgl.nn, -ray.d) > 0.f && emits) Ld = Ld + pend;
            }
            state = S_AFTER_DIRECT;
        }

        if (state == S_AFTER_LIGHT) {
            // ---- EstimateDirect, BSDF-sampling half (integrator.cpp:145-172): choose the direction
            state = S_AFTER_DIRECT;
            const PbrtGpuLight &light = sc.lights[lightNum];
            if (light.type != PBRTGPU_LIGHT_POINT) {
                BxSet bx;
                MaterialBxDFs(sc, ctx.material, &bx);
                int sampledType;
                if (BSDF_SampleDir(ctx, bx, wo, bsU0, bsU1, bsUc, BSDF_ALL & ~BSDF_SPECULAR, &evalWo, &evalWi, &wiW,
                                   &sPdf, &sampledType, &evalFlags)) {
                    evalReq = true;
                    state = S_E_MIS;
                }
            }
        }

        if (state == S_AFTER_DIRECT) {
            // ---- path.cpp:71-94: add direct light, choose the continuation direction
            L = L + T * ((float)(int)sc.nLights * Ld);
            float u0, u1, uc;
            if (bounces < 3) {
                const uint16_t *row = permBase + (size_t)s * PERM_STRIDE;
                const int d = bounces;
                const uint32_t pp = row[17 + 3 * d + 2];
                u0 = VanDerCorput(pp, scr[19 + 2 * (3 * d + 2)]);
                u1 = Sobol2(pp, scr[20 + 2 * (3 * d + 2)]);
                uc = VanDerCorput(row[3 + 4 * d + 3], scr[5 + 4 * d + 3]);
            } else {
                u0 = rngF(0); u1 = rngF(1); uc = rngF(2);
                mti += 3;
            }
            BxSet bx;
            MaterialBxDFs(sc, ctx.material, &bx);
            int sampledType;
            state = S_END_SAMPLE;   // f.IsBlack() || pdf == 0.  -> break
            if (BSDF_SampleDir(ctx, bx, wo, u0, u1, uc, BSDF_ALL, &evalWo, &evalWi, &wiW, &sPdf, &sampledType, &evalFlags)) {
                evalReq = true;
                state = S_E_PATH;
            }
        }

        // ================= the single BSDF-evaluation site (sum of matching BxDF::f, reflection.cpp:572-576,620-623)
        Spec fEval(0.f);
        if (evalReq) {
            BxSet bx;
            MaterialBxDFs(sc, ctx.material, &bx);
            for (int i = 0; i < bx.n; ++i)
                if (Matches(bx.type[i], evalFlags)) fEval = fEval + BxF(sc, bx, i, evalWo, evalWi);
        }

        // ================= part B: use the BSDF value, run up to the next ray
        if (state == S_E_LIGHT) {
            state = S_AFTER_LIGHT;   // (picked up by part A of the next iteration if no shadow ray is needed)
            if (!fEval.IsBlack()) {
                const PbrtGpuLight &light = sc.lights[lightNum];
                if (light.type == PBRTGPU_LIGHT_POINT)
                    pend = fEval * pend * (AbsDot(wiW, ctx.nn) / sPdf);
                else {
                    BxSet bx;
                    MaterialBxDFs(sc, ctx.material, &bx);
                    float bsdfPdf = BSDF_PdfLocal(bx, evalWo, evalWi, BSDF_ALL & ~BSDF_SPECULAR);
                    float weight = PowerHeuristic(1, sPdf, 1, bsdfPdf);
                    pend = fEval * pend * (AbsDot(wiW, ctx.nn) * weight / sPdf);
                }
                ray.o = ctx.p; ray.d = wiW; ray.mint = ctx.rayEpsilon; ray.maxt = visMaxt;
                wantTrace = true; anyHit = true;
                state = S_SHADOW_RESULT;
            }
        } else if (state == S_E_MIS) {
            state = S_AFTER_DIRECT;   // (next iteration)
            if (!fEval.IsBlack() && sPdf > 0.f) {
                const PbrtGpuLight &light = sc.lights[lightNum];
                // sampledType is never specular for the in-scope BxDFs
                float lightPdf = LightPdf(sc, light, ctx.p, wiW);
                if (lightPdf != 0.f) {   // else `return Ld`
                    float weight = PowerHeuristic(1, sPdf, 1, lightPdf);
                    // Li is either Lemit or black (diffuse.h:51-53), so the term can be formed now
                    pend = fEval * Spec(light.L[0], light.L[1], light.L[2]) * AbsDot(wiW, ctx.nn) * weight / sPdf;
                    ray.o = ctx.p; ray.d = wiW; ray.mint = ctx.rayEpsilon; ray.maxt = INF_F;
                    wantTrace = true; anyHit = false;
                    state = S_MIS_RESULT;
                }
            }
        } else if (state == S_E_PATH) {
            // ---- path.cpp:96-121: throughput, Russian roulette, next ray
            state = S_END_SAMPLE;
            if (!fEval.IsBlack()) {
                T = T * (fEval * AbsDot(wiW, ctx.nn) / sPdf);
                ray.o = ctx.p; ray.d = wiW; ray.mint = ctx.rayEpsilon; ray.maxt = INF_F;
                bool alive = true;
                if (bounces > 3) {
                    float continueProbability = minf(.5f, T.y());
                    const float r = rngF(0);
                    mti += 1;
                    if (r > continueProbability) alive = false;
                    else T = T / continueProbability;
                }
                if (alive && bounces != rp.maxDepth) {
                    ++bounces;
                    wantTrace = true; anyHit = false;
                    state = S_HIT_RESULT;
                }
            }
        }

        if (state == S_END_SAMPLE) {
            if (PHASE == 1) {
                const Spec Ls = SanitizeRadiance(L);
                lsBase[s] = make_float4(Ls.r, Ls.g, Ls.b, __int_as_float(code));
                state = S_NEW_SAMPLE;   // next sample index is handed out at the top of the loop
            } else {
                // samplerrenderer.cpp:111-146: sanity checks, then ImageFilm::AddSample
                FilmAdd(a, xPos, yPos, code, SanitizeRadiance(L), acc, direct);
                ++s;
                state = S_NEW_SAMPLE;
            }
        }

        if (state == S_NEW_SAMPLE) {
            if (PHASE == 1) {
                if (fresh) {
                    // LDPixelSample's camera sample + PerspectiveCamera::GenerateRayDifferential (perspective.cpp:81-138)
                    const uint16_t *row = permBase + (size_t)s * PERM_STRIDE;
                    const uint32_t pi = row[0];
                    const float imageX = xPos + VanDerCorput(pi, scr[0]);
                    const float imageY = yPos + Sobol2(pi, scr[1]);
                    code = FilmFootprint(rp, xPos, yPos, imageX, imageY);
                    V3 Pcamera = XfPoint(rp.rasterToCamera, V3(imageX, imageY, 0));
                    V3 dir = Normalize(V3(Pcamera.x, Pcamera.y, Pcamera.z));
                    ray.o = V3(0, 0, 0); ray.d = dir; ray.mint = 0.f; ray.maxt = INF_F;
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
            } else {
                // walk ls[] in sample order: add finished samples, resume the next parked one
                for (int k = 0; k < 64 && state == S_NEW_SAMPLE; ++k) {
                    if (s == spp) {
                        if (!direct) {
                            PbrtGpuPixel *px = a.film + (size_t)(yPos - rp.yPixelStart) * rp.xPixelCount + (xPos - rp.xPixelStart);
                            atomicAdd(&px->Lxyz[0], acc[0]);
                            atomicAdd(&px->Lxyz[1], acc[1]);
                            atomicAdd(&px->Lxyz[2], acc[2]);
                            atomicAdd(&px->weightSum, acc[3]);
                        }
                        a.b.mti[bi] = mti;
                        state = S_DONE;
                        break;
                    }
                    const float4 v = __ldcs(lsBase + s);
                    code = __float_as_int(v.w);
                    if (code & LS_TAIL) {
                        const float4 *r = recBase + (size_t)s * 4;
                        const float4 r0 = __ldcs(r), r1 = __ldcs(r + 1), r2 = __ldcs(r + 2), r3 = __ldcs(r + 3);
                        ray.o = V3(r0.x, r0.y, r0.z); ray.mint = r0.w;
                        ray.d = V3(r1.x, r1.y, r1.z); ray.maxt = r1.w;
                        hit.t = r1.w;
                        T = Spec(r2.x, r2.y, r2.z);
                        hit.prim = __float_as_int(r2.w);
                        hit.b1 = r3.x; hit.b2 = r3.y; hit.maxtIn = r3.z;
                        L = Spec(v.x, v.y, v.z);
                        code &= ~LS_TAIL;
                        bounces = 3;
                        found = true;
                        state = S_HIT_RESULT;   // picked up (after the RNG top-up) next iteration
                    } else {
                        FilmAdd(a, xPos, yPos, code, Spec(v.x, v.y, v.z), acc, direct);
                        ++s;
                    }
                }
            }
        }

        // ---- the single traversal site
        if (wantTrace) {
            if (STATS) { if (anyHit) ++nShadow; else ++nClosest; }
            found = Traverse<STATS>(sc, ray, anyHit, &hit, &st);
        }
    }

    if (STATS) {
        unsigned long long v[5] = {nClosest, nShadow, st.nodes, st.prims, nShadedN};
        for (int k = 0; k < 5; ++k) {
            unsigned long long x = v[k];
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(FULL, x, o);
            v[k] = x;
        }
        if (lane == 0) {
            atomicAdd(&a.stats->closestRays, v[0]);
            atomicAdd(&a.stats->shadowRays, v[1]);
            atomicAdd(&a.stats->nodes, v[2]);
            atomicAdd(&a.stats->prims, v[3]);
            atomicAdd(&a.stats->shadedN, v[4]);
        }
    }
}

// ------------------------------------------------------------- ray-batch kernels
template <bool ANY>
__global__ void k_intersect(DScene sc, const PbrtGpuRay *rays, int n, PbrtGpuHit *hits, uint8_t *occ) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray ray;
    ray.o = V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]);
    ray.d = V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
    ray.mint = rays[i].mint; ray.maxt = rays[i].maxt;
    Hit hit;
    TraceStats st;
    const bool found = Traverse<false>(sc, ray, ANY, &hit, &st);
    if (ANY) { occ[i] = found ? 1 : 0; return; }
    PbrtGpuHit h;
    memset(&h, 0, sizeof(h));
    h.t = ray.maxt;
    if (found) {
        DiffGeom dg;
        float eps;
        int mat, al;
        uint4 tri;
        uint32_t ptype;
        Ray r0 = ray;
        HitGeometry(sc, r0, hit, &dg, &eps, &mat, &al, &tri, &ptype);
        h.hit = 1; h.primitive = (uint32_t)hit.prim; h.b1 = hit.b1; h.b2 = hit.b2;
        h.p[0] = dg.p.x; h.p[1] = dg.p.y; h.p[2] = dg.p.z;
        h.n[0] = dg.nn.x; h.n[1] = dg.nn.y; h.n[2] = dg.nn.z;
    }
    hits[i] = h;
}

// ======================================================================= host
namespace {

thread_local std::string g_lastError;

int Fail(int code, const std::string &msg) {
    g_lastError = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return Fail(PBRTGPU_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));          \
    } while (0)

template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t Ensure(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t Upload(const T *src, size_t count) {
        cudaError_t e = Ensure(count ? count : 1);
        if (e != cudaSuccess) return e;
        return count ? cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice) : cudaSuccess;
    }
};

}  // namespace

struct PbrtGpuScene {
    int device = 0;
    DScene d;
    DevBuf<float4> nodes, prims;
    DevBuf<uint4> tris;
    DevBuf<PbrtGpuMesh> meshes;
    DevBuf<float> N, uv;
    DevBuf<PbrtGpuQuadric> quadrics;
    DevBuf<PbrtGpuXform> xforms;
    DevBuf<PbrtGpuMaterial> materials;
    DevBuf<PbrtGpuMeasured> measured;
    DevBuf<PbrtGpuKdNode> kdNodes;
    DevBuf<PbrtGpuBrdfSample> kdData;
    DevBuf<PbrtGpuLight> lights;
    DevBuf<PbrtGpuLightShape> lightShapes;
    // per-render scratch, kept between calls
    DevBuf<TaskRec> tasks;
    DevBuf<uint16_t> perm;
    DevBuf<uint32_t> scr, rngCur, rngNext;
    DevBuf<int> mti;
    DevBuf<float4> ls, rec;
    DevBuf<PbrtGpuPixel> film;
    DevBuf<DeviceStats> stats;
    DevBuf<PbrtGpuRay> rays;
    DevBuf<PbrtGpuHit> hits;
    DevBuf<uint8_t> occ;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    bool collectStats = false;
};

namespace {

// host copies of the window arithmetic: Sampler::ComputeSubWindow (core/sampler.cpp:55-74),
// ImageFilm extent (film/image.cpp:48-51,157-166).  Plain float arithmetic, same on any IEEE host.
inline int HFloor2Int(float v) { return (int)floorf(v); }
inline int HCeil2Int(float v) { return (int)ceilf(v); }
inline float HLerp(float t, float v1, float v2) { return (1.f - t) * v1 + t * v2; }

struct FilmExtent {
    int xPixelStart, yPixelStart, xPixelCount, yPixelCount;
    int sx0, sx1, sy0, sy1;
};
FilmExtent ComputeExtent(const PbrtGpuRenderDesc &rd) {
    FilmExtent e;
    e.xPixelStart = HCeil2Int(rd.xResolution * rd.cropWindow[0]);
    e.xPixelCount = std::max(1, HCeil2Int(rd.xResolution * rd.cropWindow[1]) - e.xPixelStart);
    e.yPixelStart = HCeil2Int(rd.yResolution * rd.cropWindow[2]);
    e.yPixelCount = std::max(1, HCeil2Int(rd.yResolution * rd.cropWindow[3]) - e.yPixelStart);
    e.sx0 = HFloor2Int(e.xPixelStart + 0.5f - rd.filterXWidth);
    e.sx1 = HCeil2Int(e.xPixelStart - 0.5f + e.xPixelCount + rd.filterXWidth);
    e.sy0 = HFloor2Int(e.yPixelStart + 0.5f - rd.filterYWidth);
    e.sy1 = HCeil2Int(e.yPixelStart - 0.5f + e.yPixelCount + rd.filterYWidth);
    return e;
}
void ComputeSubWindow(const FilmExtent &e, int num, int count, int *x0, int *x1, int *y0, int *y1) {
    int dx = e.sx1 - e.sx0, dy = e.sy1 - e.sy0;
    int nx = count, ny = 1;
    while ((nx & 0x1) == 0 && 2 * dx * ny < dy * nx) { nx >>= 1; ny <<= 1; }
    int xo = num % nx, yo = num / nx;
    float tx0 = float(xo) / float(nx), tx1 = float(xo + 1) / float(nx);
    float ty0 = float(yo) / float(ny), ty1 = float(yo + 1) / float(ny);
    *x0 = HFloor2Int(HLerp(tx0, e.sx0, e.sx1));
    *x1 = HFloor2Int(HLerp(tx1, e.sx0, e.sx1));
    *y0 = HFloor2Int(HLerp(ty0, e.sy0, e.sy1));
    *y1 = HFloor2Int(HLerp(ty1, e.sy0, e.sy1));
}

int ValidateRenderDesc(const PbrtGpuRenderDesc &rd) {
    if (rd.xResolution <= 0 || rd.yResolution <= 0) return Fail(PBRTGPU_EINVAL, "bad film resolution");
    const int spp = rd.samplesPerPixel;
    if (spp <= 0 || (spp & (spp - 1))) return Fail(PBRTGPU_EINVAL, "samplesPerPixel must be a power of two (LDSampler)");
    if (spp > 8192) return Fail(PBRTGPU_EUNSUPPORTED, "samplesPerPixel > 8192 does not fit the table kernel's shared memory");
    if (rd.filterXWidth > 0.5f || rd.filterYWidth > 0.5f)
        return Fail(PBRTGPU_EUNSUPPORTED, "filters wider than the 0.5 box need cross-task splats (out of scope)");
    if (rd.nTasks <= 0 || rd.taskBegin < 0 || rd.taskEnd > rd.nTasks || rd.taskBegin > rd.taskEnd)
        return Fail(PBRTGPU_EINVAL, "bad task range");
    if (rd.maxDepth < 0) return Fail(PBRTGPU_EINVAL, "bad maxDepth");
    if (rd.shardCount > 1 && (rd.shardIndex < 0 || rd.shardIndex >= rd.shardCount))
        return Fail(PBRTGPU_EINVAL, "bad shard index");
    return PBRTGPU_OK;
}

struct TableLaunch {
    int G, wpb;
    size_t smemBytes;
};
TableLaunch PlanTables(int spp) {
    TableLaunch t;
    const size_t perArray = (size_t)4 * (spp + 2);   // T + P, u16 each
    const size_t budget = 56 * 1024;
    size_t g = (budget - MT_N * 4) / perArray;
    t.G = (int)std::max<size_t>(1, std::min<size_t>(N_ARRAYS, g));
    const size_t perWarp = MT_N * 4 + (size_t)t.G * perArray;
    t.wpb = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / perWarp));
    t.smemBytes = perWarp * t.wpb;
    return t;
}

template <int WPB> cudaError_t LaunchTables(const BatchArgs &b, int spp, const TableLaunch &tl, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(k_tables<WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl.smemBytes);
    if (e != cudaSuccess) return e;
    const int blocks = (b.nBatch + WPB - 1) / WPB;
    k_tables<WPB><<<blocks, 32 * WPB, tl.smemBytes, stream>>>(b, spp, tl.G);
    return cudaGetLastError();
}
cudaError_t LaunchTablesAny(const BatchArgs &b, int spp, const TableLaunch &tl, cudaStream_t stream) {
    switch (tl.wpb) {
    case 4: return LaunchTables<4>(b, spp, tl, stream);
    case 3: return LaunchTables<3>(b, spp, tl, stream);
    case 2: return LaunchTables<2>(b, spp, tl, stream);
    default: return LaunchTables<1>(b, spp, tl, stream);
    }
}

size_t ScratchBudgetBytes() {
    const char *env = getenv("PBRTGPU_SCRATCH_GB");
    double gb = env ? atof(env) : 24.0;
    if (gb < 0.25) gb = 0.25;
    return (size_t)(gb * 1024.0 * 1024.0 * 1024.0);
}

}  // namespace

extern "C" {

int pbrtgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int pbrtgpu_set_device(int ordinal) {
    CUDA_TRY(cudaSetDevice(ordinal));
    return PBRTGPU_OK;
}

const char *pbrtgpu_last_error(void) { return g_lastError.c_str(); }

int pbrtgpu_scene_create(const PbrtGpuSceneDesc *desc, PbrtGpuScene **out) {
    if (!desc || !out) return Fail(PBRTGPU_EINVAL, "null argument");
    *out = nullptr;
    if (pbrtgpu_device_count() <= 0)
        return Fail(PBRTGPU_ENODEVICE, "no CUDA device visible; libpbrtgpu has no CPU fallback");
    if (desc->nNodes == 0 || desc->nPrimitives == 0) return Fail(PBRTGPU_EINVAL, "empty scene");
    if (desc->nMaterials > 65535 || desc->nLights > 65534) return Fail(PBRTGPU_EUNSUPPORTED, "too many materials/lights");
    for (uint32_t i = 0; i < desc->nLights; ++i)
        if (desc->lights[i].type == PBRTGPU_LIGHT_DIFFUSE_AREA && desc->lights[i].nShapes != 1)
            return Fail(PBRTGPU_EUNSUPPORTED, "area lights over more than one shape are out of scope");
    for (uint32_t i = 0; i < desc->nMaterials; ++i) {
        const PbrtGpuMaterial &m = desc->materials[i];
        if (m.type > PBRTGPU_MAT_MEASURED) return Fail(PBRTGPU_EINVAL, "unknown material type");
        if (m.type == PBRTGPU_MAT_MATTE && m.sigma != 0.f) return Fail(PBRTGPU_EUNSUPPORTED, "Oren-Nayar (sigma != 0) is out of scope");
        if (m.type == PBRTGPU_MAT_MEASURED && (m.measured < 0 || (uint32_t)m.measured >= desc->nMeasured))
            return Fail(PBRTGPU_EINVAL, "measured material without data");
    }
    std::unique_ptr<PbrtGpuScene> s(new PbrtGpuScene);
    CUDA_TRY(cudaGetDevice(&s->device));

    // nodes: 32-byte LinearBVHNode -> two float4 (bit copy; the byte pair nPrimitives/axis lands in .w)
    {
        std::vector<float4> h(2 * (size_t)desc->nNodes);
        for (uint32_t i = 0; i < desc->nNodes; ++i) {
            const PbrtGpuBVHNode &n = desc->nodes[i];
            uint32_t meta = (uint32_t)n.nPrimitives | ((uint32_t)n.axis << 8);
            float off, mf;
            memcpy(&off, &n.offset, 4);
            memcpy(&mf, &meta, 4);
            h[2 * i] = make_float4(n.bmin[0], n.bmin[1], n.bmin[2], n.bmax[0]);
            h[2 * i + 1] = make_float4(n.bmax[1], n.bmax[2], off, mf);
        }
        CUDA_TRY(s->nodes.Upload(h.data(), h.size()));
    }
    // primitives: pre-gathered 48-byte records in leaf order
    {
        std::vector<float4> h(3 * (size_t)desc->nPrimitives);
        for (uint32_t i = 0; i < desc->nPrimitives; ++i) {
            const PbrtGpuPrimitive &p = desc->primitives[i];
            if (p.material < 0 || (uint32_t)p.material >= desc->nMaterials) return Fail(PBRTGPU_EINVAL, "primitive without material");
            float v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (p.type == PBRTGPU_PRIM_TRIANGLE) {
                if (p.index >= desc->nTriangles) return Fail(PBRTGPU_EINVAL, "triangle index out of range");
                const PbrtGpuTriangle &t = desc->triangles[p.index];
                for (int k = 0; k < 3; ++k) {
                    if (t.v[k] >= desc->nVertices) return Fail(PBRTGPU_EINVAL, "vertex index out of range");
                    memcpy(v + 3 * k, desc->P + 3 * (size_t)t.v[k], 12);
                }
            } else if (p.index >= desc->nQuadrics) return Fail(PBRTGPU_EINVAL, "quadric index out of range");
            uint32_t w0 = p.type, w1 = p.index, w2 = (uint32_t)p.material | ((uint32_t)(p.areaLight + 1) << 16);
            float f0, f1, f2;
            memcpy(&f0, &w0, 4); memcpy(&f1, &w1, 4); memcpy(&f2, &w2, 4);
            h[3 * i] = make_float4(v[0], v[1], v[2], f0);
            h[3 * i + 1] = make_float4(v[3], v[4], v[5], f1);
            h[3 * i + 2] = make_float4(v[6], v[7], v[8], f2);
        }
        CUDA_TRY(s->prims.Upload(h.data(), h.size()));
    }
    {
        std::vector<uint4> h(desc->nTriangles);
        for (uint32_t i = 0; i < desc->nTriangles; ++i)
            h[i] = make_uint4(desc->triangles[i].v[0], desc->triangles[i].v[1], desc->triangles[i].v[2], desc->triangles[i].mesh);
        CUDA_TRY(s->tris.Upload(h.data(), h.size()));
    }
    CUDA_TRY(s->meshes.Upload(desc->meshes, desc->nMeshes));
    CUDA_TRY(s->N.Upload(desc->N, desc->N ? 3 * (size_t)desc->nVertices : 0));
    CUDA_TRY(s->uv.Upload(desc->uv, desc->uv ? 2 * (size_t)desc->nVertices : 0));
    CUDA_TRY(s->quadrics.Upload(desc->quadrics, desc->nQuadrics));
    CUDA_TRY(s->xforms.Upload(desc->xforms, desc->nXforms));
    CUDA_TRY(s->materials.Upload(desc->materials, desc->nMaterials));
    CUDA_TRY(s->measured.Upload(desc->measured, desc->nMeasured));
    CUDA_TRY(s->kdNodes.Upload(desc->kdNodes, desc->nKdNodes));
    CUDA_TRY(s->kdData.Upload(desc->kdData, desc->nKdNodes));
    CUDA_TRY(s->lights.Upload(desc->lights, desc->nLights));
    CUDA_TRY(s->lightShapes.Upload(desc->lightShapes, desc->nLightShapes));
    for (uint32_t i = 0; i < desc->nMeshes; ++i) {
        if ((desc->meshes[i].flags & PBRTGPU_MESH_HAS_N) && !desc->N) return Fail(PBRTGPU_EINVAL, "mesh flags N but desc.N is null");
        if ((desc->meshes[i].flags & PBRTGPU_MESH_HAS_UV) && !desc->uv) return Fail(PBRTGPU_EINVAL, "mesh flags uv but desc.uv is null");
    }
    DScene &d = s->d;
    d.nodes = s->nodes.p; d.prims = s->prims.p; d.tris = s->tris.p; d.meshes = s->meshes.p;
    d.N = s->N.p; d.uv = s->uv.p; d.quadrics = s->quadrics.p; d.xforms = s->xforms.p;
    d.materials = s->materials.p; d.measured = s->measured.p; d.kdNodes = s->kdNodes.p; d.kdData = s->kdData.p;
    d.lights = s->lights.p; d.lightShapes = s->lightShapes.p;
    d.nNodes = desc->nNodes; d.nPrims = desc->nPrimitives; d.nLights = desc->nLights;
    for (int i = 0; i < 5; ++i) CUDA_TRY(cudaEventCreate(&s->ev[i]));
    s->collectStats = getenv("PBRTGPU_STATS") != nullptr;
    *out = s.release();
    return PBRTGPU_OK;
}

void pbrtgpu_scene_destroy(PbrtGpuScene *scene) {
    if (!scene) return;
    for (int i = 0; i < 5; ++i)
        if (scene->ev[i]) cudaEventDestroy(scene->ev[i]);
    delete scene;
}

// counting pass switch: PbrtGpuStats.nodesVisited etc. are only filled when the
// scene was created with PBRTGPU_STATS set or via pbrtgpu_set_stats(scene, 1); the
// timed path runs the kernel instantiation without counters.
int pbrtgpu_set_stats(PbrtGpuScene *scene, int on) {
    if (!scene) return Fail(PBRTGPU_EINVAL, "null scene");
    scene->collectStats = on != 0;
    return PBRTGPU_OK;
}

int pbrtgpu_render_device(PbrtGpuScene *scene, const PbrtGpuRenderDesc *rd, void *d_pixels, void *streamv,
                          PbrtGpuStats *stats) {
    if (!scene || !rd || !d_pixels) return Fail(PBRTGPU_EINVAL, "null argument");
    int rc = ValidateRenderDesc(*rd);
    if (rc != PBRTGPU_OK) return rc;
    cudaStream_t stream = (cudaStream_t)streamv;
    const FilmExtent ext = ComputeExtent(*rd);
    const int spp = rd->samplesPerPixel;

    // ---- non-empty tasks of [taskBegin, taskEnd)  (GetSubSampler returns NULL for empty tiles)
    std::vector<TaskRec> tasks;
    int maxPixels = 0;
    uint64_t totalPixels = 0;
    const int shardCount = std::max(1, rd->shardCount), shardBlock = std::max(1, rd->shardBlock);
    for (int t = rd->taskBegin; t < rd->taskEnd; ++t) {
        if (shardCount > 1 && (t / shardBlock) % shardCount != rd->shardIndex) continue;
        TaskRec tr;
        tr.taskNum = t;
        ComputeSubWindow(ext, t, rd->nTasks, &tr.x0, &tr.x1, &tr.y0, &tr.y1);
        if (tr.x0 == tr.x1 || tr.y0 == tr.y1) continue;
        tasks.push_back(tr);
        const int npx = (tr.x1 - tr.x0) * (tr.y1 - tr.y0);
        maxPixels = std::max(maxPixels, npx);
        totalPixels += (uint64_t)npx;
    }
    PbrtGpuPixel *film = (PbrtGpuPixel *)d_pixels;
    const size_t nFilm = (size_t)ext.xPixelCount * ext.yPixelCount;
    if (stats) CUDA_TRY(cudaEventRecord(scene->ev[0], stream));
    CUDA_TRY(cudaMemsetAsync(film, 0, nFilm * sizeof(PbrtGpuPixel), stream));

    RenderParams rp;
    memcpy(rp.rasterToCamera, rd->rasterToCamera, sizeof(rp.rasterToCamera));
    memcpy(rp.cameraToWorld, rd->cameraToWorld, sizeof(rp.cameraToWorld));
    rp.lensRadius = rd->lensRadius; rp.focalDistance = rd->focalDistance;
    rp.filterXWidth = rd->filterXWidth; rp.filterYWidth = rd->filterYWidth;
    rp.xPixelStart = ext.xPixelStart; rp.yPixelStart = ext.yPixelStart;
    rp.xPixelCount = ext.xPixelCount; rp.yPixelCount = ext.yPixelCount;
    rp.spp = spp; rp.maxDepth = rd->maxDepth;

    // per task: permutation tables (64 B/sample) + finished radiance (16 B/sample) + parked
    // bounce-3 vertices (64 B/sample, sparse) + scrambles + two MT19937 blocks
    const size_t perTask = (size_t)spp * (PERM_STRIDE * 2 + 16 + 64) + SCR_STRIDE * 4 + 2 * MT_N * 4 + 4 + sizeof(TaskRec);
    size_t maxBatch = std::max<size_t>(1024, ScratchBudgetBytes() / perTask);
    maxBatch = std::min<size_t>(maxBatch, std::max<size_t>(tasks.size(), 1));
    CUDA_TRY(scene->tasks.Ensure(maxBatch));
    CUDA_TRY(scene->perm.Ensure(maxBatch * spp * PERM_STRIDE));
    CUDA_TRY(scene->scr.Ensure(maxBatch * SCR_STRIDE));
    CUDA_TRY(scene->rngCur.Ensure(maxBatch * MT_N));
    CUDA_TRY(scene->rngNext.Ensure(maxBatch * MT_N));
    CUDA_TRY(scene->mti.Ensure(maxBatch));
    CUDA_TRY(scene->ls.Ensure(maxBatch * spp));
    CUDA_TRY(scene->rec.Ensure(maxBatch * spp * 4));
    CUDA_TRY(scene->stats.Ensure(1));
    const bool wantCounters = stats && scene->collectStats;
    if (wantCounters) CUDA_TRY(cudaMemsetAsync(scene->stats.p, 0, sizeof(DeviceStats), stream));

    const TableLaunch tl = PlanTables(spp);
    uint64_t launches = 0;
    float tableMs = 0.f, traceMs = 0.f, tailMs = 0.f;
    for (size_t begin = 0; begin < tasks.size(); begin += maxBatch) {
        const size_t nb = std::min(maxBatch, tasks.size() - begin);
        int rounds = 0;
        for (size_t i = 0; i < nb; ++i)
            rounds = std::max(rounds, (tasks[begin + i].x1 - tasks[begin + i].x0) * (tasks[begin + i].y1 - tasks[begin + i].y0));
        CUDA_TRY(cudaMemcpyAsync(scene->tasks.p, tasks.data() + begin, nb * sizeof(TaskRec), cudaMemcpyHostToDevice, stream));
        BatchArgs b;
        b.tasks = scene->tasks.p; b.nBatch = (int)nb; b.round = 0;
        b.perm = scene->perm.p; b.scr = scene->scr.p; b.rngCur = scene->rngCur.p; b.rngNext = scene->rngNext.p;
        b.mti = scene->mti.p; b.ls = scene->ls.p; b.rec = scene->rec.p;
        for (int r = 0; r < rounds; ++r) {
            b.round = r;
            if (stats) CUDA_TRY(cudaEventRecord(scene->ev[2], stream));
            CUDA_TRY(LaunchTablesAny(b, spp, tl, stream));
            if (stats) CUDA_TRY(cudaEventRecord(scene->ev[3], stream));
            PathArgs pa;
            pa.sc = scene->d; pa.rp = rp; pa.b = b; pa.film = film; pa.stats = scene->stats.p;
            pa.directFilm = rounds > 1 ? 1 : 0;
            { const char *e = getenv("PBRTGPU_LOCKSTEP"); pa.lockstep = e ? atoi(e) : 1; }
            const int blocks1 = (int)((nb + 3) / 4);       // phase 1: one warp per task-pixel
            const int blocks2 = (int)((nb + 127) / 128);   // phase 2: one thread per task-pixel
            if (wantCounters) k_path<1, true><<<blocks1, 128, 0, stream>>>(pa);
            else k_path<1, false><<<blocks1, 128, 0, stream>>>(pa);
            CUDA_TRY(cudaGetLastError());
            if (stats) CUDA_TRY(cudaEventRecord(scene->ev[4], stream));
            if (wantCounters) k_path<2, true><<<blocks2, 128, 0, stream>>>(pa);
            else k_path<2, false><<<blocks2, 128, 0, stream>>>(pa);
            CUDA_TRY(cudaGetLastError());
            launches += 3;
            if (stats) {
                CUDA_TRY(cudaEventRecord(scene->ev[1], stream));
                CUDA_TRY(cudaEventSynchronize(scene->ev[1]));
                float ms = 0.f;
                CUDA_TRY(cudaEventElapsedTime(&ms, scene->ev[2], scene->ev[3]));
                tableMs += ms;
                CUDA_TRY(cudaEventElapsedTime(&ms, scene->ev[3], scene->ev[4]));
                traceMs += ms;
                CUDA_TRY(cudaEventElapsedTime(&ms, scene->ev[4], scene->ev[1]));
                tailMs += ms;
            }
        }
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        CUDA_TRY(cudaEventRecord(scene->ev[1], stream));
        CUDA_TRY(cudaEventSynchronize(scene->ev[1]));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, scene->ev[0], scene->ev[1]));
        stats->renderMs = ms;
        stats->tableMs = tableMs;
        stats->traceMs = traceMs;
        stats->tailMs = tailMs;
        stats->cameraSamples = totalPixels * (uint64_t)spp;
        stats->kernelLaunches = launches;
        if (wantCounters) {
            DeviceStats ds;
            CUDA_TRY(cudaMemcpy(&ds, scene->stats.p, sizeof(ds), cudaMemcpyDeviceToHost));
            stats->closestRays = ds.closestRays; stats->shadowRays = ds.shadowRays;
            stats->nodesVisited = ds.nodes; stats->primTests = ds.prims;
        }
    }
    return PBRTGPU_OK;
}

int pbrtgpu_render(PbrtGpuScene *scene, const PbrtGpuRenderDesc *rd, PbrtGpuPixel *pixels, PbrtGpuStats *stats) {
    if (!scene || !rd || !pixels) return Fail(PBRTGPU_EINVAL, "null argument");
    int rc = ValidateRenderDesc(*rd);
    if (rc != PBRTGPU_OK) return rc;
    const FilmExtent ext = ComputeExtent(*rd);
    const size_t nFilm = (size_t)ext.xPixelCount * ext.yPixelCount;
    CUDA_TRY(scene->film.Ensure(nFilm));
    PbrtGpuStats local;
    rc = pbrtgpu_render_device(scene, rd, scene->film.p, nullptr, stats ? stats : &local);
    if (rc != PBRTGPU_OK) return rc;
    CUDA_TRY(cudaMemcpy(pixels, scene->film.p, nFilm * sizeof(PbrtGpuPixel), cudaMemcpyDeviceToHost));
    return PBRTGPU_OK;
}

int pbrtgpu_intersect(PbrtGpuScene *scene, const PbrtGpuRay *rays, int n, PbrtGpuHit *hits) {
    if (!scene || (!rays && n) || (!hits && n) || n < 0) return Fail(PBRTGPU_EINVAL, "bad argument");
    if (n == 0) return PBRTGPU_OK;
    CUDA_TRY(scene->rays.Upload(rays, n));
    CUDA_TRY(scene->hits.Ensure(n));
    k_intersect<false><<<(n + 127) / 128, 128>>>(scene->d, scene->rays.p, n, scene->hits.p, nullptr);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(hits, scene->hits.p, (size_t)n * sizeof(PbrtGpuHit), cudaMemcpyDeviceToHost));
    return PBRTGPU_OK;
}

int pbrtgpu_intersect_p(PbrtGpuScene *scene, const PbrtGpuRay *rays, int n, uint8_t *occluded) {
    if (!scene || (!rays && n) || (!occluded && n) || n < 0) return Fail(PBRTGPU_EINVAL, "bad argument");
    if (n == 0) return PBRTGPU_OK;
    CUDA_TRY(scene->rays.Upload(rays, n));
    CUDA_TRY(scene->occ.Ensure(n));
    k_intersect<true><<<(n + 127) / 128, 128>>>(scene->d, scene->rays.p, n, nullptr, scene->occ.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(occluded, scene->occ.p, (size_t)n, cudaMemcpyDeviceToHost));
    return PBRTGPU_OK;
}

int pbrtgpu_ld_samples(const PbrtGpuRenderDesc *rd, int taskNum, float *out) {
    if (!rd || !out) return Fail(PBRTGPU_EINVAL, "null argument");
    int rc = ValidateRenderDesc(*rd);
    if (rc != PBRTGPU_OK) return rc;
    if (pbrtgpu_device_count() <= 0) return Fail(PBRTGPU_ENODEVICE, "no CUDA device visible; no CPU fallback");
    if (taskNum < 0 || taskNum >= rd->nTasks) return Fail(PBRTGPU_EINVAL, "taskNum out of range");
    const FilmExtent ext = ComputeExtent(*rd);
    TaskRec tr;
    tr.taskNum = taskNum;
    ComputeSubWindow(ext, taskNum, rd->nTasks, &tr.x0, &tr.x1, &tr.y0, &tr.y1);
    if (tr.x0 == tr.x1 || tr.y0 == tr.y1) return Fail(PBRTGPU_EINVAL, "task has an empty pixel window");
    const int spp = rd->samplesPerPixel;
    DevBuf<TaskRec> tasks;
    DevBuf<uint16_t> perm;
    DevBuf<uint32_t> scr, cur, next;
    DevBuf<int> mti;
    DevBuf<float> dout;
    CUDA_TRY(tasks.Upload(&tr, 1));
    CUDA_TRY(perm.Ensure((size_t)spp * PERM_STRIDE));
    CUDA_TRY(scr.Ensure(SCR_STRIDE));
    CUDA_TRY(cur.Ensure(MT_N));
    CUDA_TRY(next.Ensure(MT_N));
    CUDA_TRY(mti.Ensure(1));
    CUDA_TRY(dout.Ensure((size_t)spp * 37));
    BatchArgs b;
    b.tasks = tasks.p; b.nBatch = 1; b.round = 0; b.perm = perm.p; b.scr = scr.p; b.rngCur = cur.p; b.rngNext = next.p;
    b.mti = mti.p;
    const TableLaunch tl = PlanTables(spp);
    CUDA_TRY(LaunchTablesAny(b, spp, tl, nullptr));
    k_expand_samples<<<(spp + 127) / 128, 128>>>(b, spp, rd->shutterOpen, rd->shutterClose, dout.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, dout.p, (size_t)spp * 37 * sizeof(float), cudaMemcpyDeviceToHost));
    return PBRTGPU_OK;
}

}  // extern "C"
