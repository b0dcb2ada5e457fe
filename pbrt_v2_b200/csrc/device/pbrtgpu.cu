// pbrtgpu.cu -- CUDA kernels and C ABI (include/pbrtgpu.h) of the B200 path-tracing
// hot path.  sm_100a only; no CPU fallback anywhere in this file.
//
// Work decomposition:
//   unit        = one reference render task (SamplerRendererTask, its own MT19937 stream seeded with taskNum,
//                 samplerrenderer.cpp:72-73); a batch of tasks is processed pixel by pixel ("rounds")
//   k_tables / k_tables_general
//               = one WARP per task: the LD sample tables of the task's current pixel (LDPixelSample) and the task's
//                 RNG position after them
//   phase 1     = everything that draws from the tables only (PathIntegrator: bounces 0..2; DirectLightingIntegrator:
//                 all of Li): a wavefront over a pool of path slots -- k_logic (state machine, path_body.inc),
//                 k_eval (measured-BRDF kd-tree queries), k_trace (persistent BVH walk) -- with one request queue per
//                 resume state.  PathIntegrator samples that reach bounce 3 are parked (`rec`)
//   phase 2     = k_path<2>: one THREAD per task resumes the parked samples in sample order with the task's RNG (the
//                 offset of one tail depends on how many draws the previous one consumed) and adds all samples to the
//                 film in order; DirectLightingIntegrator: k_film_direct (film + the RNG draws of SpecularReflect/Transmit)
//   lanes       = batches are dealt to concurrent lanes (stream + scratch + host thread), see pbrtgpu_render_device
//   k_path<1>   = the megakernel form of phase 1 (one warp per pixel), kept as the A/B baseline (PBRTGPU_WAVEFRONT=0)
#include "dshade.cuh"
#include "drng.cuh"
#include "dpath.cuh"
#include "dtables.cuh"
#include "kcommon.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace dev;

// ===================================================================== device
__global__ void k_expand_samples(BatchArgs a, int spp, float shutterOpen, float shutterClose, SampleLayoutDev L,
                                 int floatsPerSample, float *out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= spp) return;
    ExpandSample(a, s, shutterOpen, shutterClose, L, floatsPerSample, out);
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
// resident blocks per SM the register allocator must allow (A/B knobs; the defaults are the measured best on C2 at
// 256 spp: k_logic 4 -> 5: 221 -> 214 Msamples/s; k_path 4 -> 5: 220; k_trace unbounded (79 regs, 6 blocks) -> 7 (72 regs)
// with k_eval -> 8: 221 -> 232; k_trace -> 8 (64 regs): 234, within noise of 7)
#ifndef K_PATH_MIN_BLOCKS
#define K_PATH_MIN_BLOCKS 4
#endif
#ifndef K_LOGIC_MIN_BLOCKS
#define K_LOGIC_MIN_BLOCKS 4
#endif
#ifndef K_EVAL_MIN_BLOCKS
#define K_EVAL_MIN_BLOCKS 8
#endif
struct PathArgs {
    DScene sc;
    RenderParams rp;
    BatchArgs b;
    PbrtGpuPixel *film;
    DeviceStats *stats;
    int directFilm;   // 1: every AddSample goes straight to the film (multi-pixel tasks)
    int lockstep;     // phase 1: hand out camera samples 32 at a time (see k_path)
};

// ImageFilm::AddSample, film/image.cpp:97-137, box filter (all table weights are 1)
DEV void FilmAdd(const RenderParams &rp, PbrtGpuPixel *film, int ownX, int ownY, int code, const Spec &L, float acc[4],
                 bool direct) {
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
                PbrtGpuPixel *px = film + (size_t)(y - rp.yPixelStart) * rp.xPixelCount + (x - rp.xPixelStart);
                atomicAdd(&px->Lxyz[0], xyz[0]);
                atomicAdd(&px->Lxyz[1], xyz[1]);
                atomicAdd(&px->Lxyz[2], xyz[2]);
                atomicAdd(&px->weightSum, 1.f);
            }
        }
}

// one sample into the film at the sites below (xPos, yPos, acc, direct, rp, a in scope): through the weight table when the
// render has one (PBRT_FILTER builds), else the 0.5 box's footprint bits
#if PBRT_FILTER
#define FILM_ADD(row, scr, code, L)                                                                     \
    do {                                                                                                \
        if (rp.filterTable) FilmAddFilteredRow(rp, a.film, (row), (scr), xPos, yPos, (L));              \
        else FilmAdd(rp, a.film, xPos, yPos, (code), (L), acc, direct);                                 \
    } while (0)
#else
#define FILM_ADD(row, scr, code, L) FilmAdd(rp, a.film, xPos, yPos, (code), (L), acc, direct)
#endif

template <int PHASE, bool STATS, bool ANIM>
// 4 CTAs/SM (<= 128 registers): measured +9 % over the unconstrained 168-register build
__global__ void __launch_bounds__(128, K_PATH_MIN_BLOCKS) k_path(const __grid_constant__ PathArgs a) {
    __shared__ uint32_t smt[4][MT_N];
    const DScene &sc = a.sc;
    const RenderParams &rp = a.rp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    const int ti = PHASE == 1 ? (int)(blockIdx.x * 4 + warp) : (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int spp = rp.spp;
    // phase 2 serves the tasks that parked samples first (stable partition: image order is kept inside both
    // groups, so neighbouring lanes still walk neighbouring pixels); warps of tail-free pixels only sum up.
    // (Sorting by the parked-sample COUNT was measured too: +15 % on bunny, -37 % on killeroo, where nearly
    // every pixel has tails and the sort only destroys ray coherence.)
    const int bi = (PHASE == 2 && a.b.order && ti < a.b.nBatch) ? (int)a.b.order[ti] : ti;

    bool valid = ti < a.b.nBatch;
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
    ray.o = V3(0, 0, 0); ray.d = V3(0, 0, 1); ray.mint = 0.f; ray.maxt = 0.f; ray.time = 0.f;
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
#if PBRT_SPECULAR
    bool specularBounce = false;          // path.cpp:53; travels with a parked sample as LS_SPEC
    int dlDepth = 0, dlStage = 0, dlNodes = 0;   // (DirectLightingIntegrator's recursion: unused, PATH_INTEG == 0 here)
    bool dlReturn = false;
    DlFrame dlFrames[1];
    (void)dlStage; (void)dlNodes; (void)dlReturn; (void)dlFrames;
#endif
    int dlI = 0, dlJ = 0;                 // (DirectLightingIntegrator state of path_body.inc: unused, PATH_INTEG == 0 here)
    Spec dlSum(0.f), dlAll(0.f);
    const int rowStride = PERM_STRIDE;
    bool ctxDirty = false;
    (void)ctxDirty;
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

#define PATH_FEVAL_IN Spec(0.f)
#define PATH_EVALWO_IN V3(0, 0, 1)
#define PATH_EVALWI_IN V3(0, 0, 1)
#define PATH_DEFER_MEASURED false
#define PATH_ANIM ANIM
#define PATH_INTEG 0
#define PATH_COUNT_TAIL() (atomicAdd(a.b.tailCount + bi, 1u), a.b.tailMask ? (void)atomicOr(a.b.tailMask + (size_t)bi * a.b.maskWords + (s >> 5), 1u << (s & 31)) : (void)0)
#define PATH_DL_FRAME(i) dlFrames[((i), 0)]
#include "path_body.inc"
#undef PATH_DL_FRAME
#undef PATH_COUNT_TAIL
#undef PATH_FEVAL_IN
#undef PATH_EVALWO_IN
#undef PATH_EVALWI_IN
#undef PATH_DEFER_MEASURED
#undef PATH_ANIM
#undef PATH_INTEG

        if (state == S_END_SAMPLE) {
            if (PHASE == 1) {
                const Spec Ls = SanitizeRadiance(L);
                lsBase[s] = make_float4(Ls.r, Ls.g, Ls.b, __int_as_float(code));
                state = S_NEW_SAMPLE;   // next sample index is handed out at the top of the loop
            } else {
                // samplerrenderer.cpp:111-146: sanity checks, then ImageFilm::AddSample
                FILM_ADD(permBase + (size_t)s * PERM_STRIDE, scr, code, SanitizeRadiance(L));
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
                    ray.time = Lerp(VanDerCorput(row[2], scr[4]), rp.shutterOpen, rp.shutterClose);   // montecarlo.cpp:243
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
#if PBRT_SPECULAR
                    specularBounce = false;
#endif
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
                        ray.time = r3.w;
                        L = Spec(v.x, v.y, v.z);
#if PBRT_SPECULAR
                        specularBounce = (code & LS_SPEC) != 0;
                        code &= ~LS_SPEC;
#endif
                        code &= ~LS_TAIL;
                        bounces = 3;
                        found = true;
                        state = S_HIT_RESULT;   // picked up (after the RNG top-up) next iteration
                    } else {
                        FILM_ADD(permBase + (size_t)s * PERM_STRIDE, scr, code, Spec(v.x, v.y, v.z));
                        ++s;
                    }
                }
            }
        }

        // ---- the single traversal site
        if (wantTrace) {
            if (STATS) { if (anyHit) ++nShadow; else ++nClosest; }
            found = Traverse<STATS, ANIM>(sc, ray, anyHit, &hit, &st);
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

#include "kvertex.cuh"

// ------------------------------------------------------------------- wavefront
// Phase 1 as a wavefront (the default): the same state machine, but every path lives in a pool slot
// in HBM and each of its blocking requests is served by a kernel specialised for it:
//   k_logic  one thread per pool slot: consume the previous result, run the state machine up to the
//            next ray or (measured-BRDF) evaluation, park / finish / start camera samples
//   k_trace  persistent lanes pull rays from the queue; a lane whose walk ends takes the next ray at
//            once (no lane waits for its warp's longest walk)
//   k_eval   kd-tree queries of IrregIsotropicBRDF::f for the queued requests
// Slots are SoA float4 fields: field f of slot i is pool[f * P + i].
enum {
    F_RAY_O = 0,   // o.xyz, mint
    F_RAY_D,       // d.xyz, maxt (in) / t (out)
    F_HIT,         // prim, b1, b2, maxtIn
    F_T,           // T.rgb, packed {state:8, bounces:8, lightNum:16}
    F_L,           // L.rgb, code
    F_LD,          // Ld.rgb, sPdf
    F_PEND,        // pend.rgb, visMaxt
    F_CTX_P,       // p.xyz, rayEpsilon
    F_CTX_NN,      // nn.xyz, material | (areaLight+1) << 16
    F_CTX_NG,      // ng.xyz, bsU0
    F_CTX_SN,      // sn.xyz, bsU1
    F_CTX_TN,      // tn.xyz, bsUc
    F_WO,          // wo.xyz, task (batch index)
    F_WI,          // wiW.xyz, sample index
    F_EVAL_WO,     // evalWo.xyz, evalFlags
    F_EVAL_WI,     // evalWi.xyz
    F_FEVAL,       // fEval.rgb
    F_ACC,         // phase 2: the task's film accumulator {Lxyz, weightSum} (ImageFilm::Pixel of its current pixel)
    F_TIME,        // ray.time of the camera path (scenes with TransformedPrimitives only)
    F_DL_SUM,      // directlighting: Ld summed over this light's samples so far, packed {light:16, sample:16}
    F_DL_ALL,      // directlighting: UniformSampleAllLights' L so far
    F_COUNT
};
// Request queues, one per resume state ("kind"), so that the lanes of a warp of k_logic are all in the same
// state, k_trace sees shadow rays and extension rays in separate runs, and nothing iterates over retired slots.
// Two sets are ping-ponged: iteration i of k_logic drains set i&1 and fills set (i+1)&1.
enum { K_E_LIGHT = 0, K_E_MIS, K_E_PATH, K_HIT, K_SHADOW, K_MIS,
       K_RESUME,   // phase 2: a parked sample picked up while its task's RNG window needs a top-up (no ray to trace)
       K_COUNT };
enum { C_QN = 0, C_TRACE_CUR = 2 * K_COUNT, C_EVAL_CUR, C_NEXT_SAMPLE, C_COUNT };

struct WfArgs {
    DScene sc;
    RenderParams rp;
    BatchArgs b;
    float4 *pool;
    uint32_t P;
    uint32_t *q;         // [2 sets][K_COUNT][P] slot indices
    unsigned *counters;  // C_QN + set * K_COUNT + kind: queue lengths; cursors; next camera sample
    unsigned totalSamples;
    int iter;            // 0: every slot starts a camera sample
    PbrtGpuPixel *film;  // phase 2
    int directFilm;
    DeviceStats *stats;
    DlFrame *dlFrames;   // [P][DL_MAX_DEPTH] DirectLightingIntegrator over specular materials, else null
};

DEV unsigned AggregatedAdd(unsigned *ctr) {   // warp-aggregated atomicAdd(ctr, 1) for the currently converged lanes
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(ctr, (unsigned)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

// PHASE 1: slots are camera samples (bounces 0..2, table-driven).  PHASE 2: slot i is task-pixel i's chain of
// parked samples, resumed one after the other with the task's sequential RNG stream; finished samples are added
// to the film in sample order in between (what k_path<2> does per thread, here as a wavefront).
template <int PHASE, bool STATS, bool ANIM, int INTEG>
__global__ void __launch_bounds__(128, K_LOGIC_MIN_BLOCKS) k_logic(const __grid_constant__ WfArgs a) {
    __shared__ uint32_t smt[PHASE == 2 ? 4 : 1][PHASE == 2 ? MT_N : 1];
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const DScene &sc = a.sc;
    const RenderParams &rp = a.rp;
    const int spp = rp.spp;
    float4 *pool = a.pool;
    const size_t P = a.P;
    const int setIn = a.iter & 1, setOut = setIn ^ 1;
    // which queue of the input set does this thread serve?
    uint32_t slot = 0;
    int kindIn = -1;
    bool live = true;
    if (a.iter == 0) {
        if (idx >= (PHASE == 1 ? a.P : (uint32_t)a.b.nBatch)) live = false;
        slot = idx;
    } else {
        uint32_t base = 0;
        const unsigned *qn = a.counters + C_QN + setIn * K_COUNT;
        for (int k = 0; k < K_COUNT; ++k) {
            const uint32_t n = qn[k];
            if (idx < base + n) { kindIn = k; break; }
            base += n;
        }
        if (kindIn < 0) live = false;
        else slot = a.q[((size_t)setIn * K_COUNT + kindIn) * P + (idx - base)];
    }
    if (PHASE == 1 && !live) return;
#define FLD(f) pool[(size_t)(f) * P + slot]

    int state = S_NEW_SAMPLE, bounces = 0, lightNum = 0;
    Spec T(1.f);
    int s = 0, code = 0, bi = 0, mti = 0;
    Ray ray;
    ray.o = V3(0, 0, 0); ray.d = V3(0, 0, 1); ray.mint = 0.f; ray.maxt = 0.f; ray.time = 0.f;
    Hit hit;
    hit.prim = -1; hit.t = 0.f; hit.b1 = hit.b2 = 0.f; hit.maxtIn = 0.f;
    bool found = false;
    Spec L(0.f), Ld(0.f), pend(0.f), fLoaded(0.f);
    ShadeCtx ctx;
    ctx.p = ctx.nn = ctx.ng = ctx.sn = ctx.tn = V3(0, 0, 1);
    ctx.material = 0; ctx.areaLight = -1; ctx.rayEpsilon = 0.f;
    V3 wo(0, 0, 1), wiW(0, 0, 1), evalWoIn(0, 0, 1), evalWiIn(0, 0, 1);
    float bsU0 = 0.f, bsU1 = 0.f, bsUc = 0.f, sPdf = 0.f, visMaxt = 0.f;
    unsigned long long nShadedN = 0;
    bool ctxDirty = false;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#if PBRT_SPECULAR
    bool specularBounce = false;           // path.cpp:53; bit 31 of the packed state word between iterations
    // DirectLightingIntegrator's SpecularReflect / SpecularTransmit recursion (integrator.cpp:177-258) as a post-order walk:
    // depth / stage / return flag / drawing-node count travel in F_DL_ALL.w, the frames of the ancestors live in the
    // slot's column of a.dlFrames (global memory: only scenes with specular materials have any)
    int dlDepth = 0, dlStage = 0, dlNodes = 0;
    bool dlReturn = false;
#endif
    int dlI = 0, dlJ = 0;                  // DirectLightingIntegrator's light loop
    Spec dlSum(0.f), dlAll(0.f);
    const int rowStride = INTEG == PBRTGPU_INTEGRATOR_DIRECTLIGHTING ? rp.lay.rowStride : PERM_STRIDE;

    if (PHASE == 2 && live && kindIn < 0) {   // iteration 0: the task-pixel behind this slot
        bi = a.b.order ? (int)a.b.order[idx] : (int)idx;
        const TaskRec tr0 = a.b.tasks[bi];
        if (a.b.round >= (tr0.x1 - tr0.x0) * (tr0.y1 - tr0.y0)) live = false;
        else mti = a.b.mti[bi];
    }
    if (live && kindIn >= 0) {   // load only what this resume state reads
        const float4 fT = FLD(F_T), fL = FLD(F_L), fWo = FLD(F_WO), fWi = FLD(F_WI);
        const uint32_t packed = __float_as_uint(fT.w);
#if PBRT_SPECULAR
        state = packed & 0xffu; bounces = (packed >> 8) & 0xffu; lightNum = (packed >> 16) & 0x7fffu;
        specularBounce = (packed >> 31) != 0;
#else
        state = packed & 0xffu; bounces = (packed >> 8) & 0xffu; lightNum = packed >> 16;
#endif
        T = Spec(fT.x, fT.y, fT.z);
        L = Spec(fL.x, fL.y, fL.z); code = __float_as_int(fL.w);
        wo = V3(fWo.x, fWo.y, fWo.z); bi = __float_as_int(fWo.w);
        wiW = V3(fWi.x, fWi.y, fWi.z); s = __float_as_int(fWi.w);
        if (kindIn >= K_HIT) {   // a traversal result
            const float4 r0 = FLD(F_RAY_O), r1 = FLD(F_RAY_D), h = FLD(F_HIT);
            ray.o = V3(r0.x, r0.y, r0.z); ray.mint = r0.w;
            ray.d = V3(r1.x, r1.y, r1.z); ray.maxt = r1.w;
            hit.prim = __float_as_int(h.x); hit.b1 = h.y; hit.b2 = h.z; hit.maxtIn = h.w; hit.t = r1.w;
            found = hit.prim >= 0;
        }
        if (kindIn != K_HIT && kindIn != K_RESUME) {   // mid-vertex states need the shading context
            const float4 c0 = FLD(F_CTX_P), c1 = FLD(F_CTX_NN), c2 = FLD(F_CTX_NG), c3 = FLD(F_CTX_SN), c4 = FLD(F_CTX_TN);
            ctx.p = V3(c0.x, c0.y, c0.z); ctx.rayEpsilon = c0.w;
            ctx.nn = V3(c1.x, c1.y, c1.z);
            const uint32_t ml = __float_as_uint(c1.w);
            ctx.material = (int)(ml & 0xffffu); ctx.areaLight = (int)(ml >> 16) - 1;
            ctx.ng = V3(c2.x, c2.y, c2.z); bsU0 = c2.w;
            ctx.sn = V3(c3.x, c3.y, c3.z); bsU1 = c3.w;
            ctx.tn = V3(c4.x, c4.y, c4.z); bsUc = c4.w;
            const float4 d0 = FLD(F_LD), d1 = FLD(F_PEND);
            Ld = Spec(d0.x, d0.y, d0.z); sPdf = d0.w;
            pend = Spec(d1.x, d1.y, d1.z); visMaxt = d1.w;
            if (kindIn <= K_E_PATH) {
                const float4 e0 = FLD(F_EVAL_WO), e1 = FLD(F_EVAL_WI), e2 = FLD(F_FEVAL);
                evalWoIn = V3(e0.x, e0.y, e0.z); evalWiIn = V3(e1.x, e1.y, e1.z);
                fLoaded = Spec(e2.x, e2.y, e2.z);
            }
        }
        if (PHASE == 2) {
            const float4 ac = FLD(F_ACC);
            acc[0] = ac.x; acc[1] = ac.y; acc[2] = ac.z; acc[3] = ac.w;
            mti = a.b.mti[bi];
        }
        if (ANIM) ray.time = FLD(F_TIME).x;
        if (INTEG == PBRTGPU_INTEGRATOR_DIRECTLIGHTING) {
            const float4 d0 = FLD(F_DL_SUM), d1 = FLD(F_DL_ALL);
            dlSum = Spec(d0.x, d0.y, d0.z); dlAll = Spec(d1.x, d1.y, d1.z);
            dlI = (int)(__float_as_uint(d0.w) & 0xffffu); dlJ = (int)(__float_as_uint(d0.w) >> 16);
#if PBRT_SPECULAR
            const uint32_t rw = __float_as_uint(d1.w);
            dlDepth = (int)(rw & 0xfu); dlStage = (int)((rw >> 4) & 0x3u); dlReturn = ((rw >> 6) & 1u) != 0; dlNodes = (int)(rw >> 8);
#endif
        }
    }
    if (PHASE == 2) {
        // cooperative RNG top-up before a bounce (<= 11 draws must be addressable), as in k_path<2>
        const unsigned FULL = 0xffffffffu;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const bool need = live && state == S_HIT_RESULT && found && mti >= MT_N;
        unsigned m = __ballot_sync(FULL, need);
        while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            const int tb = __shfl_sync(FULL, bi, l);
            uint32_t *c = a.b.rngCur + (size_t)tb * MT_N, *n = a.b.rngNext + (size_t)tb * MT_N;
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
        if (!live) return;
    }

    // per-task environment (re-derived whenever the slot starts a sample of another task)
    const uint16_t *permBase = nullptr;
    const uint32_t *scr = nullptr;
    float4 *lsBase = nullptr, *recBase = nullptr;
    int xPos = 0, yPos = 0;
    auto bindTask = [&](int task) -> bool {
        const TaskRec tr = a.b.tasks[task];
        const int tw = tr.x1 - tr.x0;
        if (a.b.round >= tw * (tr.y1 - tr.y0)) return false;
        xPos = tr.x0 + a.b.round % tw;
        yPos = tr.y0 + a.b.round / tw;
        permBase = a.b.perm + (size_t)task * spp * rowStride;
        scr = a.b.scr + (size_t)task * SCR_STRIDE;
        lsBase = a.b.ls + (size_t)task * spp;
        recBase = a.b.rec + (size_t)task * spp * 4;
        return true;
    };
    if (state != S_NEW_SAMPLE || PHASE == 2) bindTask(bi);
    const uint32_t *rngCur = a.b.rngCur + (size_t)bi * MT_N, *rngNext = a.b.rngNext + (size_t)bi * MT_N;
    // draw `k` positions past mti (RNG::RandomUInt order), without advancing; phase 1 never draws
    auto rngF = [&](int k) -> float {
        if (PHASE == 1) return 0.f;
        const int j = mti + k;
        return UIntToFloat(Temper(j < MT_N ? __ldcg(rngCur + j) : __ldcg(rngNext + (j - MT_N))));
    };
    const bool direct = a.directFilm != 0;
    bool resumeOut = false;

    ctxDirty = (kindIn < 0 || kindIn == K_HIT || kindIn == K_RESUME);   // the context is (re)built only at a path vertex
    bool wantTraceOut = false, anyHitOut = false, deferOut = false;
    V3 evalWoOut(0, 0, 1), evalWiOut(0, 0, 1);
    int evalFlagsOut = 0;
    for (;;) {
#define PATH_FEVAL_IN fLoaded
#define PATH_EVALWO_IN evalWoIn
#define PATH_EVALWI_IN evalWiIn
#define PATH_DEFER_MEASURED true
#define PATH_ANIM ANIM
#define PATH_INTEG INTEG
#define PATH_COUNT_TAIL() (atomicAdd(a.b.tailCount + bi, 1u), a.b.tailMask ? (void)atomicOr(a.b.tailMask + (size_t)bi * a.b.maskWords + (s >> 5), 1u << (s & 31)) : (void)0)
#define PATH_DL_FRAME(i) a.dlFrames[(size_t)slot * DL_MAX_DEPTH + (size_t)(i)]
#include "path_body.inc"
#undef PATH_DL_FRAME
#undef PATH_COUNT_TAIL
#undef PATH_FEVAL_IN
#undef PATH_EVALWO_IN
#undef PATH_EVALWI_IN
#undef PATH_DEFER_MEASURED
#undef PATH_ANIM
#undef PATH_INTEG
        if (deferEval) {
            deferOut = true;
            evalWoOut = evalWo; evalWiOut = evalWi; evalFlagsOut = evalFlags;
            break;
        }
        fLoaded = Spec(0.f);
        if (PHASE == 2) {
            if (state == S_END_SAMPLE) {
                // samplerrenderer.cpp:111-146: sanity checks, then ImageFilm::AddSample
                FILM_ADD(permBase + (size_t)s * rowStride, scr, code, SanitizeRadiance(L));
                ++s;
                state = S_NEW_SAMPLE;
            }
            if (state == S_NEW_SAMPLE) {
                // walk ls[] in sample order: add finished samples, resume the next parked one
                while (state == S_NEW_SAMPLE) {
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
                        ray.time = r3.w;
                        L = Spec(v.x, v.y, v.z);
#if PBRT_SPECULAR
                        specularBounce = (code & LS_SPEC) != 0;
                        code &= ~LS_SPEC;
#endif
                        code &= ~LS_TAIL;
                        bounces = 3;
                        found = true;
                        state = S_HIT_RESULT;
                        ctxDirty = true;
                    } else {
                        FILM_ADD(permBase + (size_t)s * rowStride, scr, code, Spec(v.x, v.y, v.z));
                        ++s;
                    }
                }
                if (state == S_DONE) break;
                if (mti >= MT_N) { resumeOut = true; break; }   // RNG window must be topped up first (next iteration)
                continue;                                       // shade the parked vertex right away
            }
        } else {
        if (state == S_END_SAMPLE) {
            const Spec Ls = SanitizeRadiance(L);
#if PBRT_SPECULAR
            // DirectLightingIntegrator: the number of Li nodes that made the two specular calls (6 RNG draws each) rides in
            // the code word's upper bits to k_film_direct
            if (INTEG == PBRTGPU_INTEGRATOR_DIRECTLIGHTING) code |= dlNodes << 8;
            dlDepth = dlStage = dlNodes = 0; dlReturn = false;
#endif
            lsBase[s] = make_float4(Ls.r, Ls.g, Ls.b, __int_as_float(code));
            state = S_NEW_SAMPLE;
        }
        if (state == S_NEW_SAMPLE) {
            // next camera sample of the batch: (task, sample) = (g / spp, g % spp)
            bool have = false;
            for (;;) {
                const unsigned g = AggregatedAdd(a.counters + C_NEXT_SAMPLE);
                if (g >= a.totalSamples) break;
                bi = (int)(g / (unsigned)spp);
                s = (int)(g - (unsigned)bi * (unsigned)spp);
                if (bindTask(bi)) { have = true; break; }
            }
            if (!have) { state = S_DONE; break; }
            // LDPixelSample's camera sample + PerspectiveCamera::GenerateRayDifferential (perspective.cpp:81-138)
            const uint16_t *row = permBase + (size_t)s * rowStride;
            const uint32_t pi = row[0];
            const float imageX = xPos + VanDerCorput(pi, scr[0]);
            const float imageY = yPos + Sobol2(pi, scr[1]);
            code = FilmFootprint(rp, xPos, yPos, imageX, imageY);
            V3 Pcamera = XfPoint(rp.rasterToCamera, V3(imageX, imageY, 0));
            V3 dir = Normalize(V3(Pcamera.x, Pcamera.y, Pcamera.z));
            ray.o = V3(0, 0, 0); ray.d = dir; ray.mint = 0.f; ray.maxt = INF_F;
                    ray.time = Lerp(VanDerCorput(row[2], scr[4]), rp.shutterOpen, rp.shutterClose);   // montecarlo.cpp:243
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
#if PBRT_SPECULAR
            specularBounce = false;
#endif
            bounces = 0;
            wantTrace = true; anyHit = false;
            state = S_HIT_RESULT;
        }
        }
        if (wantTrace) { wantTraceOut = true; anyHitOut = anyHit; break; }
        // otherwise the state is non-blocking (S_AFTER_LIGHT / S_AFTER_DIRECT): keep going
    }

    // ---- write the slot back and queue its request under its resume state
    if (state == S_DONE) return;   // no more camera samples: the slot retires (it is in no queue any more)
    const int kindOut = resumeOut ? K_RESUME : state == S_E_LIGHT ? K_E_LIGHT : state == S_E_MIS ? K_E_MIS
                        : state == S_E_PATH ? K_E_PATH : state == S_HIT_RESULT ? K_HIT
                        : state == S_SHADOW_RESULT ? K_SHADOW : K_MIS;
    if (PHASE == 2) {
        FLD(F_ACC) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        a.b.mti[bi] = mti;
    }
    if (ANIM) FLD(F_TIME) = make_float4(ray.time, 0.f, 0.f, 0.f);
    if (INTEG == PBRTGPU_INTEGRATOR_DIRECTLIGHTING) {
        FLD(F_DL_SUM) = make_float4(dlSum.r, dlSum.g, dlSum.b, __uint_as_float((uint32_t)dlI | ((uint32_t)dlJ << 16)));
#if PBRT_SPECULAR
        FLD(F_DL_ALL) = make_float4(dlAll.r, dlAll.g, dlAll.b, __uint_as_float((uint32_t)dlDepth | ((uint32_t)dlStage << 4) |
                                                                                   (dlReturn ? 0x40u : 0u) | ((uint32_t)dlNodes << 8)));
#else
        FLD(F_DL_ALL) = make_float4(dlAll.r, dlAll.g, dlAll.b, 0.f);
#endif
    }
#if PBRT_SPECULAR
    FLD(F_T) = make_float4(T.r, T.g, T.b, __uint_as_float((uint32_t)state | ((uint32_t)bounces << 8) | ((uint32_t)lightNum << 16) |
                                                          (specularBounce ? 0x80000000u : 0u)));
#else
    FLD(F_T) = make_float4(T.r, T.g, T.b, __uint_as_float((uint32_t)state | ((uint32_t)bounces << 8) | ((uint32_t)lightNum << 16)));
#endif
    FLD(F_L) = make_float4(L.r, L.g, L.b, __int_as_float(code));
    FLD(F_WO) = make_float4(wo.x, wo.y, wo.z, __int_as_float(bi));
    FLD(F_WI) = make_float4(wiW.x, wiW.y, wiW.z, __int_as_float(s));
    if (kindOut != K_HIT) {
        if (ctxDirty) {
            FLD(F_CTX_P) = make_float4(ctx.p.x, ctx.p.y, ctx.p.z, ctx.rayEpsilon);
            FLD(F_CTX_NN) = make_float4(ctx.nn.x, ctx.nn.y, ctx.nn.z,
                                       __uint_as_float((uint32_t)ctx.material | ((uint32_t)(ctx.areaLight + 1) << 16)));
            FLD(F_CTX_NG) = make_float4(ctx.ng.x, ctx.ng.y, ctx.ng.z, bsU0);
            FLD(F_CTX_SN) = make_float4(ctx.sn.x, ctx.sn.y, ctx.sn.z, bsU1);
            FLD(F_CTX_TN) = make_float4(ctx.tn.x, ctx.tn.y, ctx.tn.z, bsUc);
        }
        FLD(F_LD) = make_float4(Ld.r, Ld.g, Ld.b, sPdf);
        FLD(F_PEND) = make_float4(pend.r, pend.g, pend.b, visMaxt);
    }
    if (deferOut) {
        FLD(F_EVAL_WO) = make_float4(evalWoOut.x, evalWoOut.y, evalWoOut.z, __int_as_float(evalFlagsOut));
        FLD(F_EVAL_WI) = make_float4(evalWiOut.x, evalWiOut.y, evalWiOut.z, 0.f);
    } else {
        FLD(F_RAY_O) = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint);
        FLD(F_RAY_D) = make_float4(ray.d.x, ray.d.y, ray.d.z, ray.maxt);
        if (resumeOut) FLD(F_HIT) = make_float4(__int_as_float(hit.prim), hit.b1, hit.b2, hit.maxtIn);
    }
    {   // warp-aggregated push, one atomic per (warp, kind)
        const unsigned peers = __match_any_sync(__activemask(), kindOut);
        const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(a.counters + C_QN + setOut * K_COUNT + kindOut, (unsigned)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        a.q[((size_t)setOut * K_COUNT + kindOut) * P + base + __popc(peers & ((1u << lane) - 1u))] = slot;
    }
    (void)wantTraceOut; (void)anyHitOut;
    if (STATS && nShadedN) atomicAdd(&a.stats->shadedN, nShadedN);
#undef FLD
}

// DirectLightingIntegrator's "phase 2": the samples of a pixel are added to the film in sample order
// (samplerrenderer.cpp:147-157) and the task's RNG moves past the draws SpecularReflect / SpecularTransmit made for
// every camera ray that hit (BSDFSample(rng) twice = 6 draws, integrator.cpp:184,227; only when depth + 1 < maxDepth,
// directlighting.cpp:108).  One thread per task-pixel; the stream position is advanced block by block,
// warp-cooperatively, only when another pixel of the task follows (`advance`).
__global__ void __launch_bounds__(128) k_film_direct(const __grid_constant__ PathArgs a, int advance) {
    __shared__ uint32_t smt[4][MT_N];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bi = blockIdx.x * blockDim.x + threadIdx.x;
    const RenderParams &rp = a.rp;
    const int spp = rp.spp;
    bool live = bi < a.b.nBatch;
    int mti = 0;
    if (live) {
        const TaskRec tr = a.b.tasks[bi];
        const int tw = tr.x1 - tr.x0;
        if (a.b.round >= tw * (tr.y1 - tr.y0)) live = false;
        else {
            const int xPos = tr.x0 + a.b.round % tw, yPos = tr.y0 + a.b.round / tw;
            const float4 *ls = a.b.ls + (size_t)bi * spp;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            int hits = 0;
            const bool direct = a.directFilm != 0;
            for (int s = 0; s < spp; ++s) {
                const float4 v = __ldcs(ls + s);
                const int code = __float_as_int(v.w);
#if PBRT_SPECULAR
                hits += code >> 8;   // Li nodes that drew (the camera vertex and, under mirror / glass, its specular descendants)
#else
                hits += ((code & LS_HIT) && 0 + 1 < rp.maxDepth) ? 1 : 0;
#endif
                FILM_ADD(a.b.perm + ((size_t)bi * spp + s) * rp.lay.rowStride, a.b.scr + (size_t)bi * SCR_STRIDE,
                         code & (LS_LEFT | LS_RIGHT | LS_UP | LS_DOWN), Spec(v.x, v.y, v.z));
            }
            if (!direct) {
                PbrtGpuPixel *px = a.film + (size_t)(yPos - rp.yPixelStart) * rp.xPixelCount + (xPos - rp.xPixelStart);
                atomicAdd(&px->Lxyz[0], acc[0]);
                atomicAdd(&px->Lxyz[1], acc[1]);
                atomicAdd(&px->Lxyz[2], acc[2]);
                atomicAdd(&px->weightSum, acc[3]);
            }
            mti = a.b.mti[bi] + 6 * hits;
        }
    }
    if (!advance) return;
    // bring the two-block window (cur, next) forward until mti lies in [0, 2 * MT_N) again (what k_tables expects)
    for (;;) {
        unsigned m = __ballot_sync(FULL, live && mti >= MT_N);
        if (!m) break;
        while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            const int tb = __shfl_sync(FULL, bi, l);
            uint32_t *c = a.b.rngCur + (size_t)tb * MT_N, *n = a.b.rngNext + (size_t)tb * MT_N;
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
    if (live) a.b.mti[bi] = mti;
}

// persistent ray queue consumer (tuning constants: kcommon.cuh)
template <bool STATS, bool ANIM>
__global__ void __launch_bounds__(128, K_TRACE_MIN_BLOCKS) k_trace(const __grid_constant__ WfArgs a) {
    const DScene &sc = a.sc;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // this iteration's rays: the K_HIT, K_SHADOW and K_MIS queues of the set k_logic just filled, back to back
    const int setOut = (a.iter & 1) ^ 1;
    const unsigned *qn = a.counters + C_QN + setOut * K_COUNT;
    const unsigned n0 = qn[K_HIT], n1 = n0 + qn[K_SHADOW], n = n1 + qn[K_MIS];
    float4 *pool = a.pool;
    const size_t P = a.P;
    const uint32_t *qHit = a.q + ((size_t)setOut * K_COUNT + K_HIT) * P;
    const uint32_t *qShadow = a.q + ((size_t)setOut * K_COUNT + K_SHADOW) * P;
    const uint32_t *qMis = a.q + ((size_t)setOut * K_COUNT + K_MIS) * P;
    Trav t;
#if PBRT_SMEM_STACK
    __shared__ uint32_t sStack[PBRT_SMEM_STACK * 128];
    t.sstk = sStack + threadIdx.x;
#endif
#if PBRT_TOP_LEVELS
    __shared__ __align__(128) float4 sTop[2 * PBRT_TOP_SLOTS];
    __shared__ __align__(8) uint64_t sTopBar;
    StageTopNodes(sTop, &sTopBar, sc.topNodes);
    const float4 *smemTop = sTop;
#else
    const float4 *smemTop = nullptr;
#endif
    bool active = false;
    bool leaf = false;        // the lane stands before a leaf (TravDescend returned 1)
    uint32_t slot = 0;
    TraceStats st;
    st.nodes = st.prims = 0;
    unsigned long long nClosest = 0, nShadow = 0;
    // Rays are handed out in chunks: a warp reserves TRACE_CHUNK queue entries with ONE atomic on the shared cursor
    // and deals them to its lanes from registers (a per-refill atomic on one address was 40 % of this kernel's stall
    // samples, profiles/r01_k_trace_*).  cPos/cEnd are warp-uniform.
    unsigned cPos = 0, cEnd = 0;
    bool exhausted = false;   // the cursor has passed the end of the queue
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !active);
        if (idle && !(exhausted && cPos >= cEnd)) {
            if (cPos >= cEnd) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(a.counters + C_TRACE_CUR, (unsigned)TRACE_CHUNK);
                base = __shfl_sync(FULL, base, 0);
                if (base >= n) { exhausted = true; cPos = cEnd = 0; }
                else { cPos = base; cEnd = min(base + (unsigned)TRACE_CHUNK, n); }
            }
            if (!active) {
                const unsigned idx = cPos + __popc(idle & ((1u << lane) - 1u));
                if (idx < cEnd) {
                    const bool any = idx >= n0 && idx < n1;
                    slot = idx < n0 ? qHit[idx] : (idx < n1 ? qShadow[idx - n0] : qMis[idx - n1]);
                    const float4 r0 = pool[(size_t)F_RAY_O * P + slot], r1 = pool[(size_t)F_RAY_D * P + slot];
                    Ray ray;
                    ray.o = V3(r0.x, r0.y, r0.z); ray.mint = r0.w;
                    ray.d = V3(r1.x, r1.y, r1.z); ray.maxt = r1.w;
                    ray.time = ANIM ? pool[(size_t)F_TIME * P + slot].x : 0.f;
                    TravInit(t, ray, any);
                    if (STATS) { if (any) ++nShadow; else ++nClosest; }
                    active = true; leaf = false;
                }
            }
            cPos = min(cEnd, cPos + (unsigned)__popc(idle));
        }
        const bool more = !(exhausted && cPos >= cEnd);   // could idle lanes still be given rays?
        if (!__any_sync(FULL, active)) {
            if (!more) break;
            continue;
        }
        // walk until a good share of the warp has finished, then refill (nothing to refill with: walk on)
#if PBRT_TRACE_LOOP == 0 || PBRT_PAIR_NODES
        for (int it = 0; it < 64; ++it) {
            const bool fin = active && TravStep<STATS, ANIM>(sc, t, &st, smemTop);
#elif PBRT_TRACE_LOOP == 2
        for (int it = 0; it < 64; ++it) {
            // one box test per walking lane; a lane that reaches a leaf waits until PBRT_LEAF_BATCH lanes stand before one (or
            // nobody walks any more), then those lanes test their primitives together (the double-precision cross products of
            // a lone leaf test no longer run with 5 of 32 lanes, profiles/r02b_k_wtrace_ncu_full.txt)
            bool fin = false;
            if (active && !leaf) {
                const int r = TravDescend<STATS, ANIM>(sc, t, &st, 1, smemTop);
                leaf = r == 1;
                fin = r == 2;
            }
            const unsigned atLeaf = __ballot_sync(FULL, active && leaf);
            const unsigned walking = __ballot_sync(FULL, active && !leaf && !fin);
            if (atLeaf && (__popc(atLeaf) >= PBRT_LEAF_BATCH || !walking)) {
                if (active && leaf) {
                    leaf = false;
                    fin = TravLeaf<STATS, ANIM>(sc, t, &st);
                }
            }
#else
        for (int it = 0; it < 8; ++it) {
            bool fin = false;
            for (int k = 0; k < TRACE_DESCEND_STEPS; ++k) {   // the warp's loop: every round ends in a vote
                const bool walking = active && !leaf && !fin;
                if (!__any_sync(FULL, walking)) break;
                if (walking) {
                    const int r = TravDescend<STATS, ANIM>(sc, t, &st, 1, smemTop);
                    leaf = r == 1;
                    fin = r == 2;
                }
            }
            if (active && leaf) {
                leaf = false;
                fin = TravLeaf<STATS, ANIM>(sc, t, &st);
            }
#endif
            if (fin) {
                pool[(size_t)F_HIT * P + slot] = make_float4(__int_as_float(t.found ? t.hit.prim : -1), t.hit.b1, t.hit.b2, t.hit.maxtIn);
                pool[(size_t)F_RAY_D * P + slot].w = t.ray.maxt;
                active = false;
            }
            const int nActive = __popc(__ballot_sync(FULL, active));
            if (nActive == 0 || (more && nActive < TRACE_REFILL_BELOW)) break;
        }
    }
    if (STATS) {
        unsigned long long v[4] = {nClosest, nShadow, st.nodes, st.prims};
        for (int k = 0; k < 4; ++k) {
            unsigned long long x = v[k];
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(FULL, x, o);
            v[k] = x;
        }
        if (lane == 0) {
            atomicAdd(&a.stats->closestRays, v[0]);
            atomicAdd(&a.stats->shadowRays, v[1]);
            atomicAdd(&a.stats->nodes, v[2]);
            atomicAdd(&a.stats->prims, v[3]);
        }
    }
}

#include "kwave.cuh"

// measured-BRDF evaluations (IrregIsotropicBRDF::f) of the queued requests
__global__ void __launch_bounds__(128, K_EVAL_MIN_BLOCKS) k_eval(const __grid_constant__ WfArgs a) {
    const int setOut = (a.iter & 1) ^ 1;
    const unsigned *qn = a.counters + C_QN + setOut * K_COUNT;
    const unsigned n0 = qn[K_E_LIGHT], n1 = n0 + qn[K_E_MIS], n = n1 + qn[K_E_PATH];
    const size_t P = a.P;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int kind = idx < n0 ? K_E_LIGHT : (idx < n1 ? K_E_MIS : K_E_PATH);
        const uint32_t slot = a.q[((size_t)setOut * K_COUNT + kind) * P + (idx - (idx < n0 ? 0 : (idx < n1 ? n0 : n1)))];
        const float4 c1 = a.pool[(size_t)F_CTX_NN * P + slot];
        const float4 e0 = a.pool[(size_t)F_EVAL_WO * P + slot], e1 = a.pool[(size_t)F_EVAL_WI * P + slot];
        const int material = (int)(__float_as_uint(c1.w) & 0xffffu);
        const int flags = __float_as_int(e0.w);
        BxSet bx;
        MaterialBxDFs(a.sc, material, &bx);
        Spec f(0.f);
        for (int i = 0; i < bx.n; ++i)
            if (Matches(bx.type(i), flags)) f = f + BxF(a.sc, bx, i, V3(e0.x, e0.y, e0.z), V3(e1.x, e1.y, e1.z));
        a.pool[(size_t)F_FEVAL * P + slot] = make_float4(f.r, f.g, f.b, 0.f);
    }
}

// ------------------------------------------------------------- ray-batch kernels
template <bool ANY>
__global__ void k_intersect(const __grid_constant__ DScene sc, const PbrtGpuRay *rays, int n, PbrtGpuHit *hits, uint8_t *occ) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray ray;
    ray.o = V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]);
    ray.d = V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
    ray.mint = rays[i].mint; ray.maxt = rays[i].maxt; ray.time = rays[i].time;
    Hit hit;
    TraceStats st;
    const bool found = Traverse<false, true>(sc, ray, ANY, &hit, &st);
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
        DXf w2p;
        HitGeometry<true>(sc, r0, hit, &dg, &eps, &mat, &al, &tri, &ptype, &w2p);
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

// One lane = one CUDA stream + one set of scratch buffers (+ one host thread while a render call runs).
struct Lane {
    cudaStream_t stream = nullptr;
    DevBuf<TaskRec> tasks;
    DevBuf<uint16_t> perm;
    DevBuf<uint32_t> scr, rngCur, rngNext;
    DevBuf<int> mti;
    DevBuf<float4> ls, rec;
    DevBuf<unsigned> tailCount, tailKeysOut, orderIn, orderOut, tailFlag, workCounter;
    DevBuf<uint32_t> tailMask;        // [task][max(1, spp / 32)] parked-sample bits of the task's current pixel
    DevBuf<unsigned char> sortTemp;
    DevBuf<uint8_t> tgt;              // general table kernel: shuffle targets
    DevBuf<float4> pool;              // wavefront path slots (F_COUNT fields x P)
    DevBuf<uint32_t> traceQ;
    DevBuf<uint32_t> evalQ;           // vertex wavefront: (slot, direction) evaluation requests
    DevBuf<uint32_t> liveList;        // chain wavefront: [2][maxBatch] live chains, compacted
    DevBuf<uint32_t> sortList;        // vertex wavefront: [2][P] slots sorted by their next action
    DevBuf<uint8_t> sortClass;        // ... and the class bytes the lists are built from
    DevBuf<DlFrame> dlFrames;         // DirectLightingIntegrator over specular materials: [P][DL_MAX_DEPTH]
    DevBuf<unsigned> counters;
    unsigned *hostCounters = nullptr;  // pinned
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    ~Lane() {
        for (int i = 0; i < 4; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
        if (hostCounters) cudaFreeHost(hostCounters);
        if (stream) cudaStreamDestroy(stream);
    }
};

// Lanes outlive scene handles: released lanes wait here, per device, for the next render call.
struct LaneCache {
    std::mutex m;
    std::vector<Lane *> free[64];
};
static LaneCache &Lanes() {
    static LaneCache *c = new LaneCache;   // never destroyed: no cudaFree during process teardown
    return *c;
}
static Lane *AcquireLane(int device) {
    LaneCache &c = Lanes();
    std::lock_guard<std::mutex> g(c.m);
    std::vector<Lane *> &f = c.free[device & 63];
    if (f.empty()) return new Lane;
    Lane *l = f.back();
    f.pop_back();
    return l;
}
static void ReleaseLane(int device, Lane *l) {
    LaneCache &c = Lanes();
    std::lock_guard<std::mutex> g(c.m);
    c.free[device & 63].push_back(l);
}

struct PbrtGpuScene {
    int device = 0;
    DScene d;
    DevBuf<float4> nodes, prims;
#if PBRT_PAIR_NODES
    DevBuf<float4> pairs;
    DevBuf<uint32_t> pairOfNode;
#endif
    DevBuf<float4> topNodes;
    DevBuf<uint4> tris;
    DevBuf<PbrtGpuMesh> meshes;
    DevBuf<float> N, uv;
    DevBuf<PbrtGpuQuadric> quadrics;
    DevBuf<PbrtGpuXform> xforms;
    DevBuf<PbrtGpuMaterial> materials;
    DevBuf<PbrtGpuMeasured> measured;
    DevBuf<PbrtGpuKdNode> kdNodes;
    DevBuf<PbrtGpuBrdfSample> kdData;
#if PBRT_KD_INTERLEAVED
    DevBuf<float4> kdRec;
#endif
    DevBuf<PbrtGpuLight> lights;
#if PBRT_INFINITE
    DevBuf<PbrtGpuEnvMap> envMaps;
    DevBuf<float> envData;
#endif
    DevBuf<PbrtGpuLightShape> lightShapes;
    DevBuf<PbrtGpuAnimated> animated;
    // per-render scratch: one Lane per concurrent batch pipeline (see pbrtgpu_render_device), borrowed from a
    // process-wide cache so that a new scene handle does not pay for tens of GB of cudaMalloc again
    static const int MAX_LANES = 4;
    Lane *lanes[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    DevBuf<PbrtGpuPixel> film;
    DevBuf<DeviceStats> stats;
    DevBuf<PbrtGpuRay> rays;
    DevBuf<PbrtGpuHit> hits;
    DevBuf<uint8_t> occ;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};   // render start / end, the lanes' start gate
    int collectStats = 0;   // 1: ray/node/primitive counters, 2: per-launch timing of the traversal kernel
    int nMeasuredMaterials = 0;
    int nSpecularMaterials = 0;
    int nInfiniteLights = 0;
#if PBRT_FILTER
    DevBuf<float> filterTable;   // the current render's ImageFilm::filterTable (RenderParams::filterTable)
#endif
    bool valid = false;     // false while an upload is in progress or after one failed half way
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
    if (rd.xResolution > 65536 || rd.yResolution > 65536 || (long long)rd.xResolution * rd.yResolution > (1ll << 28))
        return Fail(PBRTGPU_EUNSUPPORTED, "film resolution beyond 2^28 pixels");
    // ImageFilm's crop window is four numbers in [0, 1], low <= high (CreateImageFilm clamps and orders what the scene file
    // says, film/image.cpp:246-257); the film extent is computed from them in int
    for (int i = 0; i < 4; ++i)
        if (!(rd.cropWindow[i] >= 0.f && rd.cropWindow[i] <= 1.f)) return Fail(PBRTGPU_EINVAL, "bad crop window");
    if (rd.cropWindow[0] > rd.cropWindow[1] || rd.cropWindow[2] > rd.cropWindow[3]) return Fail(PBRTGPU_EINVAL, "bad crop window");
    {   // camera: every ray's origin, direction and time come out of these
        bool finite = std::isfinite(rd.lensRadius) && std::isfinite(rd.focalDistance) && std::isfinite(rd.shutterOpen) && std::isfinite(rd.shutterClose);
        for (int i = 0; i < 16; ++i) finite = finite && std::isfinite(rd.rasterToCamera[i]) && std::isfinite(rd.cameraToWorld[i]);
        if (!finite) return Fail(PBRTGPU_EINVAL, "non-finite camera parameter");
    }
    const int spp = rd.samplesPerPixel;
    if (spp <= 0 || (spp & (spp - 1))) return Fail(PBRTGPU_EINVAL, "samplesPerPixel must be a power of two (LDSampler)");
    if (spp > 8192) return Fail(PBRTGPU_EUNSUPPORTED, "samplesPerPixel > 8192 does not fit the table kernel's shared memory");
#if PBRT_FILTER
    if (!(rd.filterXWidth > 0.f) || !(rd.filterYWidth > 0.f)) return Fail(PBRTGPU_EINVAL, "bad filter width");
    if (rd.filterXWidth > 31.f || rd.filterYWidth > 31.f)
        return Fail(PBRTGPU_EUNSUPPORTED, "filters wider than 31 pixels exceed the film code's 64-pixel footprint");
#else
    if (rd.filterXWidth > 0.5f || rd.filterYWidth > 0.5f)
        return Fail(PBRTGPU_EUNSUPPORTED, "filters wider than the 0.5 box need cross-task splats (out of scope)");
    // (narrower boxes leave some samples without any pixel, image.cpp:85-91: the footprint bits of FilmAdd cannot say that)
    if (rd.filterXWidth != 0.5f || rd.filterYWidth != 0.5f)
        return Fail(PBRTGPU_EUNSUPPORTED, "filters narrower than the 0.5 box are rendered by PBRT_FILTER builds only");
    if (rd.hasFilterTable)
        for (int i = 0; i < 256; ++i)
            if (rd.filterTable[i] != 1.f) return Fail(PBRTGPU_EUNSUPPORTED, "pixel filters other than the box filter are rendered by the oracle only");
#endif
    if (rd.nTasks <= 0 || rd.taskBegin < 0 || rd.taskEnd > rd.nTasks || rd.taskBegin > rd.taskEnd)
        return Fail(PBRTGPU_EINVAL, "bad task range");
    if (rd.maxDepth < 0) return Fail(PBRTGPU_EINVAL, "bad maxDepth");
    if (rd.integrator != PBRTGPU_INTEGRATOR_PATH && rd.integrator != PBRTGPU_INTEGRATOR_DIRECTLIGHTING)
        return Fail(PBRTGPU_EUNSUPPORTED, "unknown surface integrator");
    if (rd.n1D < 0 || rd.n1D > PBRTGPU_MAX_SAMPLE_ARRAYS || rd.n2D < 0 || rd.n2D > PBRTGPU_MAX_SAMPLE_ARRAYS)
        return Fail(PBRTGPU_EINVAL, "bad sample-array counts");
    for (int i = 0; i < rd.n1D + rd.n2D; ++i) {
        const int n = i < rd.n1D ? rd.size1D[i] : rd.size2D[i - rd.n1D];
        if (n < 1 || n > MAX_ARRAY_SAMPLES) return Fail(PBRTGPU_EUNSUPPORTED, "a sample array asks for 0 or more than 64 values per pixel sample");
        if ((long long)n * spp > 65536) return Fail(PBRTGPU_EUNSUPPORTED, "samplesPerPixel x array size exceeds the 16-bit table index");
    }
    if (rd.integrator == PBRTGPU_INTEGRATOR_PATH && !IsPathLayout(rd))
        return Fail(PBRTGPU_EINVAL, "PathIntegrator's Sample layout is 14 one-value 1-D and 9 one-value 2-D arrays (path.cpp:41-49)");
    {   // Sampler::ComputeSubWindow multiplies in int (core/sampler.cpp:58: 2*dx*ny < dy*nx); a task grid that overflows
        // it is undefined behaviour in the reference, so it is not a configuration parity can be defined for
        const FilmExtent e = ComputeExtent(rd);
        const long long dx = e.sx1 - e.sx0, dy = e.sy1 - e.sy0;
        long long nx = rd.nTasks, ny = 1;
        bool overflow = false;
        for (;;) {
            if (2 * dx * ny > 0x7fffffffLL || dy * nx > 0x7fffffffLL) { overflow = true; break; }
            if (!((nx & 1) == 0 && 2 * dx * ny < dy * nx)) break;
            nx >>= 1; ny <<= 1;
        }
        if (overflow)
            return Fail(PBRTGPU_EINVAL, "nTasks x film extent overflows the int arithmetic of Sampler::ComputeSubWindow");
    }
    if (rd.shardCount > 1 && (rd.shardIndex < 0 || rd.shardIndex >= rd.shardCount))
        return Fail(PBRTGPU_EINVAL, "bad shard index");
    return PBRTGPU_OK;
}

template <int WPB> cudaError_t LaunchTables(const BatchArgs &b, int spp, const TableLaunch &tl, int stages, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(k_tables<WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl.smemBytes);
    if (e != cudaSuccess) return e;
    const int blocks = (b.nBatch + WPB - 1) / WPB;
    k_tables<WPB><<<blocks, 32 * WPB, tl.smemBytes, stream>>>(b, spp, tl.G, stages);
    return cudaGetLastError();
}
// PathIntegrator's tables in two launches: the draw-stream walk with eight warps per block (one MT19937 block of shared
// memory each), then the shuffles + row assembly with as many warps as the index arrays leave room for
cudaError_t LaunchTablesAny(const BatchArgs &b, int spp, const TableLaunch &, cudaStream_t stream) {
    TableLaunch s1;
    s1.G = 0; s1.wpb = 8; s1.smemBytes = (size_t)8 * MT_N * 4;
    cudaError_t e = LaunchTables<8>(b, spp, s1, 1, stream);
    if (e != cudaSuccess) return e;
    const TableLaunch tl = PlanTables(spp, 2);
    switch (tl.wpb) {
    case 4: return LaunchTables<4>(b, spp, tl, 2, stream);
    case 3: return LaunchTables<3>(b, spp, tl, 2, stream);
    case 2: return LaunchTables<2>(b, spp, tl, 2, stream);
    default: return LaunchTables<1>(b, spp, tl, 2, stream);
    }
}

template <int WPB> cudaError_t LaunchTablesGeneral(const BatchArgs &b, int spp, const TableLaunch &tl, const TableArrays &T,
                                                   uint8_t *tgt, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(k_tables_general<WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl.smemBytes);
    if (e != cudaSuccess) return e;
    const int blocks = (b.nBatch + WPB - 1) / WPB;
    k_tables_general<WPB><<<blocks, 32 * WPB, tl.smemBytes, stream>>>(b, spp, tl.G, T, tgt);
    return cudaGetLastError();
}
cudaError_t LaunchTablesGeneralAny(const BatchArgs &b, int spp, const TableLaunch &tl, const TableArrays &T, uint8_t *tgt,
                                   cudaStream_t stream) {
    switch (tl.wpb) {
    case 4: return LaunchTablesGeneral<4>(b, spp, tl, T, tgt, stream);
    case 3: return LaunchTablesGeneral<3>(b, spp, tl, T, tgt, stream);
    case 2: return LaunchTablesGeneral<2>(b, spp, tl, T, tgt, stream);
    default: return LaunchTablesGeneral<1>(b, spp, tl, T, tgt, stream);
    }
}

// Development switches, read once per render CALL (never inside its batch / round / iteration loops).
struct RenderConfig {
    bool wavefront = true;     // PBRTGPU_WAVEFRONT=0: phase 1 as the one-warp-per-pixel megakernel k_path<1> (A/B baseline)
    bool wavefront2 = false;   // PBRTGPU_WAVEFRONT2=1: phase 2 as a wavefront (measured slower)
    // PBRTGPU_VERTEX bit 0: phase 1 by k_vertex<1>, bit 1: phase 2 by k_vertex<2> + k_film.  Measured on C2 (bunny 1024x1024,
    // 1024 spp, one lane; profiles/r02_*): phase 1 wavefront 3.2 s vs k_vertex<1> 4.8 s (8.2 of 32 lanes active: a camera
    // sample's 0..3 evaluations and 0..3 rays per vertex are too uneven for lock-step lanes; the wavefront compacts them),
    // phase 2 k_path<2> 814 ms vs k_vertex<2> + k_film 567 ms.  Default: wavefront phase 1, vertex-synchronous phase 2.
    int vertex = 2;
    // PBRTGPU_WAVE=0: PathIntegrator's phase 1 on the request-granular wavefront (k_logic / k_eval / k_trace) instead of the
    // vertex wavefront (kwave.cuh: k_wshade / k_weval / k_wtrace)
    // bit 1: phase 2 (the task chains) on it as well instead of k_vertex<2>.  Measured on C2, one GPU (profiles/r02d_*, r02e_*):
    // k_vertex<2> 541 ms per step, the chain wavefront 794 ms -- every iteration costs the LATENCY of one vertex (one shade,
    // one kd-tree evaluation, one ray: ~150 us even for a single chain, r02e_launches.csv) and a batch runs as many
    // iterations as its longest chain has vertices; the megakernel's lanes pay only their own chain's latency.  Default 1.
    int wave = 1;
    bool wsort = true;         // PBRTGPU_WSORT=0: k_wshade<1> serves slot i with thread i instead of the slots sorted by next action
    int tailSort = 2;          // PBRTGPU_TAIL_SORT 0: chains in task order, 2: longest chain first
    int lockstep = 1;
    size_t pool = (size_t)1 << 22;
    int lanes = 0;             // 0: automatic
    bool traceIters = false;
    int vertexBlocksPerSM = K_VERTEX_MIN_BLOCKS;
};
inline int WvLiveCounterHost(int list) { return list ? WC_LIVE2 : WC_LIVE; }
RenderConfig ReadConfig() {
    {
        RenderConfig c;
        if (const char *e = getenv("PBRTGPU_WAVEFRONT")) c.wavefront = atoi(e) != 0;
        if (const char *e = getenv("PBRTGPU_WAVEFRONT2")) c.wavefront2 = atoi(e) != 0;
        if (const char *e = getenv("PBRTGPU_VERTEX")) c.vertex = atoi(e);
        if (const char *e = getenv("PBRTGPU_WAVE")) c.wave = atoi(e);
        if (const char *e = getenv("PBRTGPU_WSORT")) c.wsort = atoi(e) != 0;
        if (const char *e = getenv("PBRTGPU_TAIL_SORT")) c.tailSort = atoi(e);
        if (const char *e = getenv("PBRTGPU_LOCKSTEP")) c.lockstep = atoi(e);
        if (const char *e = getenv("PBRTGPU_POOL")) { if (atol(e) > 0) c.pool = (size_t)atol(e); }
        if (const char *e = getenv("PBRTGPU_LANES")) c.lanes = atoi(e);
        if (const char *e = getenv("PBRTGPU_VERTEX_BLOCKS")) { if (atoi(e) > 0) c.vertexBlocksPerSM = atoi(e); }
        c.traceIters = getenv("PBRTGPU_TRACE_ITERS") != nullptr;
        if (!c.wavefront) c.wavefront2 = false;
        return c;
    }
}

size_t ScratchBudgetBytes() {
    const char *env = getenv("PBRTGPU_SCRATCH_GB");
    double gb = env ? atof(env) : 40.0;
    if (gb < 0.25) gb = 0.25;
    return (size_t)(gb * 1024.0 * 1024.0 * 1024.0);
}

}  // namespace

extern "C" {

// (internal, for the library's other translation units: records the thread's last-error text)
int pbrtgpu_set_error_text(int code, const char *text) { return Fail(code, text ? text : ""); }

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

// Everything pbrtgpu_scene_create / _update promise to answer with PBRTGPU_EINVAL / EUNSUPPORTED, checked BEFORE any device
// buffer is touched: a failed update leaves the handle exactly as it was.  Every index the kernels follow is
// range-checked here, so a consumer other than the in-tree host front end cannot make them read out of bounds.
static int ValidateSceneDesc(const PbrtGpuSceneDesc *desc) {
    if (desc->nNodes == 0 || desc->nPrimitives == 0) return Fail(PBRTGPU_EINVAL, "empty scene");
    if (!desc->nodes || !desc->primitives) return Fail(PBRTGPU_EINVAL, "null node / primitive array");
    if (desc->nMaterials > 65535 || desc->nLights > 65534) return Fail(PBRTGPU_EUNSUPPORTED, "too many materials/lights");
    if (desc->nAnimated > 255) return Fail(PBRTGPU_EUNSUPPORTED, "more than 255 TransformedPrimitives");
    for (uint32_t i = 0; i < desc->nLights; ++i) {
        const PbrtGpuLight &l = desc->lights[i];
#if PBRT_INFINITE
        if (l.type == PBRTGPU_LIGHT_INFINITE) {
            if (l.firstShape >= desc->nXforms) return Fail(PBRTGPU_EINVAL, "infinite light's transform index out of range");
            if (l.nShapes >= desc->nEnvMaps || !desc->envMaps || !desc->envData) return Fail(PBRTGPU_EINVAL, "infinite light's environment map index out of range");
            const PbrtGpuEnvMap &m = desc->envMaps[l.nShapes];
            auto pow2 = [](uint32_t v) { return v && !(v & (v - 1)); };
            if (!pow2(m.width) || !pow2(m.height) || m.width > 65536 || m.height > 65536)
                return Fail(PBRTGPU_EINVAL, "environment map: level 0 must be a power of two in both directions");
            if (m.nu < 1 || m.nv < 1 || m.nu > 65536 || m.nv > 65536) return Fail(PBRTGPU_EINVAL, "environment map: bad distribution size");
            const uint64_t n = desc->nEnvData, nu = m.nu, nv = m.nv;
            if (m.texels + 3ull * m.width * m.height > n || m.func + nu * nv > n || m.cdf + nv * (nu + 1) > n || m.funcInt + nv > n ||
                m.margCdf + nv + 1 > n)
                return Fail(PBRTGPU_EINVAL, "environment map: array range out of bounds");
        }
#else
        if (l.type == PBRTGPU_LIGHT_INFINITE)
            return Fail(PBRTGPU_EUNSUPPORTED, "infinite area lights: this library was built with PBRT_INFINITE=0 (the device code has not run on a GPU yet)");
#endif
        if (l.type > PBRTGPU_LIGHT_INFINITE) return Fail(PBRTGPU_EINVAL, "unknown light type");
        if (l.type == PBRTGPU_LIGHT_DIFFUSE_AREA) {
            if (l.nShapes != 1) return Fail(PBRTGPU_EUNSUPPORTED, "area lights over more than one shape are out of scope");
            if ((uint64_t)l.firstShape + l.nShapes > desc->nLightShapes) return Fail(PBRTGPU_EINVAL, "light shape range out of bounds");
        }
    }
    for (uint32_t i = 0; i < desc->nLightShapes; ++i)
        if (desc->lightShapes[i].index >= desc->nQuadrics) return Fail(PBRTGPU_EINVAL, "light shape index out of range");
    for (uint32_t i = 0; i < desc->nMaterials; ++i) {
        const PbrtGpuMaterial &m = desc->materials[i];
#if !PBRT_SPECULAR
        if (m.type == PBRTGPU_MAT_MIRROR || m.type == PBRTGPU_MAT_GLASS)
            return Fail(PBRTGPU_EUNSUPPORTED, "specular materials (mirror, glass): this library was built with PBRT_SPECULAR=0");
#endif
#if !PBRT_METAL
        if (m.type == PBRTGPU_MAT_METAL)
            return Fail(PBRTGPU_EUNSUPPORTED, "metal material: this library was built with PBRT_METAL=0 (the device code has not run on a GPU yet)");
#endif
        if (m.type > PBRTGPU_MAT_METAL) return Fail(PBRTGPU_EINVAL, "unknown material type");
        if (m.type == PBRTGPU_MAT_MATTE && m.sigma != 0.f) return Fail(PBRTGPU_EUNSUPPORTED, "Oren-Nayar (sigma != 0) is out of scope");
        if (m.type == PBRTGPU_MAT_MEASURED && (m.measured < 0 || (uint32_t)m.measured >= desc->nMeasured))
            return Fail(PBRTGPU_EINVAL, "measured material without data");
    }
    for (uint32_t i = 0; i < desc->nMeasured; ++i) {
        const PbrtGpuMeasured &m = desc->measured[i];
        if ((uint64_t)m.firstNode + m.nNodes > desc->nKdNodes || m.nNodes == 0) return Fail(PBRTGPU_EINVAL, "measured kd-tree range out of bounds");
        for (uint32_t k = 0; k < m.nNodes; ++k) {
            const uint32_t bits = desc->kdNodes[m.firstNode + k].bits;
            // (rightChild >= nNodes means "no right child", kdtree.h:56; a left child is the next node)
            if ((bits & 3u) != 3u && ((bits >> 2) & 1u) && k + 1 >= m.nNodes) return Fail(PBRTGPU_EINVAL, "kd node's left child out of range");
            // children come after their parent (KdTree::recursiveBuild numbers the nodes in preorder, kdtree.h:93-118):
            // the device walks rely on it to terminate
            if ((bits & 3u) != 3u && (bits >> 3) < m.nNodes && (bits >> 3) <= k) return Fail(PBRTGPU_EINVAL, "kd node's right child does not follow it");
        }
        {   // depth: the evaluation kernels keep two stack entries per level (k_weval: WEVAL_STACK, dshade.cuh KdGather: 40)
            std::vector<uint8_t> depth(m.nNodes, 0);
            depth[0] = 1;
            for (uint32_t k = 0; k < m.nNodes; ++k) {
                const uint32_t bits = desc->kdNodes[m.firstNode + k].bits;
                if (depth[k] == 0) continue;   // (unreachable node)
                if (depth[k] > 18) return Fail(PBRTGPU_EUNSUPPORTED, "measured-BRDF kd-tree deeper than 18 levels (not a median-split tree)");
                if ((bits & 3u) == 3u) continue;
                if ((bits >> 2) & 1u) depth[k + 1] = (uint8_t)(depth[k] + 1);
                if ((bits >> 3) < m.nNodes) depth[bits >> 3] = (uint8_t)(depth[k] + 1);
            }
        }
    }
    for (uint32_t i = 0; i < desc->nMeshes; ++i) {
        if (desc->meshes[i].xform >= desc->nXforms) return Fail(PBRTGPU_EINVAL, "mesh transform index out of range");
        if ((desc->meshes[i].flags & PBRTGPU_MESH_HAS_N) && !desc->N) return Fail(PBRTGPU_EINVAL, "mesh flags N but desc.N is null");
        if ((desc->meshes[i].flags & PBRTGPU_MESH_HAS_UV) && !desc->uv) return Fail(PBRTGPU_EINVAL, "mesh flags uv but desc.uv is null");
    }
    for (uint32_t i = 0; i < desc->nQuadrics; ++i)
        if (desc->quadrics[i].xform >= desc->nXforms) return Fail(PBRTGPU_EINVAL, "quadric transform index out of range");
    for (uint32_t i = 0; i < desc->nTriangles; ++i) {
        const PbrtGpuTriangle &t = desc->triangles[i];
        if (t.mesh >= desc->nMeshes) return Fail(PBRTGPU_EINVAL, "triangle mesh index out of range");
        for (int k = 0; k < 3; ++k)
            if (t.v[k] >= desc->nVertices) return Fail(PBRTGPU_EINVAL, "vertex index out of range");
    }
    std::vector<uint8_t> nested(desc->nPrimitives, 0);
    std::vector<uint32_t> nodeBase(desc->nNodes, 0u), nodeEnd(desc->nNodes, 0u), primBase(desc->nNodes, 0u), primEnd(desc->nNodes, 0u);
    uint32_t topNodes = desc->nNodes, topPrims = desc->nPrimitives;
    for (uint32_t k = 0; k < desc->nAnimated; ++k) {
        const PbrtGpuAnimated &a = desc->animated[k];
        {   // every number the interpolation reads must be one (AnimatedTransform::Interpolate, transform.cpp:437-465)
            bool finite = std::isfinite(a.startTime) && std::isfinite(a.endTime);
            for (int i = 0; i < 16; ++i)
                finite = finite && std::isfinite(a.start.m[i]) && std::isfinite(a.start.mInv[i]) && std::isfinite(a.end.m[i]) && std::isfinite(a.end.mInv[i]);
            for (int e = 0; e < 2; ++e) {
                for (int i = 0; i < 3; ++i) finite = finite && std::isfinite(a.T[e][i]);
                for (int i = 0; i < 4; ++i) finite = finite && std::isfinite(a.R[e][i]);
                for (int i = 0; i < 16; ++i) finite = finite && std::isfinite(a.S[e][i]);
            }
            if (!finite) return Fail(PBRTGPU_EINVAL, "TransformedPrimitive with a non-finite transform or time");
        }
        if ((uint64_t)a.firstPrimitive + a.nPrimitives > desc->nPrimitives || (uint64_t)a.firstNode + a.nNodes > desc->nNodes ||
            a.nPrimitives == 0 || (a.nNodes == 0 && a.nPrimitives != 1))
            return Fail(PBRTGPU_EINVAL, "TransformedPrimitive range out of bounds");
        for (uint32_t i = 0; i < a.nPrimitives; ++i) {
            // a primitive record names the one TransformedPrimitive it lies under (dscene.cuh HitGeometry): instances
            // that share an aggregate must come with a copy each
            if (nested[a.firstPrimitive + i]) return Fail(PBRTGPU_EUNSUPPORTED, "TransformedPrimitives with overlapping primitive ranges (a shared instance aggregate)");
            nested[a.firstPrimitive + i] = 1;
        }
        for (uint32_t i = 0; i < a.nNodes; ++i) {
            nodeBase[a.firstNode + i] = a.firstNode; nodeEnd[a.firstNode + i] = a.firstNode + a.nNodes;
            primBase[a.firstNode + i] = a.firstPrimitive; primEnd[a.firstNode + i] = a.firstPrimitive + a.nPrimitives;
        }
        if (a.nNodes) topNodes = std::min(topNodes, a.firstNode);
        topPrims = std::min(topPrims, a.firstPrimitive);
    }
    for (uint32_t i = 0; i < topNodes; ++i) { nodeEnd[i] = topNodes; primEnd[i] = topPrims; }
    // depth of every node below its BVH's root (0: not reached yet).  flattenBVHTree numbers the nodes in preorder
    // (bvh.cpp:383-399): both children come after their parent, which is what makes one forward pass enough here and
    // what lets the walks terminate; the walks keep one stack entry per level (todo[64], bvh.cpp:410, as in the reference)
    std::vector<uint8_t> depth(desc->nNodes, 0);
    depth[0] = 1;
    for (uint32_t k = 0; k < desc->nAnimated; ++k)
        if (desc->animated[k].nNodes) depth[desc->animated[k].firstNode] = 1;
    for (uint32_t i = 0; i < desc->nNodes; ++i) {   // child / primitive offsets are relative to the BVH the node belongs to
        const PbrtGpuBVHNode &n = desc->nodes[i];
        if (nodeEnd[i] == 0) continue;   // (a node outside every BVH range is never reached)
        if (n.nPrimitives) {
            if ((uint64_t)primBase[i] + n.offset + n.nPrimitives > primEnd[i]) return Fail(PBRTGPU_EINVAL, "BVH leaf's primitive range out of bounds");
        } else {
            if ((uint64_t)nodeBase[i] + n.offset >= nodeEnd[i] || i + 1 >= nodeEnd[i]) return Fail(PBRTGPU_EINVAL, "BVH child offset out of bounds");
            if (n.axis > 2) return Fail(PBRTGPU_EINVAL, "BVH split axis out of range");
            const uint32_t second = nodeBase[i] + n.offset;
            if (second <= i + 1) return Fail(PBRTGPU_EINVAL, "BVH second child does not follow the first child's subtree");
            if (depth[i]) {
                if (depth[i] >= 64) return Fail(PBRTGPU_EUNSUPPORTED, "BVH deeper than 64 levels (the traversal stack of the reference and of the kernels)");
                depth[i + 1] = std::max<uint8_t>(depth[i + 1], (uint8_t)(depth[i] + 1));
                depth[second] = std::max<uint8_t>(depth[second], (uint8_t)(depth[i] + 1));
            }
        }
    }
    for (uint32_t i = 0; i < desc->nPrimitives; ++i) {
        const PbrtGpuPrimitive &p = desc->primitives[i];
        if (p.type == PBRTGPU_PRIM_ANIMATED) {
            if (p.index >= desc->nAnimated) return Fail(PBRTGPU_EINVAL, "animated index out of range");
            if (nested[i]) return Fail(PBRTGPU_EUNSUPPORTED, "nested TransformedPrimitives are out of scope");
            continue;
        }
        if (p.type > PBRTGPU_PRIM_ANIMATED) return Fail(PBRTGPU_EINVAL, "unknown primitive type");
        if (p.material < 0 || (uint32_t)p.material >= desc->nMaterials) return Fail(PBRTGPU_EINVAL, "primitive without material");
        if (p.areaLight < -1 || p.areaLight >= (int)desc->nLights) return Fail(PBRTGPU_EINVAL, "primitive's area light out of range");
        if (p.type == PBRTGPU_PRIM_TRIANGLE) {
            if (p.index >= desc->nTriangles) return Fail(PBRTGPU_EINVAL, "triangle index out of range");
        } else if (p.index >= desc->nQuadrics) return Fail(PBRTGPU_EINVAL, "quadric index out of range");
    }
    return PBRTGPU_OK;
}

// copies a validated `desc` into the scene's device buffers (reused when large enough).  A CUDA failure half way leaves the
// handle marked invalid (render / intersect refuse it) until a later update succeeds.
static int UploadScene(PbrtGpuScene *s, const PbrtGpuSceneDesc *desc) {
    int vrc = ValidateSceneDesc(desc);
    if (vrc != PBRTGPU_OK) return vrc;
    s->valid = false;

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
    {
        std::vector<float4> top;
        BuildTopNodes(desc, PBRT_TOP_LEVELS, &top);
        CUDA_TRY(s->topNodes.Upload(top.data(), top.size()));
    }
#if PBRT_PAIR_NODES
    {
        std::vector<float4> pairs;
        std::vector<uint32_t> pairOfNode;
        BuildNodePairs(desc, &pairs, &pairOfNode);
        CUDA_TRY(s->pairs.Upload(pairs.data(), pairs.size()));
        CUDA_TRY(s->pairOfNode.Upload(pairOfNode.data(), pairOfNode.size()));
    }
#endif
    // primitives: pre-gathered 48-byte records in leaf order
    {
        std::vector<float4> h(3 * (size_t)desc->nPrimitives);
        // which TransformedPrimitive (if any) owns each nested primitive
        std::vector<int> owner(desc->nPrimitives, -1);
        for (uint32_t k = 0; k < desc->nAnimated; ++k) {
            const PbrtGpuAnimated &a = desc->animated[k];
            for (uint32_t i = 0; i < a.nPrimitives; ++i) owner[a.firstPrimitive + i] = (int)k;
        }
        for (uint32_t i = 0; i < desc->nPrimitives; ++i) {
            const PbrtGpuPrimitive &p = desc->primitives[i];
            float v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (p.type == PBRTGPU_PRIM_ANIMATED) {
                uint32_t w0 = p.type, w1 = p.index;
                float f0, f1;
                memcpy(&f0, &w0, 4); memcpy(&f1, &w1, 4);
                h[3 * i] = make_float4(0, 0, 0, f0);
                h[3 * i + 1] = make_float4(0, 0, 0, f1);
                h[3 * i + 2] = make_float4(0, 0, 0, 0);
                continue;
            }
            if (p.type == PBRTGPU_PRIM_TRIANGLE) {
                const PbrtGpuTriangle &t = desc->triangles[p.index];
                for (int k = 0; k < 3; ++k) memcpy(v + 3 * k, desc->P + 3 * (size_t)t.v[k], 12);
            }
            uint32_t w0 = p.type | ((uint32_t)(owner[i] + 1) << 8), w1 = p.index;
            uint32_t w2 = (uint32_t)p.material | ((uint32_t)(p.areaLight + 1) << 16);
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
#if PBRT_KD_INTERLEAVED
    {
        std::vector<float4> rec;
        BuildKdRecords(desc, &rec);
        CUDA_TRY(s->kdRec.Upload(rec.data(), rec.size()));
    }
#endif
    CUDA_TRY(s->lights.Upload(desc->lights, desc->nLights));
#if PBRT_INFINITE
    CUDA_TRY(s->envMaps.Upload(desc->envMaps, desc->nEnvMaps));
    CUDA_TRY(s->envData.Upload(desc->envData, desc->nEnvData));
#endif
    CUDA_TRY(s->lightShapes.Upload(desc->lightShapes, desc->nLightShapes));
    CUDA_TRY(s->animated.Upload(desc->animated, desc->nAnimated));
    DScene &d = s->d;
    d.nodes = s->nodes.p; d.prims = s->prims.p; d.tris = s->tris.p; d.meshes = s->meshes.p;
    d.topNodes = s->topNodes.p;
#if PBRT_PAIR_NODES
    d.pairs = s->pairs.p; d.pairOfNode = s->pairOfNode.p;
#endif
#if PBRT_KD_INTERLEAVED
    d.kdRec = s->kdRec.p;
#endif
    d.N = s->N.p; d.uv = s->uv.p; d.quadrics = s->quadrics.p; d.xforms = s->xforms.p;
    d.materials = s->materials.p; d.measured = s->measured.p; d.kdNodes = s->kdNodes.p; d.kdData = s->kdData.p;
    d.lights = s->lights.p; d.lightShapes = s->lightShapes.p; d.animated = s->animated.p;
#if PBRT_INFINITE
    d.envMaps = s->envMaps.p; d.envData = s->envData.p;
#endif
    d.nNodes = desc->nNodes; d.nPrims = desc->nPrimitives; d.nLights = desc->nLights; d.nAnimated = desc->nAnimated;
    s->nMeasuredMaterials = 0;
    s->nSpecularMaterials = 0;
    for (uint32_t i = 0; i < desc->nMaterials; ++i) {
        if (desc->materials[i].type == PBRTGPU_MAT_MEASURED) s->nMeasuredMaterials++;
        if (desc->materials[i].type == PBRTGPU_MAT_MIRROR || desc->materials[i].type == PBRTGPU_MAT_GLASS) s->nSpecularMaterials++;
    s->nInfiniteLights = 0;
    for (uint32_t i = 0; i < desc->nLights; ++i)
        if (desc->lights[i].type == PBRTGPU_LIGHT_INFINITE) s->nInfiniteLights++;
    }
    s->valid = true;
    return PBRTGPU_OK;
}

// what a render call checks once it knows the scene: the Sample layout against the scene's lights, recursion depth
// against the frames a path slot carries
static int ValidateRenderForScene(const PbrtGpuRenderDesc &rdr, uint32_t nLights, uint32_t nSpecularMaterials) {
    const PbrtGpuRenderDesc *rd = &rdr;
    const bool directInteg = rd->integrator == PBRTGPU_INTEGRATOR_DIRECTLIGHTING;
#if PBRT_SPECULAR
    if (directInteg && nSpecularMaterials && rd->maxDepth > DL_MAX_DEPTH)
        return Fail(PBRTGPU_EUNSUPPORTED, "DirectLightingIntegrator over specular materials: maxdepth beyond the 8 recursion frames a path slot carries");
    if (nSpecularMaterials && nLights > 32767) return Fail(PBRTGPU_EUNSUPPORTED, "too many lights");
#else
    (void)nSpecularMaterials;
#endif
    if (directInteg) {   // the layout must be what DirectLightingIntegrator::RequestSamples builds for this scene's lights
        const int nl = (int)nLights;
        const bool all = rd->lightStrategy == PBRTGPU_STRATEGY_ALL;
        if (rd->lightStrategy != PBRTGPU_STRATEGY_ALL && rd->lightStrategy != PBRTGPU_STRATEGY_ONE)
            return Fail(PBRTGPU_EINVAL, "bad lightStrategy");
        if (all ? (rd->n1D != 2 * nl + 2 || rd->n2D != 2 * nl) : (rd->n1D != 5 || rd->n2D != 2))
            return Fail(PBRTGPU_EINVAL, "Sample layout does not match DirectLightingIntegrator::RequestSamples for this scene (directlighting.cpp:53-78)");
        for (int i = 0; i < (all ? nl : 0); ++i)
            if (rd->size1D[2 * i] != rd->size1D[2 * i + 1] || rd->size1D[2 * i] != rd->size2D[2 * i] || rd->size1D[2 * i] != rd->size2D[2 * i + 1])
                return Fail(PBRTGPU_EINVAL, "a light's four sample arrays must have the same size (light.cpp:64-68, reflection.cpp:502-506)");
    }
    return PBRTGPU_OK;
}

int pbrtgpu_render_validate_scene(const PbrtGpuSceneDesc *desc, const PbrtGpuRenderDesc *rd) {
    if (!desc || !rd) return Fail(PBRTGPU_EINVAL, "null argument");
    int rc = ValidateSceneDesc(desc);
    if (rc != PBRTGPU_OK) return rc;
    rc = ValidateRenderDesc(*rd);
    if (rc != PBRTGPU_OK) return rc;
    uint32_t nSpecular = 0;
    for (uint32_t i = 0; i < desc->nMaterials; ++i)
        if (desc->materials[i].type == PBRTGPU_MAT_MIRROR || desc->materials[i].type == PBRTGPU_MAT_GLASS) ++nSpecular;
    return ValidateRenderForScene(*rd, desc->nLights, nSpecular);
}

int pbrtgpu_render_validate(const PbrtGpuRenderDesc *rd) {
    if (!rd) return Fail(PBRTGPU_EINVAL, "null argument");
    return ValidateRenderDesc(*rd);
}

int pbrtgpu_film_extent(const PbrtGpuRenderDesc *rd, int32_t extent[4]) {
    if (!rd || !extent) return Fail(PBRTGPU_EINVAL, "null argument");
    const int rc = ValidateRenderDesc(*rd);
    if (rc != PBRTGPU_OK) return rc;
    const FilmExtent e = ComputeExtent(*rd);
    extent[0] = e.xPixelStart; extent[1] = e.yPixelStart; extent[2] = e.xPixelCount; extent[3] = e.yPixelCount;
    return PBRTGPU_OK;
}

int pbrtgpu_scene_validate(const PbrtGpuSceneDesc *desc) {
    if (!desc) return Fail(PBRTGPU_EINVAL, "null argument");
    return ValidateSceneDesc(desc);
}

int pbrtgpu_scene_create(const PbrtGpuSceneDesc *desc, PbrtGpuScene **out) {
    if (!desc || !out) return Fail(PBRTGPU_EINVAL, "null argument");
    *out = nullptr;
    if (pbrtgpu_device_count() <= 0)
        return Fail(PBRTGPU_ENODEVICE, "no CUDA device visible; libpbrtgpu has no CPU fallback");
    std::unique_ptr<PbrtGpuScene> s(new PbrtGpuScene);
    CUDA_TRY(cudaGetDevice(&s->device));
    int rc = UploadScene(s.get(), desc);
    if (rc != PBRTGPU_OK) return rc;
    s->collectStats = getenv("PBRTGPU_STATS") != nullptr ? 1 : 0;
    for (int i = 0; i < 3; ++i) CUDA_TRY(cudaEventCreate(&s->ev[i]));
    *out = s.release();
    return PBRTGPU_OK;
}

int pbrtgpu_scene_update(PbrtGpuScene *scene, const PbrtGpuSceneDesc *desc) {
    if (!scene || !desc) return Fail(PBRTGPU_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(scene->device));
    return UploadScene(scene, desc);
}

void pbrtgpu_scene_destroy(PbrtGpuScene *scene) {
    if (!scene) return;
    for (int i = 0; i < 3; ++i)
        if (scene->ev[i]) cudaEventDestroy(scene->ev[i]);
    for (int l = 0; l < PbrtGpuScene::MAX_LANES; ++l)
        if (scene->lanes[l]) ReleaseLane(scene->device, scene->lanes[l]);
    delete scene;
}

// counting pass switch: PbrtGpuStats.nodesVisited etc. are only filled when the
// scene was created with PBRTGPU_STATS set or via pbrtgpu_set_stats(scene, 1); the
// timed path runs the kernel instantiation without counters.
int pbrtgpu_set_stats(PbrtGpuScene *scene, int on) {
    if (!scene) return Fail(PBRTGPU_EINVAL, "null scene");
    scene->collectStats = on;
    return PBRTGPU_OK;
}

}  // extern "C"

namespace {

// One lane = one CUDA stream + one set of scratch buffers + one host thread.  A render call deals its task
// batches round-robin to the lanes; batches share nothing (own RNG streams, own pixels; edge splats are atomic),
// so the lanes' kernels overlap on the device and fill the SMs the latency-bound tail / traversal kernels of a
// single batch leave idle.
struct LaneJob {
    PbrtGpuScene *scene;
    Lane *ln;
    const std::vector<TaskRec> *tasks;
    size_t maxBatch;
    int laneIndex, nLanes;   // this lane renders batches laneIndex, laneIndex + nLanes, ...
    RenderParams rp;
    int spp;
    PbrtGpuPixel *film;
    bool timing, wantCounters, timeRays;
    RenderConfig cfg;
    TableArrays tables;      // the Sample layout, array by array
    bool pathLayout;         // ... is PathIntegrator's: the specialised table kernel applies
    // results
    int rc = PBRTGPU_OK;
    std::string err;
    float tableMs = 0.f, traceMs = 0.f, tailMs = 0.f, rayMs = 0.f;
    uint64_t launches = 0, rayLaunches = 0;
};

int RunLane(LaneJob &job) {
    PbrtGpuScene *scene = job.scene;
    Lane *ln = job.ln;
    CUDA_TRY(cudaSetDevice(scene->device));   // a new host thread starts on device 0
    cudaStream_t stream = ln->stream;
    const std::vector<TaskRec> &tasks = *job.tasks;
    const int spp = job.spp;
    const RenderParams &rp = job.rp;
    PbrtGpuPixel *film = job.film;
    const bool stats = job.timing, wantCounters = job.wantCounters, timeRays = job.timeRays;
    const size_t maxBatch = job.maxBatch;
    std::vector<cudaEvent_t> rayEvents;   // timeRays: a pair of events around every k_trace launch of a round
    size_t rayEventsUsed = 0;
    float rayMs = 0.f;
    uint64_t rayLaunches = 0;
    auto rayEvent = [&]() -> cudaEvent_t {
        if (rayEventsUsed == rayEvents.size()) {
            cudaEvent_t e = nullptr;
            cudaEventCreate(&e);
            rayEvents.push_back(e);
        }
        return rayEvents[rayEventsUsed++];
    };
    CUDA_TRY(cudaStreamWaitEvent(stream, scene->ev[2], 0));   // the film has been cleared
    CUDA_TRY(ln->tasks.Ensure(maxBatch));
    CUDA_TRY(ln->perm.Ensure(maxBatch * spp * (size_t)rp.lay.rowStride));
    if (!job.pathLayout) CUDA_TRY(ln->tgt.Ensure(maxBatch * spp * (size_t)job.tables.tgtPerSample + 16));
    CUDA_TRY(ln->scr.Ensure(maxBatch * SCR_STRIDE));
    CUDA_TRY(ln->rngCur.Ensure(maxBatch * MT_N));
    CUDA_TRY(ln->rngNext.Ensure(maxBatch * MT_N));
    CUDA_TRY(ln->mti.Ensure(maxBatch));
    CUDA_TRY(ln->ls.Ensure(maxBatch * spp));
    CUDA_TRY(ln->rec.Ensure(rp.integrator == PBRTGPU_INTEGRATOR_DIRECTLIGHTING ? 4 : maxBatch * spp * 4));   // parked bounce-3 vertices: PathIntegrator only
    CUDA_TRY(ln->tailCount.Ensure(maxBatch));
    CUDA_TRY(ln->tailKeysOut.Ensure(maxBatch));
    CUDA_TRY(ln->orderIn.Ensure(maxBatch));
    CUDA_TRY(ln->orderOut.Ensure(maxBatch));
    {
        std::vector<unsigned> iota(maxBatch);
        for (size_t i = 0; i < maxBatch; ++i) iota[i] = (unsigned)i;
        CUDA_TRY(cudaMemcpyAsync(ln->orderIn.p, iota.data(), maxBatch * sizeof(unsigned), cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
    }
    size_t sortBytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, sortBytes, ln->tailCount.p, ln->tailKeysOut.p,
                                              ln->orderIn.p, ln->orderOut.p, (int)maxBatch, 0, 14, stream);
    CUDA_TRY(ln->sortTemp.Ensure(sortBytes + 16));

    // PathIntegrator: both phases by the vertex-synchronous kernel (k_vertex) unless PBRTGPU_VERTEX says otherwise;
    // DirectLightingIntegrator: the wavefront (k_logic / k_eval / k_trace)
    const RenderConfig &cfg = job.cfg;
    const bool direct = rp.integrator == PBRTGPU_INTEGRATOR_DIRECTLIGHTING;
    const bool vertex1 = !direct && (cfg.vertex & 1), vertex2 = !direct && (cfg.vertex & 2);
    const bool wave1 = !direct && (cfg.wave & 1) && cfg.wavefront && !vertex1;
    // phase 2 on the vertex wavefront (its slot word keeps the bounce index in 8 bits: deeper paths stay with k_vertex<2>)
    const bool wave2 = !direct && (cfg.wave & 2) && cfg.wavefront && rp.maxDepth <= 250;   // phase 1 on the vertex wavefront (kwave.cuh)
    const bool useWavefront = cfg.wavefront && !vertex1 && !wave1, useWavefront2 = cfg.wavefront2 && useWavefront && !vertex2;
    size_t poolSize = cfg.pool;   // 4 Mi slots x 272 B = 1.1 GB (measured: +17 % over 1 Mi)
    const bool dlSpecular = direct && scene->nSpecularMaterials > 0;
    if (dlSpecular) poolSize = std::min<size_t>(poolSize, (size_t)1 << 20);   // + 960 B of recursion frames per slot
    poolSize = (poolSize + 127) / 128 * 128;
    const int maskWords = std::max(1, spp / 32);
    CUDA_TRY(ln->tailMask.Ensure(maxBatch * (size_t)maskWords));
    CUDA_TRY(ln->tailFlag.Ensure(maxBatch));
    CUDA_TRY(ln->workCounter.Ensure(4));
    int nSM = 148;
    cudaDeviceGetAttribute(&nSM, cudaDevAttrMultiProcessorCount, scene->device);
    bool hasMeasured = scene->measured.n > 0 && scene->nMeasuredMaterials > 0;
    if (useWavefront) {
        CUDA_TRY(ln->pool.Ensure(poolSize * F_COUNT));
        CUDA_TRY(ln->traceQ.Ensure(poolSize * 2 * K_COUNT));   // request queues: 2 sets x K_COUNT kinds
        CUDA_TRY(ln->counters.Ensure(C_COUNT));
        if (dlSpecular) CUDA_TRY(ln->dlFrames.Ensure(poolSize * DL_MAX_DEPTH));
        if (!ln->hostCounters) CUDA_TRY(cudaMallocHost((void **)&ln->hostCounters, C_COUNT * sizeof(unsigned)));
    }
    if (wave1 || wave2) {
        CUDA_TRY(ln->pool.Ensure(std::max(poolSize, maxBatch) * std::max<size_t>(scene->d.nAnimated ? W_COUNT_ANIM : W_COUNT, F_COUNT)));
        CUDA_TRY(ln->traceQ.Ensure(std::max(poolSize, maxBatch) * 3));
        CUDA_TRY(ln->evalQ.Ensure(std::max(poolSize, maxBatch) * 3));
        CUDA_TRY(ln->liveList.Ensure(maxBatch * 2));
        if (cfg.wsort) {
            const size_t nClass = (poolSize + 1023) / 1024 * 1024;
            const bool fresh = ln->sortClass.n < nClass;
            CUDA_TRY(ln->sortList.Ensure(poolSize * 2));
            CUDA_TRY(ln->sortClass.Ensure(nClass));
            if (fresh) CUDA_TRY(cudaMemsetAsync(ln->sortClass.p, 0, ln->sortClass.n, stream));
        }
        CUDA_TRY(ln->counters.Ensure(std::max<int>(WC_COUNT, C_COUNT)));
        if (!ln->hostCounters) CUDA_TRY(cudaMallocHost((void **)&ln->hostCounters, std::max<int>(WC_COUNT, C_COUNT) * sizeof(unsigned)));
    }
    const bool traceIters = cfg.traceIters;   // debugging aid: queue occupancy on stderr
    const TableLaunch tl = job.pathLayout ? PlanTables(spp) : PlanTablesGeneral(spp, job.tables.nArrays);
    if (direct && !useWavefront) return Fail(PBRTGPU_EUNSUPPORTED, "DirectLightingIntegrator runs on the wavefront only (PBRTGPU_WAVEFRONT=0 is a PathIntegrator A/B switch)");
    uint64_t launches = 0;
    float tableMs = 0.f, traceMs = 0.f, tailMs = 0.f;
    for (size_t bi = (size_t)job.laneIndex; bi * maxBatch < tasks.size(); bi += (size_t)job.nLanes) {
        const size_t begin = bi * maxBatch;
        const size_t nb = std::min(maxBatch, tasks.size() - begin);
        int rounds = 0;
        for (size_t i = 0; i < nb; ++i)
            rounds = std::max(rounds, (tasks[begin + i].x1 - tasks[begin + i].x0) * (tasks[begin + i].y1 - tasks[begin + i].y0));
        CUDA_TRY(cudaMemcpyAsync(ln->tasks.p, tasks.data() + begin, nb * sizeof(TaskRec), cudaMemcpyHostToDevice, stream));
        BatchArgs b;
        b.tasks = ln->tasks.p; b.nBatch = (int)nb; b.round = 0;
        b.perm = ln->perm.p; b.scr = ln->scr.p; b.rngCur = ln->rngCur.p; b.rngNext = ln->rngNext.p;
        b.mti = ln->mti.p; b.ls = ln->ls.p; b.rec = ln->rec.p;
        b.tailCount = ln->tailCount.p; b.order = nullptr;
        b.tailMask = ln->tailMask.p; b.maskWords = maskWords;
        for (int r = 0; r < rounds; ++r) {
            b.round = r;
            if (stats) CUDA_TRY(cudaEventRecord(ln->ev[0], stream));
            if (job.pathLayout) CUDA_TRY(LaunchTablesAny(b, spp, tl, stream));
            else CUDA_TRY(LaunchTablesGeneralAny(b, spp, tl, job.tables, ln->tgt.p, stream));
            if (stats) CUDA_TRY(cudaEventRecord(ln->ev[1], stream));
            CUDA_TRY(cudaMemsetAsync(ln->tailCount.p, 0, nb * sizeof(unsigned), stream));
            CUDA_TRY(cudaMemsetAsync(ln->tailMask.p, 0, nb * (size_t)maskWords * sizeof(uint32_t), stream));
            PathArgs pa;
            pa.sc = scene->d; pa.rp = rp; pa.b = b; pa.film = film; pa.stats = scene->stats.p;
            pa.directFilm = rounds > 1 ? 1 : 0;
#if PBRT_FILTER
            if (rp.filterTable) pa.directFilm = 1;   // no own-pixel accumulator: the pixel may lie outside the film
#endif
            pa.lockstep = cfg.lockstep;
            VxArgs va;
            va.sc = scene->d; va.rp = rp; va.b = b; va.workCounter = ln->workCounter.p;
            va.totalSamples = (unsigned)(nb * (size_t)spp); va.sortedCount = nullptr; va.stats = scene->stats.p;
            const int vblocks = nSM * cfg.vertexBlocksPerSM;   // persistent: every warp pulls work until none is left
            const int blocks1 = (int)((nb + 3) / 4);       // phase 1 (megakernel form): one warp per task-pixel
            const int blocks2 = (int)((nb + 127) / 128);   // phase 2: one thread per task-pixel
            if (vertex1) {
                if (timeRays) CUDA_TRY(cudaEventRecord(rayEvent(), stream));
                CUDA_TRY(cudaMemsetAsync(ln->workCounter.p, 0, sizeof(unsigned), stream));
                if (scene->d.nAnimated) {
                    if (wantCounters) k_vertex<1, true, true><<<vblocks, 128, 0, stream>>>(va);
                    else k_vertex<1, false, true><<<vblocks, 128, 0, stream>>>(va);
                } else {
                    if (wantCounters) k_vertex<1, true, false><<<vblocks, 128, 0, stream>>>(va);
                    else k_vertex<1, false, false><<<vblocks, 128, 0, stream>>>(va);
                }
                CUDA_TRY(cudaGetLastError());
                if (timeRays) CUDA_TRY(cudaEventRecord(rayEvent(), stream));
                launches += 1;
            } else if (wave1) {
                // ---- phase 1 as a vertex wavefront: one k_wshade pass per path vertex
                WvArgs wa;
                wa.sc = scene->d; wa.rp = rp; wa.b = b; wa.pool = ln->pool.p; wa.P = (uint32_t)poolSize;
                wa.qRay = ln->traceQ.p; wa.qEval = ln->evalQ.p; wa.counters = ln->counters.p;
                wa.totalSamples = (unsigned)(nb * (size_t)spp); wa.stats = scene->stats.p; wa.liveList = nullptr;
                wa.sortList = cfg.wsort ? ln->sortList.p : nullptr; wa.sortClass = ln->sortClass.p;
                if (cfg.wsort) CUDA_TRY(cudaMemsetAsync(ln->sortClass.p, 0, ln->sortClass.n, stream));   // (a batch that ended early leaves none, but be sure)
                CUDA_TRY(cudaMemsetAsync(ln->counters.p, 0, WC_COUNT * sizeof(unsigned), stream));
                const int sblocks = (int)((poolSize + 127) / 128);
                const int tblocks = nSM * 8;
                for (int iter = 0;; ++iter) {
                    wa.iter = iter;
                    // empty the queues k_wshade is about to fill, the consumers' cursors, the live-slot count and the sorted
                    // lists this iteration writes
                    CUDA_TRY(cudaMemsetAsync(ln->counters.p, 0, WC_NEXT_SAMPLE * sizeof(unsigned), stream));

                    if (scene->d.nAnimated) {
                        if (wantCounters) k_wshade<1, true, true><<<sblocks, 128, 0, stream>>>(wa);
                        else k_wshade<1, false, true><<<sblocks, 128, 0, stream>>>(wa);
                    } else {
                        if (wantCounters) k_wshade<1, true, false><<<sblocks, 128, 0, stream>>>(wa);
                        else k_wshade<1, false, false><<<sblocks, 128, 0, stream>>>(wa);
                    }
                    if (hasMeasured) k_weval<<<tblocks, 128, 0, stream>>>(wa);
                    if (timeRays) CUDA_TRY(cudaEventRecord(rayEvent(), stream));
                    if (scene->d.nAnimated) {
                        if (wantCounters) k_wtrace<true, true><<<tblocks, 128, 0, stream>>>(wa);
                        else k_wtrace<false, true><<<tblocks, 128, 0, stream>>>(wa);
                    } else {
                        if (wantCounters) k_wtrace<true, false><<<tblocks, 128, 0, stream>>>(wa);
                        else k_wtrace<false, false><<<tblocks, 128, 0, stream>>>(wa);
                    }
                    if (timeRays) CUDA_TRY(cudaEventRecord(rayEvent(), stream));
                    launches += hasMeasured ? 3 : 2;
                    if (cfg.wsort) {   // this iteration's class bytes -> the slot lists of the next
                        CUDA_TRY(cudaMemsetAsync(ln->counters.p + WC_SORT, 0, 2 * sizeof(unsigned), stream));
                        k_wcompact<<<(int)((poolSize + 1023) / 1024), WCOMPACT_THREADS, 0, stream>>>(wa);
                        launches += 1;
                    }
                    if ((iter & 7) == 7) {
                        CUDA_TRY(cudaMemcpyAsync(ln->hostCounters, ln->counters.p, WC_COUNT * sizeof(unsigned),
                                                 cudaMemcpyDeviceToHost, stream));
                        CUDA_TRY(cudaStreamSynchronize(stream));
                        if (traceIters)
                            fprintf(stderr, "[pbrtgpu] batch@%zu round %d iter %d: live %u rays %u/%u/%u evals %u next sample %u of %u\n",
                                    begin, r, iter, ln->hostCounters[WC_LIVE], ln->hostCounters[WC_RAYS], ln->hostCounters[WC_RAYS + 1],
                                    ln->hostCounters[WC_RAYS + 2], ln->hostCounters[WC_EVALS], ln->hostCounters[WC_NEXT_SAMPLE], wa.totalSamples);
                        if (ln->hostCounters[WC_LIVE] == 0) break;
                    }
                    if (iter > 4000000) return Fail(PBRTGPU_ECUDA, "vertex wavefront did not terminate");
                }
                CUDA_TRY(cudaGetLastError());
            } else if (!useWavefront) {
                if (scene->d.nAnimated) {
                    if (wantCounters) k_path<1, true, true><<<blocks1, 128, 0, stream>>>(pa);
                    else k_path<1, false, true><<<blocks1, 128, 0, stream>>>(pa);
                } else {
                    if (wantCounters) k_path<1, true, false><<<blocks1, 128, 0, stream>>>(pa);
                    else k_path<1, false, false><<<blocks1, 128, 0, stream>>>(pa);
                }
                CUDA_TRY(cudaGetLastError());
                launches += 1;
            } else {
                // ---- phase 1 as a wavefront over a pool of path slots
                WfArgs wa;
                wa.sc = scene->d; wa.rp = rp; wa.b = b; wa.pool = ln->pool.p; wa.P = (uint32_t)poolSize;
                wa.q = ln->traceQ.p; wa.counters = ln->counters.p;
                wa.totalSamples = (unsigned)(nb * (size_t)spp); wa.stats = scene->stats.p;
                wa.film = film; wa.directFilm = 0;
                wa.dlFrames = dlSpecular ? ln->dlFrames.p : nullptr;
                CUDA_TRY(cudaMemsetAsync(ln->counters.p, 0, C_COUNT * sizeof(unsigned), stream));
                const int lblocks = (int)((poolSize + 127) / 128);
                const int tblocks = nSM * 8;
                for (int iter = 0;; ++iter) {
                    wa.iter = iter;
                    const int setOut = (iter & 1) ^ 1;
                    // empty the queues k_logic is about to fill, and the consumers' cursors
                    CUDA_TRY(cudaMemsetAsync(ln->counters.p + C_QN + setOut * K_COUNT, 0, K_COUNT * sizeof(unsigned), stream));
                    CUDA_TRY(cudaMemsetAsync(ln->counters.p + C_TRACE_CUR, 0, 2 * sizeof(unsigned), stream));
                    if (direct) {   // DirectLightingIntegrator: one instantiation (TransformedPrimitives handled at run time)
                        if (wantCounters) k_logic<1, true, true, 1><<<lblocks, 128, 0, stream>>>(wa);
                        else k_logic<1, false, true, 1><<<lblocks, 128, 0, stream>>>(wa);
                    } else if (scene->d.nAnimated) {
                        if (wantCounters) k_logic<1, true, true, 0><<<lblocks, 128, 0, stream>>>(wa);
                        else k_logic<1, false, true, 0><<<lblocks, 128, 0, stream>>>(wa);
                    } else {
                        if (wantCounters) k_logic<1, true, false, 0><<<lblocks, 128, 0, stream>>>(wa);
                        else k_logic<1, false, false, 0><<<lblocks, 128, 0, stream>>>(wa);
                    }
                    if (hasMeasured) k_eval<<<tblocks, 128, 0, stream>>>(wa);
                    if (timeRays) CUDA_TRY(cudaEventRecord(rayEvent(), stream));
                    if (scene->d.nAnimated) {
                        if (wantCounters) k_trace<true, true><<<tblocks, 128, 0, stream>>>(wa);
                        else k_trace<false, true><<<tblocks, 128, 0, stream>>>(wa);
                    } else {
                        if (wantCounters) k_trace<true, false><<<tblocks, 128, 0, stream>>>(wa);
                        else k_trace<false, false><<<tblocks, 128, 0, stream>>>(wa);
                    }
                    if (timeRays) CUDA_TRY(cudaEventRecord(rayEvent(), stream));
                    launches += hasMeasured ? 3 : 2;
                    if ((iter & 7) == 7) {
                        CUDA_TRY(cudaMemcpyAsync(ln->hostCounters, ln->counters.p, C_COUNT * sizeof(unsigned),
                                                 cudaMemcpyDeviceToHost, stream));
                        CUDA_TRY(cudaStreamSynchronize(stream));
                        unsigned alive = 0;
                        for (int k = 0; k < K_COUNT; ++k) alive += ln->hostCounters[C_QN + setOut * K_COUNT + k];
                        if (traceIters)
                            fprintf(stderr, "[pbrtgpu] batch@%zu round %d iter %d: queued %u (eval %u/%u/%u hit %u shadow %u mis %u) next sample %u of %u\n",
                                    begin, r, iter, alive, ln->hostCounters[C_QN + setOut * K_COUNT + K_E_LIGHT],
                                    ln->hostCounters[C_QN + setOut * K_COUNT + K_E_MIS], ln->hostCounters[C_QN + setOut * K_COUNT + K_E_PATH],
                                    ln->hostCounters[C_QN + setOut * K_COUNT + K_HIT], ln->hostCounters[C_QN + setOut * K_COUNT + K_SHADOW],
                                    ln->hostCounters[C_QN + setOut * K_COUNT + K_MIS], ln->hostCounters[C_NEXT_SAMPLE], wa.totalSamples);
                        if (alive == 0) break;
                    }
                    if (iter > 4000000) return Fail(PBRTGPU_ECUDA, "wavefront did not terminate");
                }
                CUDA_TRY(cudaGetLastError());
            }
            if (stats) CUDA_TRY(cudaEventRecord(ln->ev[2], stream));
            if (wave2) {
                // ---- phase 2 as a vertex wavefront over the task chains (slot c = chain c), then the film in sample order
                WvArgs wa;
                wa.sc = scene->d; wa.rp = rp; wa.b = b; wa.pool = ln->pool.p; wa.P = (uint32_t)std::max(poolSize, maxBatch);
                wa.qRay = ln->traceQ.p; wa.qEval = ln->evalQ.p; wa.counters = ln->counters.p;
                wa.totalSamples = 0; wa.stats = scene->stats.p;
                wa.liveList = ln->liveList.p; wa.sortList = nullptr; wa.sortClass = nullptr;
                // the kernels shrink with the number of live chains: the grids follow the count read back every 8 iterations
                // (it only falls in between); an empty persistent kernel at full size costs 6 - 16 us, and a chain wavefront
                // runs as many iterations as its longest chain has vertices (profiles/r02d_launches.csv)
                size_t liveEst = nb;
                for (int iter = 0;; ++iter) {
                    wa.iter = iter;
                    const int cblocks = (int)((liveEst + 127) / 128);
                    const int tblocks = (int)std::min<size_t>((size_t)nSM * 8, std::max<size_t>(1, (liveEst * 3 + 255) / 256));
                    // empty the queues, the cursors and the live list this iteration writes (its length: WC_LIVE / WC_LIVE2)
                    CUDA_TRY(cudaMemsetAsync(ln->counters.p, 0, WC_LIVE * sizeof(unsigned), stream));
                    CUDA_TRY(cudaMemsetAsync(ln->counters.p + WvLiveCounterHost((iter + 1) & 1), 0, sizeof(unsigned), stream));
                    if (scene->d.nAnimated) {
                        if (wantCounters) k_wshade<2, true, true><<<cblocks, 128, 0, stream>>>(wa);
                        else k_wshade<2, false, true><<<cblocks, 128, 0, stream>>>(wa);
                    } else {
                        if (wantCounters) k_wshade<2, true, false><<<cblocks, 128, 0, stream>>>(wa);
                        else k_wshade<2, false, false><<<cblocks, 128, 0, stream>>>(wa);
                    }
                    if (hasMeasured) k_weval<<<tblocks, 128, 0, stream>>>(wa);
                    if (scene->d.nAnimated) {
                        if (wantCounters) k_wtrace<true, true><<<tblocks, 128, 0, stream>>>(wa);
                        else k_wtrace<false, true><<<tblocks, 128, 0, stream>>>(wa);
                    } else {
                        if (wantCounters) k_wtrace<true, false><<<tblocks, 128, 0, stream>>>(wa);
                        else k_wtrace<false, false><<<tblocks, 128, 0, stream>>>(wa);
                    }
                    launches += hasMeasured ? 3 : 2;
                    if ((iter & 7) == 7) {
                        CUDA_TRY(cudaMemcpyAsync(ln->hostCounters, ln->counters.p, WC_COUNT * sizeof(unsigned),
                                                 cudaMemcpyDeviceToHost, stream));
                        CUDA_TRY(cudaStreamSynchronize(stream));
                        liveEst = ln->hostCounters[WvLiveCounterHost((iter + 1) & 1)];
                        if (liveEst == 0) break;
                    }
                    if (iter > 40000000) return Fail(PBRTGPU_ECUDA, "chain wavefront did not terminate");
                }
                CUDA_TRY(cudaGetLastError());
                k_film<<<blocks2, 128, 0, stream>>>(pa);
                CUDA_TRY(cudaGetLastError());
                launches += 1;
            } else if (vertex2) {
                // ---- phase 2, vertex-synchronous: chains handed out longest first, then the film in sample order
                if (cfg.tailSort) {
                    size_t tmp = ln->sortTemp.n;
                    cub::DeviceRadixSort::SortPairsDescending(ln->sortTemp.p, tmp, ln->tailCount.p, ln->tailKeysOut.p,
                                                              ln->orderIn.p, ln->orderOut.p, (int)nb, 0, 14, stream);
                    va.b.order = ln->orderOut.p;
                    va.sortedCount = ln->tailKeysOut.p;
                }
                CUDA_TRY(cudaMemsetAsync(ln->workCounter.p, 0, sizeof(unsigned), stream));
                if (scene->d.nAnimated) {
                    if (wantCounters) k_vertex<2, true, true><<<vblocks, 128, 0, stream>>>(va);
                    else k_vertex<2, false, true><<<vblocks, 128, 0, stream>>>(va);
                } else {
                    if (wantCounters) k_vertex<2, true, false><<<vblocks, 128, 0, stream>>>(va);
                    else k_vertex<2, false, false><<<vblocks, 128, 0, stream>>>(va);
                }
                CUDA_TRY(cudaGetLastError());
                k_film<<<blocks2, 128, 0, stream>>>(pa);
                CUDA_TRY(cudaGetLastError());
                launches += 1;
            } else if (!direct) {   // thread-per-task tails: tasks with parked samples first
                k_tail_flags<<<(int)((nb + 255) / 256), 256, 0, stream>>>(ln->tailCount.p, ln->tailFlag.p, (int)nb);
                size_t tmp = ln->sortTemp.n;
                cub::DeviceRadixSort::SortPairsDescending(ln->sortTemp.p, tmp, ln->tailFlag.p, ln->tailKeysOut.p,
                                                          ln->orderIn.p, ln->orderOut.p, (int)nb, 0, 1, stream);
                pa.b.order = ln->orderOut.p;
            }
            if (vertex2 || wave2) {
            } else if (direct) {
                k_film_direct<<<blocks2, 128, 0, stream>>>(pa, r + 1 < rounds ? 1 : 0);
                CUDA_TRY(cudaGetLastError());
            } else if (!useWavefront2) {
                if (scene->d.nAnimated) {
                    if (wantCounters) k_path<2, true, true><<<blocks2, 128, 0, stream>>>(pa);
                    else k_path<2, false, true><<<blocks2, 128, 0, stream>>>(pa);
                } else {
                    if (wantCounters) k_path<2, true, false><<<blocks2, 128, 0, stream>>>(pa);
                    else k_path<2, false, false><<<blocks2, 128, 0, stream>>>(pa);
                }
                CUDA_TRY(cudaGetLastError());
            } else {
                // ---- phase 2 as a wavefront: one slot per task-pixel chain
                WfArgs wa;
                wa.sc = scene->d; wa.rp = rp; wa.b = pa.b; wa.pool = ln->pool.p; wa.P = (uint32_t)poolSize;
                wa.q = ln->traceQ.p; wa.counters = ln->counters.p;
                wa.totalSamples = 0; wa.stats = scene->stats.p;
                wa.film = film; wa.directFilm = pa.directFilm; wa.dlFrames = nullptr;
                CUDA_TRY(cudaMemsetAsync(ln->counters.p, 0, C_COUNT * sizeof(unsigned), stream));
                const int tblocks = nSM * 8;
                for (int iter = 0;; ++iter) {
                    wa.iter = iter;
                    const int setOut = (iter & 1) ^ 1;
                    CUDA_TRY(cudaMemsetAsync(ln->counters.p + C_QN + setOut * K_COUNT, 0, K_COUNT * sizeof(unsigned), stream));
                    CUDA_TRY(cudaMemsetAsync(ln->counters.p + C_TRACE_CUR, 0, 2 * sizeof(unsigned), stream));
                    if (scene->d.nAnimated) {
                        if (wantCounters) k_logic<2, true, true, 0><<<blocks2, 128, 0, stream>>>(wa);
                        else k_logic<2, false, true, 0><<<blocks2, 128, 0, stream>>>(wa);
                    } else {
                        if (wantCounters) k_logic<2, true, false, 0><<<blocks2, 128, 0, stream>>>(wa);
                        else k_logic<2, false, false, 0><<<blocks2, 128, 0, stream>>>(wa);
                    }
                    if (hasMeasured) k_eval<<<tblocks, 128, 0, stream>>>(wa);
                    if (scene->d.nAnimated) {
                        if (wantCounters) k_trace<true, true><<<tblocks, 128, 0, stream>>>(wa);
                        else k_trace<false, true><<<tblocks, 128, 0, stream>>>(wa);
                    } else {
                        if (wantCounters) k_trace<true, false><<<tblocks, 128, 0, stream>>>(wa);
                        else k_trace<false, false><<<tblocks, 128, 0, stream>>>(wa);
                    }
                    launches += hasMeasured ? 3 : 2;
                    if ((iter & 7) == 7) {
                        CUDA_TRY(cudaMemcpyAsync(ln->hostCounters, ln->counters.p, C_COUNT * sizeof(unsigned),
                                                 cudaMemcpyDeviceToHost, stream));
                        CUDA_TRY(cudaStreamSynchronize(stream));
                        unsigned alive = 0;
                        for (int k = 0; k < K_COUNT; ++k) alive += ln->hostCounters[C_QN + setOut * K_COUNT + k];
                        if (alive == 0) break;
                    }
                    if (iter > 40000000) return Fail(PBRTGPU_ECUDA, "tail wavefront did not terminate");
                }
                CUDA_TRY(cudaGetLastError());
            }
            launches += 2;
            if (stats) {
                CUDA_TRY(cudaEventRecord(ln->ev[3], stream));
                CUDA_TRY(cudaEventSynchronize(ln->ev[3]));
                float ms = 0.f;
                CUDA_TRY(cudaEventElapsedTime(&ms, ln->ev[0], ln->ev[1]));
                tableMs += ms;
                CUDA_TRY(cudaEventElapsedTime(&ms, ln->ev[1], ln->ev[2]));
                traceMs += ms;
                CUDA_TRY(cudaEventElapsedTime(&ms, ln->ev[2], ln->ev[3]));
                tailMs += ms;
                for (size_t k = 0; k + 1 < rayEventsUsed; k += 2) {
                    CUDA_TRY(cudaEventElapsedTime(&ms, rayEvents[k], rayEvents[k + 1]));
                    rayMs += ms;
                    ++rayLaunches;
                }
                rayEventsUsed = 0;
            }
        }
    }
    CUDA_TRY(cudaStreamSynchronize(stream));
    job.tableMs = tableMs; job.traceMs = traceMs; job.tailMs = tailMs; job.launches = launches;
    job.rayMs = rayMs; job.rayLaunches = rayLaunches;
    for (cudaEvent_t e : rayEvents) cudaEventDestroy(e);
    return PBRTGPU_OK;
}

}  // namespace

extern "C" {

int pbrtgpu_render_device(PbrtGpuScene *scene, const PbrtGpuRenderDesc *rd, void *d_pixels, void *streamv,
                          PbrtGpuStats *stats) {
    if (!scene || !rd || !d_pixels) return Fail(PBRTGPU_EINVAL, "null argument");
    if (!scene->valid) return Fail(PBRTGPU_EINVAL, "the scene handle is invalid: its last pbrtgpu_scene_update failed on the device");
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
    rp.shutterOpen = rd->shutterOpen; rp.shutterClose = rd->shutterClose;
    rp.filterXWidth = rd->filterXWidth; rp.filterYWidth = rd->filterYWidth;
    rp.xPixelStart = ext.xPixelStart; rp.yPixelStart = ext.yPixelStart;
    rp.xPixelCount = ext.xPixelCount; rp.yPixelCount = ext.yPixelCount;
    rp.spp = spp; rp.maxDepth = rd->maxDepth;
    rp.integrator = rd->integrator; rp.lightStrategy = rd->lightStrategy;
#if PBRT_FILTER
    {   // the 0.5 box keeps its footprint-bit path (FilmAdd); everything else goes through the weight table
        bool box = rd->filterXWidth == 0.5f && rd->filterYWidth == 0.5f;
        float table[256];
        for (int i = 0; i < 256; ++i) {
            table[i] = rd->hasFilterTable ? rd->filterTable[i] : 1.f;
            if (table[i] != 1.f) box = false;
        }
        rp.filterTable = nullptr;
        if (!box) {
            CUDA_TRY(scene->filterTable.Upload(table, 256));
            rp.filterTable = scene->filterTable.p;
        }
    }
#endif

    // per task: permutation tables (64 B/sample) + finished radiance (16 B/sample) + parked
    // bounce-3 vertices (64 B/sample, sparse) + scrambles + two MT19937 blocks
    TableArrays tables;
    BuildLayout(*rd, &rp.lay, &tables);
    const bool pathLayout = IsPathLayout(*rd);
    const bool directInteg = rd->integrator == PBRTGPU_INTEGRATOR_DIRECTLIGHTING;
    {
        const int rcs = ValidateRenderForScene(*rd, scene->d.nLights, scene->nSpecularMaterials);
        if (rcs != PBRTGPU_OK) return rcs;
    }
    const size_t perTask = (size_t)spp * ((size_t)rp.lay.rowStride * 2 + 16 + (directInteg ? 0 : 64) + (pathLayout ? 0 : tables.tgtPerSample)) +
                           SCR_STRIDE * 4 + 2 * MT_N * 4 + 4 + sizeof(TaskRec) + (size_t)std::max(1, spp / 32) * 4;
    const bool wantCounters = stats && scene->collectStats == 1;
    const bool timeRays = stats && scene->collectStats == 2;
    const RenderConfig cfg = ReadConfig();
    const size_t poolSlots = cfg.pool;
    // lanes: the kernels' register allocations fill the SMs, so lanes cannot overlap two full-size kernels; what they
    // hide is each other's low-occupancy stretches (the tail kernel, the wavefront's drain at the end of a batch).
    // Measured on C2 (bunny 1024x1024): 128 spp 1/2/3/4 lanes = 210/227/240/250 Msamples/s, 1024 spp 1/2/4 lanes =
    // 197/219/188 (more lanes = smaller batches = more, equally long, tail launches).  Default 2.
    int nLanes = cfg.lanes > 0 ? cfg.lanes : 2;
    nLanes = std::max(1, std::min(nLanes, (int)PbrtGpuScene::MAX_LANES));
    if (cfg.lanes <= 0)
        while (nLanes > 1 && tasks.size() / nLanes * (size_t)spp < poolSlots) --nLanes;
    while (nLanes > 1 && tasks.size() < (size_t)nLanes * 1024) --nLanes;   // too little work to split further
    if (timeRays) nLanes = 1;   // the timed kernel must have the device to itself
    size_t maxBatch = std::max<size_t>(1024, ScratchBudgetBytes() / (size_t)nLanes / perTask);
    maxBatch = std::min(maxBatch, poolSlots);   // the tail wavefront keeps one pool slot per task of the batch
    {   // a whole number of batches per lane, all of (nearly) the same size
        size_t nBatches = (std::max<size_t>(tasks.size(), 1) + maxBatch - 1) / maxBatch;
        nBatches = (nBatches + nLanes - 1) / nLanes * nLanes;
        maxBatch = (std::max<size_t>(tasks.size(), 1) + nBatches - 1) / nBatches;
    }
    CUDA_TRY(scene->stats.Ensure(1));
    if (wantCounters) CUDA_TRY(cudaMemsetAsync(scene->stats.p, 0, sizeof(DeviceStats), stream));
    CUDA_TRY(cudaEventRecord(scene->ev[2], stream));
    std::vector<LaneJob> jobs(nLanes);
    for (int l = 0; l < nLanes; ++l) {
        if (!scene->lanes[l]) scene->lanes[l] = AcquireLane(scene->device);
        Lane &ln = *scene->lanes[l];
        if (!ln.stream) CUDA_TRY(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
        for (int i = 0; i < 4; ++i)
            if (!ln.ev[i]) CUDA_TRY(cudaEventCreate(&ln.ev[i]));
        LaneJob &j = jobs[l];
        j.scene = scene; j.ln = &ln; j.tasks = &tasks; j.maxBatch = maxBatch; j.laneIndex = l; j.nLanes = nLanes;
        j.rp = rp; j.spp = spp; j.film = film; j.timing = stats != nullptr; j.wantCounters = wantCounters; j.timeRays = timeRays;
        j.tables = tables; j.pathLayout = pathLayout; j.cfg = cfg;
    }
    {
        auto run = [](LaneJob *j) {
            j->rc = RunLane(*j);
            if (j->rc != PBRTGPU_OK) j->err = g_lastError;   // thread-local: carry it back to the caller's thread
        };
        std::vector<std::thread> workers;
        for (int l = 1; l < nLanes; ++l) workers.emplace_back(run, &jobs[l]);
        run(&jobs[0]);
        for (std::thread &w : workers) w.join();
    }
    uint64_t launches = 0, rayLaunches = 0;
    float tableMs = 0.f, traceMs = 0.f, tailMs = 0.f, rayMs = 0.f;
    for (const LaneJob &j : jobs) {
        if (j.rc != PBRTGPU_OK) return Fail(j.rc, j.err);
        launches += j.launches;
        rayMs += j.rayMs; rayLaunches += j.rayLaunches;
        tableMs += j.tableMs; traceMs += j.traceMs; tailMs += j.tailMs;   // lanes overlap: the sum can exceed renderMs
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
        stats->rayKernelMs = rayMs;
        stats->rayKernelLaunches = rayLaunches;
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

// (comm.cu)
int PbrtGpuCommReduceToRoot(PbrtGpuComm *c, float *buf, size_t count, cudaStream_t stream);
int PbrtGpuCommRank(const PbrtGpuComm *c);
int PbrtGpuCommWorld(const PbrtGpuComm *c);

int pbrtgpu_render_sharded(PbrtGpuScene *scene, const PbrtGpuRenderDesc *rdIn, PbrtGpuComm *comm, PbrtGpuPixel *pixels,
                           PbrtGpuStats *stats) {
    if (!scene || !rdIn || !comm) return Fail(PBRTGPU_EINVAL, "null argument");
    const int rank = PbrtGpuCommRank(comm), world = PbrtGpuCommWorld(comm);
    PbrtGpuRenderDesc rd = *rdIn;
    rd.shardCount = world; rd.shardIndex = rank;
    if (rd.shardBlock <= 0) rd.shardBlock = std::max(1, 4 * rd.xResolution);
    int rc = ValidateRenderDesc(rd);
    if (rc != PBRTGPU_OK) return rc;
    const FilmExtent ext = ComputeExtent(rd);
    const size_t nFilm = (size_t)ext.xPixelCount * ext.yPixelCount;
    CUDA_TRY(scene->film.Ensure(nFilm));
    PbrtGpuStats local;
    rc = pbrtgpu_render_device(scene, &rd, scene->film.p, nullptr, stats ? stats : &local);
    if (rc != PBRTGPU_OK) return rc;
    // the exchange step: one reduce of the 16 B/pixel film to rank 0 (its time is part of the caller's e2e, not of renderMs)
    rc = PbrtGpuCommReduceToRoot(comm, reinterpret_cast<float *>(scene->film.p), nFilm * 4, nullptr);
    if (rc != PBRTGPU_OK) return rc;
    if (rank == 0 && pixels) CUDA_TRY(cudaMemcpy(pixels, scene->film.p, nFilm * sizeof(PbrtGpuPixel), cudaMemcpyDeviceToHost));
    else CUDA_TRY(cudaStreamSynchronize(nullptr));
    return PBRTGPU_OK;
}

int pbrtgpu_intersect(PbrtGpuScene *scene, const PbrtGpuRay *rays, int n, PbrtGpuHit *hits) {
    if (!scene || (!rays && n) || (!hits && n) || n < 0) return Fail(PBRTGPU_EINVAL, "bad argument");
    if (!scene->valid) return Fail(PBRTGPU_EINVAL, "the scene handle is invalid");
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
    if (!scene->valid) return Fail(PBRTGPU_EINVAL, "the scene handle is invalid");
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
    DevBuf<float4> rec;
    DevBuf<uint8_t> tgt;
    SampleLayoutDev L;
    TableArrays T;
    BuildLayout(*rd, &L, &T);
    const bool pathLayout = IsPathLayout(*rd);
    const int nf = FloatsPerSample(*rd);
    CUDA_TRY(rec.Ensure((size_t)spp * 4));
    CUDA_TRY(tgt.Ensure((size_t)spp * T.tgtPerSample + 16));
    CUDA_TRY(tasks.Upload(&tr, 1));
    CUDA_TRY(perm.Ensure((size_t)spp * L.rowStride));
    CUDA_TRY(scr.Ensure(SCR_STRIDE));
    CUDA_TRY(cur.Ensure(MT_N));
    CUDA_TRY(next.Ensure(MT_N));
    CUDA_TRY(mti.Ensure(1));
    CUDA_TRY(dout.Ensure((size_t)spp * nf));
    BatchArgs b;
    b.tasks = tasks.p; b.nBatch = 1; b.round = 0; b.perm = perm.p; b.scr = scr.p; b.rngCur = cur.p; b.rngNext = next.p;
    b.mti = mti.p; b.ls = nullptr; b.rec = rec.p; b.tailCount = nullptr; b.order = nullptr;
    if (pathLayout) CUDA_TRY(LaunchTablesAny(b, spp, PlanTables(spp), nullptr));
    else CUDA_TRY(LaunchTablesGeneralAny(b, spp, PlanTablesGeneral(spp, T.nArrays), T, tgt.p, nullptr));
    k_expand_samples<<<(spp + 127) / 128, 128>>>(b, spp, rd->shutterOpen, rd->shutterClose, L, nf, dout.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, dout.p, (size_t)spp * nf * sizeof(float), cudaMemcpyDeviceToHost));
    return PBRTGPU_OK;
}

}  // extern "C"
