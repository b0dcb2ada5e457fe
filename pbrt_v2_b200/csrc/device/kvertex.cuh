// kvertex.cuh -- the vertex-synchronous path kernels (included by pbrtgpu.cu).
//
// k_vertex<PHASE>: a persistent kernel whose warps advance 32 camera paths ONE PATH VERTEX per iteration, all lanes through
// the same stages in the same order (dvertex.cuh):
//     refill -> [RNG top-up] -> VertexA (shade, choose the 3 directions) -> the BSDF-evaluation site (<= 3 values)
//            -> VertexB -> shadow rays -> MIS rays -> radiance update -> path rays
// so a warp executes each stage once per iteration with every lane that has something to do in it, instead of the
// one-state-transition-per-iteration megakernel (k_path: 7 transitions per bounce, 3.3 of 32 lanes active in the tails,
// profiles/r01_k_path_ncu_full.txt).  Path state lives in registers from the camera ray to the sample's radiance: no slot
// round trips through HBM (the wavefront's 288 B per transition, profiles/r01_k_logic_ncu_full.txt).
//   PHASE 1  lanes = camera samples of the batch, handed out by a global counter; bounces 0..2 (table-driven, independent);
//            a sample that reaches its bounce-3 vertex is parked (rec[s], LS_TAIL, a bit in the task's tail mask)
//   PHASE 2  lanes = task chains: the parked samples of one task-pixel in sample order, each resumed with the task's RNG
//            exactly where the previous one stopped (7 + 3 + [1] draws per vertex); chains are handed out by a global
//            counter, longest first, so a lane whose chain ends takes the next one
// k_film: ImageFilm::AddSample for every sample of every task-pixel, in sample order (what k_path<2> did inline).
#pragma once
#include "dvertex.cuh"
#include "kcommon.cuh"
#include "kwarp.cuh"

#ifndef K_VERTEX_MIN_BLOCKS
#define K_VERTEX_MIN_BLOCKS 4
#endif

struct VxArgs {
    DScene sc;
    RenderParams rp;
    BatchArgs b;
    unsigned *workCounter;      // next camera sample (PHASE 1) / next chain (PHASE 2)
    unsigned totalSamples;      // PHASE 1: nBatch * spp
    const unsigned *sortedCount;// PHASE 2: tail counts in `order` (descending) or null (unsorted: b.tailCount[bi])
    DeviceStats *stats;
};

#if PBRT_TOP_LEVELS
// One TMA bulk copy (cp.async.bulk.shared::cluster.global, SASS UBLKCP) of the heap-ordered top of the scene BVH into the
// CTA's shared memory; completion is signalled on an mbarrier that every thread of the CTA then waits on.
DEV void StageTopNodes(float4 *dst, uint64_t *mbar, const float4 *src) {
    constexpr uint32_t bytes = PBRT_TOP_SLOTS * 32u;
    const uint32_t mb = (uint32_t)__cvta_generic_to_shared(mbar), ds = (uint32_t)__cvta_generic_to_shared(dst);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(ds), "l"(src), "r"(bytes), "r"(mb) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(mb) : "memory");
}
#endif

template <int PHASE, bool STATS, bool ANIM>
__global__ void __launch_bounds__(128, K_VERTEX_MIN_BLOCKS) k_vertex(const __grid_constant__ VxArgs a) {
    __shared__ uint32_t smt[PHASE == 2 ? 4 : 1][PHASE == 2 ? MT_N : 1];
    const DScene &sc = a.sc;
    const RenderParams &rp = a.rp;
    const unsigned FULL = 0xffffffffu;
#if PBRT_TOP_LEVELS
    __shared__ __align__(128) float4 sTop[2 * PBRT_TOP_SLOTS];
    __shared__ __align__(8) uint64_t sTopBar;
    StageTopNodes(sTop, &sTopBar, sc.topNodes);
    const float4 *smemTop = sTop;
#else
    const float4 *smemTop = nullptr;
#endif
#if PBRT_SMEM_STACK
    __shared__ uint32_t sStack[PBRT_SMEM_STACK * 128];
    uint32_t *smemStack = sStack + threadIdx.x;
#else
    uint32_t *smemStack = nullptr;
#endif
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int spp = rp.spp;
    const int nLights = (int)sc.nLights;

    // ---- per-lane path state
    bool have = false;        // the lane owns a camera path
    bool atVertex = false;    // ... whose current ray hit something: a vertex to shade (else: a ray to trace)
    bool exhausted = false;   // no more work to hand out
    Ray ray;
    ray.o = V3(0, 0, 0); ray.d = V3(0, 0, 1); ray.mint = 0.f; ray.maxt = 0.f; ray.time = 0.f;
    Hit hit;
    hit.prim = -1; hit.t = 0.f; hit.b1 = hit.b2 = 0.f; hit.maxtIn = 0.f;
    Spec T(1.f), L(0.f);
    int bounces = 0, s = 0, bi = 0, code = 0;
    bool afterSpecular = false;
    // PHASE 2: the chain
    bool chain = false;       // the lane owns a task chain
    int mti = 0, mw = 0;      // RNG position; index of the tail-mask word being walked
    uint32_t mbits = 0;       // its remaining bits
    TraceStats st;
    st.nodes = st.prims = 0;
    unsigned long long nClosest = 0, nShadow = 0, nShadedN = 0;

    for (;;) {
        // ================================================================== refill
        if (PHASE == 1) {
            if (!have && !exhausted) {
                for (;;) {
                    const unsigned g = AggregatedAddV(a.workCounter);
                    if (g >= a.totalSamples) { exhausted = true; break; }
                    bi = (int)(g / (unsigned)spp);
                    s = (int)(g - (unsigned)bi * (unsigned)spp);
                    const TaskRec tr = a.b.tasks[bi];
                    const int tw = tr.x1 - tr.x0;
                    if (a.b.round >= tw * (tr.y1 - tr.y0)) continue;   // (this task has no pixel in this round)
                    const int xPos = tr.x0 + a.b.round % tw, yPos = tr.y0 + a.b.round / tw;
                    const uint16_t *row = a.b.perm + ((size_t)bi * spp + s) * PERM_STRIDE;
                    code = CameraRay(rp, row, a.b.scr + (size_t)bi * SCR_STRIDE, xPos, yPos, &ray);
                    T = Spec(1.f); L = Spec(0.f);
                    bounces = 0; afterSpecular = false;
                    have = true; atVertex = false;
                    break;
                }
            }
        } else {
            if (!have && !exhausted) {
                for (;;) {
                    if (!chain) {
                        const unsigned c = AggregatedAddV(a.workCounter);
                        if (c >= (unsigned)a.b.nBatch) { exhausted = true; break; }
                        bi = a.b.order ? (int)a.b.order[c] : (int)c;
                        const unsigned nT = a.sortedCount ? a.sortedCount[c] : a.b.tailCount[bi];
                        if (nT == 0) {
                            if (a.sortedCount) { exhausted = true; break; }   // sorted descending: nothing but empty chains left
                            continue;
                        }
                        chain = true;
                        mti = a.b.mti[bi];
                        mw = 0;
                        mbits = __ldcg(a.b.tailMask + (size_t)bi * a.b.maskWords);
                    }
                    while (mbits == 0 && mw + 1 < a.b.maskWords) { ++mw; mbits = __ldcg(a.b.tailMask + (size_t)bi * a.b.maskWords + mw); }
                    if (mbits == 0) {   // the chain is finished: the task's stream position for its next pixel
                        a.b.mti[bi] = mti;
                        chain = false;
                        continue;
                    }
                    s = mw * 32 + (__ffs(mbits) - 1);
                    mbits &= mbits - 1;
                    const float4 v = __ldcs(a.b.ls + (size_t)bi * spp + s);
                    const float4 *r = a.b.rec + ((size_t)bi * spp + s) * 4;
                    const float4 r0 = __ldcs(r), r1 = __ldcs(r + 1), r2 = __ldcs(r + 2), r3 = __ldcs(r + 3);
                    ray.o = V3(r0.x, r0.y, r0.z); ray.mint = r0.w;
                    ray.d = V3(r1.x, r1.y, r1.z); ray.maxt = r1.w;
                    hit.t = r1.w;
                    T = Spec(r2.x, r2.y, r2.z);
                    hit.prim = __float_as_int(r2.w);
                    hit.b1 = r3.x; hit.b2 = r3.y; hit.maxtIn = r3.z;
                    ray.time = r3.w;
                    L = Spec(v.x, v.y, v.z);
                    code = __float_as_int(v.w);
                    afterSpecular = (code & LS_SPEC) != 0;
                    code &= ~(LS_SPEC | LS_TAIL);
                    bounces = 3;
                    have = true; atVertex = true;
                    break;
                }
            }
        }
        if (!__any_sync(FULL, have)) break;

        // ================================================================== PHASE 2: cooperative RNG top-up
        if (PHASE == 2) {   // the <= 11 draws of a vertex must be addressable in the two-block window
            unsigned m = __ballot_sync(FULL, have && atVertex && mti >= MT_N);
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

        // ================================================================== the vertex
        bool doVertex = have && atVertex;
        if (PHASE == 1 && doVertex && bounces == 3) {
            // ---- park the sample: from here on the path draws from the task's RNG stream
            float4 *r = a.b.rec + ((size_t)bi * spp + s) * 4;
            r[0] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint);
            r[1] = make_float4(ray.d.x, ray.d.y, ray.d.z, hit.t);
            r[2] = make_float4(T.r, T.g, T.b, __int_as_float(hit.prim));
            r[3] = make_float4(hit.b1, hit.b2, hit.maxtIn, ray.time);
            a.b.ls[(size_t)bi * spp + s] = make_float4(L.r, L.g, L.b, __int_as_float(code | LS_TAIL | (afterSpecular ? LS_SPEC : 0)));
            atomicAdd(a.b.tailCount + bi, 1u);
            atomicOr(a.b.tailMask + (size_t)bi * a.b.maskWords + (s >> 5), 1u << (s & 31));
            have = false; doVertex = false;
        }
        Vertex v;
        Spec f0(0.f), f1(0.f), f2(0.f);
        bool cont = false;
        int k0 = 0;
        if (doVertex) {
            VertexRandoms u;
            if (PHASE == 1) {
                const uint16_t *row = a.b.perm + ((size_t)bi * spp + s) * PERM_STRIDE;
                VertexRandomsFromTables(row, a.b.scr + (size_t)bi * SCR_STRIDE, bounces, &u);
            } else {
                const uint32_t *rc = a.b.rngCur + (size_t)bi * MT_N, *rn = a.b.rngNext + (size_t)bi * MT_N;
                auto rngF = [&](int k) -> float {
                    const int j = mti + k;
                    return UIntToFloat(Temper(j < MT_N ? __ldcg(rc + j) : __ldcg(rn + (j - MT_N))));
                };
                if (nLights > 0) {   // integrator.cpp:95,107-108: lightNum, LightSample(rng), BSDFSample(rng)
                    u.lnU = rngF(0); u.lsU0 = rngF(1); u.lsU1 = rngF(2);
                    u.bsU0 = rngF(4); u.bsU1 = rngF(5); u.bsUc = rngF(6);
                    k0 = 7;
                } else u.lnU = u.lsU0 = u.lsU1 = u.bsU0 = u.bsU1 = u.bsUc = 0.f;
                u.pU0 = rngF(k0); u.pU1 = rngF(k0 + 1); u.pUc = rngF(k0 + 2);   // path.cpp:90
            }
            ShadeCtx ctx;
            VertexA<ANIM>(sc, ray, hit, bounces, afterSpecular, u, T, L, ctx, &v, STATS ? &nShadedN : nullptr);
        }
        // ---- the BSDF-evaluation site
        if (doVertex && v.need[VX_LIGHT]) f0 = VertexEval(sc, v.material, v.evalWo, v.evalWi[VX_LIGHT], v.evalFlags[VX_LIGHT]);
        if (doVertex && v.need[VX_MIS]) f1 = VertexEval(sc, v.material, v.evalWo, v.evalWi[VX_MIS], v.evalFlags[VX_MIS]);
#if PBRT_SPECULAR
        if (doVertex && v.need[VX_PATH] && !v.specSampled)
#else
        if (doVertex && v.need[VX_PATH])
#endif
            f2 = VertexEval(sc, v.material, v.evalWo, v.evalWi[VX_PATH], v.evalFlags[VX_PATH]);
        if (doVertex) {
            float rrU = 0.f;
            if (PHASE == 2 && bounces > 3) {
                const int j = mti + k0 + 3;
                rrU = UIntToFloat(Temper(j < MT_N ? __ldcg(a.b.rngCur + (size_t)bi * MT_N + j)
                                                  : __ldcg(a.b.rngNext + (size_t)bi * MT_N + (j - MT_N))));
            }
            bool rrDrawn = false;
            const Spec f[3] = {f0, f1, f2};
            cont = VertexB(f, bounces, rp.maxDepth, T, rrU, &v, &rrDrawn);
            if (PHASE == 2) mti += k0 + 3 + (rrDrawn ? 1 : 0);
        }

        // ================================================================== the rays, kind by kind
        Spec Ld(0.f);
#pragma unroll 1
        for (int r = 0; r < 3; ++r) {
            if (r == 2 && doVertex) {   // both direct-lighting rays are back: path.cpp:71-80, then the new throughput
                L = L + T * ((float)nLights * Ld);
                T = v.Tnew;
                if (cont) {
#if PBRT_SPECULAR
                    afterSpecular = v.specSampled;
#endif
                    ++bounces;
                    ray.o = v.p; ray.d = v.wiW[VX_PATH]; ray.mint = v.eps; ray.maxt = INF_F;
                }
            }
            const bool want = r == 2 ? (have && (!doVertex || cont)) : (doVertex && (r == 0 ? v.ray[VX_LIGHT] : v.ray[VX_MIS]));
            if (!__any_sync(FULL, want)) continue;
            if (want) {
                Ray q = ray;   // r == 2: the path ray (camera ray or continuation)
                if (r != 2) {
                    q.o = v.p; q.mint = v.eps;
                    q.d = r == 0 ? v.wiW[VX_LIGHT] : v.wiW[VX_MIS];
                    q.maxt = r == 0 ? v.visMaxt : INF_F;
                }
                Hit h;
                if (STATS) { if (r == 0) ++nShadow; else ++nClosest; }
                const bool found = Traverse<STATS, ANIM>(sc, q, r == 0, &h, &st, smemTop, smemStack);   // the single traversal site
                if (r == 0) { if (!found) Ld = Ld + v.pendS; }   // visibility.Unoccluded(scene), Transmittance == 1
#if PBRT_INFINITE
                else if (r == 1) { if (MisRayContributes<ANIM>(sc, q, h, found, v.lightNum)) Ld = Ld + v.pendM; }
                else {
                    atVertex = found; hit = h; ray.maxt = q.maxt;
                    if (!found) EscapedRadiance(sc, ray.d, !doVertex, afterSpecular, T, L);   // (a camera ray has no vertex behind it)
                }
#else
                else if (r == 1) { if (found && MisHitIsLight<ANIM>(sc, q, h, v.lightNum)) Ld = Ld + v.pendM; }
                else { atVertex = found; hit = h; ray.maxt = q.maxt; }
#endif
            }
        }

        // ================================================================== end of a sample
        if (have && ((doVertex && !cont) || !atVertex)) {
            // PathIntegrator::Li returned (no continuation, or the ray left the scene: path.cpp:113-118 adds nothing for
            // the in-scope lights); samplerrenderer.cpp:119-133 sanity checks
            const Spec Ls = SanitizeRadiance(L);
            a.b.ls[(size_t)bi * spp + s] = make_float4(Ls.r, Ls.g, Ls.b, __int_as_float(code));
            have = false;
        }
    }

    if (STATS) {
        unsigned long long vv[5] = {nClosest, nShadow, st.nodes, st.prims, nShadedN};
        for (int k = 0; k < 5; ++k) {
            unsigned long long x = vv[k];
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(FULL, x, o);
            vv[k] = x;
        }
        if (lane == 0) {
            atomicAdd(&a.stats->closestRays, vv[0]);
            atomicAdd(&a.stats->shadowRays, vv[1]);
            atomicAdd(&a.stats->nodes, vv[2]);
            atomicAdd(&a.stats->prims, vv[3]);
            atomicAdd(&a.stats->shadedN, vv[4]);
        }
    }
}

// ImageFilm::AddSample for the samples of every task's current pixel, in sample order (samplerrenderer.cpp:147-157):
// one thread per task-pixel.  `direct`: multi-pixel tasks add every sample straight to the film (a neighbour pixel of the
// same task may already have splatted into it: the reference's order is splat first, then this pixel's samples).
__global__ void __launch_bounds__(128) k_film(const __grid_constant__ PathArgs a) {
    const int bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= a.b.nBatch) return;
    const RenderParams &rp = a.rp;
    const TaskRec tr = a.b.tasks[bi];
    const int tw = tr.x1 - tr.x0;
    if (a.b.round >= tw * (tr.y1 - tr.y0)) return;
    const int xPos = tr.x0 + a.b.round % tw, yPos = tr.y0 + a.b.round / tw;
    const int spp = rp.spp;
    const float4 *ls = a.b.ls + (size_t)bi * spp;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bool direct = a.directFilm != 0;
    for (int s = 0; s < spp; ++s) {
        const float4 v = __ldcs(ls + s);
        FILM_ADD(a.b.perm + ((size_t)bi * spp + s) * PERM_STRIDE, a.b.scr + (size_t)bi * SCR_STRIDE,
                 __float_as_int(v.w) & (LS_LEFT | LS_RIGHT | LS_UP | LS_DOWN), Spec(v.x, v.y, v.z));
    }
    if (!direct) {
        PbrtGpuPixel *px = a.film + (size_t)(yPos - rp.yPixelStart) * rp.xPixelCount + (xPos - rp.xPixelStart);
        atomicAdd(&px->Lxyz[0], acc[0]);
        atomicAdd(&px->Lxyz[1], acc[1]);
        atomicAdd(&px->Lxyz[2], acc[2]);
        atomicAdd(&px->weightSum, acc[3]);
    }
}

// 0/1 keys for the "tasks with parked samples first" partition of the thread-per-task tail kernel (k_path<2>)
__global__ void k_tail_flags(const unsigned *count, unsigned *flag, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = count[i] ? 1u : 0u;
}
