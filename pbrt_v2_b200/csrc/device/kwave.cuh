// kwave.cuh -- phase 1 of PathIntegrator as a VERTEX wavefront (included by pbrtgpu.cu; the default since round 2).
//
// The first wavefront (k_logic / k_eval / k_trace, still what DirectLightingIntegrator runs on) cuts PathIntegrator::Li at
// every blocking request: a path vertex costs up to three BSDF evaluations and three rays, one after the other, and each of
// those six hand-overs moves the whole 288-byte slot through HBM (profiles/r01_k_logic_ncu_full.txt: 2 KB of state traffic
// per camera sample, more than the scene bytes the sample touches).  The vertex-synchronous megakernel (kvertex.cuh) showed
// that the three directions of a vertex -- light sample, MIS sample, continuation -- depend on the shading context and the
// vertex's random numbers only (dvertex.cuh), so they can be chosen together; it lost to the wavefront because its lanes
// idle through each other's stages (8 of 32 active).  This file puts that schedule on the wavefront's compaction:
//
//   k_wshade   one thread per path slot, ONE pass per path vertex: finish the previous vertex with its three ray results
//              (VertexB: L += T * nLights * Ld, T = Tnew), then -- path ray missed / path over: write the sample's radiance
//              and take the next camera sample of the batch; bounce 3 reached: park the sample for phase 2; else VertexA:
//              shade, choose the three directions, write the vertex record.  BSDF values of analytic materials are computed
//              on the spot and their rays queued; measured-BRDF directions go to the evaluation queue
//   k_weval    persistent lanes over the evaluation queue: IrregIsotropicBRDF::f per (slot, direction); a direction whose
//              value is not black gets its ray queued
//   k_wtrace   persistent lanes over the three ray queues (shadow, MIS, path rays in separate runs): BVHAccel::IntersectP /
//              Intersect step-wise as in k_trace; the MIS ray's "did it land on the sampled light" test happens here
//   k_wcompact class bytes (written by whoever learns a slot's fate) -> slot lists sorted by next action: k_wshade's warps are
//              all shading a vertex or all finishing a sample, and neighbouring threads still read neighbouring slots
// so a vertex moves its state through HBM once (~420 B in k_wshade, 48 B per evaluation, 48 B per ray) instead of six times.
// k_wshade<2> runs phase 2 (the task chains) on the same kernels (PBRTGPU_WAVE=3; not the default: see pbrtgpu.cu RenderConfig).
// Every float expression is dvertex.cuh's (shared with k_vertex and the host-compiled harness): same operands, same
// operation order per quantity, so the film is the same film.
#pragma once
#include "dvertex.cuh"
#include "kcommon.cuh"
#include "kwarp.cuh"
#include "drng.cuh"

// resident blocks per SM the register allocator must allow.  k_wshade: 4 (128 registers) / 5 (96) / 6 (80, 550 B of spills):
// 6 measured +6.8 %, +0.5 %, +4.2 % on C2 in three visits (profiles/r02c_ab_wave2.json, r02d_ab_wave3.json, r02e_ab_wave4.json)
#ifndef K_WSHADE_MIN_BLOCKS
#define K_WSHADE_MIN_BLOCKS 6
#endif
#ifndef K_WEVAL_MIN_BLOCKS
#define K_WEVAL_MIN_BLOCKS 8
#endif

// slot fields (SoA float4: field f of slot i is pool[f * P + i])
enum {
    W_P = 0,   // origin of the vertex's rays: p.xyz, rayEpsilon            (camera ray: o.xyz, 0)
    W_D0,      // light sample:  wiW.xyz, visMaxt            -> .w = 1 if the shadow ray was blocked
    W_D1,      // MIS sample:    wiW.xyz, lightNum (int)     -> .w = 1 if the ray landed on that light, facing it
    W_D2,      // continuation / camera ray: d.xyz, -        -> .w = t of the hit
    W_HIT,     // path ray result: prim (-1: none), b1, b2, maxtIn
    W_EWO,     // evalWo.xyz, material
    W_EW0,     // evalWi[k].xyz, evalFlags[k] | (ray allowed << 16)
    W_EW1,
    W_EW2,
    W_F0,      // BSDF values f[k].rgb
    W_F1,
    W_F2,
    W_T,       // T.rgb (before the vertex), packed bits (WB_*)
    W_L,       // L.rgb, film-footprint code
    W_ID,      // task (batch index), sample, ray.time, pdfP
    W_V0,      // Li.rgb, scalarL
    W_V1,      // absM, weightM, pdfM, absP
    W_V2,      // specF.rgb (PBRT_SPECULAR), -
    W_CH,      // phase 2, the slot's task chain: mti, tail-mask word index, its remaining bits, the vertex's roulette draw
    W_COUNT,
    // scenes with TransformedPrimitives: AnimatedTransform::Interpolate(ray.time) of the first ANIM_CACHE_MAX instances, filled
    // when the slot takes a camera sample (ray.time is constant along the path): 8 float4 each (dscene.cuh RayW2P)
    W_AX = W_COUNT,
    W_COUNT_ANIM = W_COUNT + 8 * ANIM_CACHE_MAX
};
// W_T.w: flags, bounces (8 bits), lightNum (16 bits)
enum { WB_VERTEX = 0x1, WB_AFTER_SPEC = 0x2, WB_SPEC_SAMPLED = 0x4, WB_NEED0 = 0x8, WB_NEED1 = 0x10, WB_NEED2 = 0x20,
       WB_RETIRED = 0x40, WB_BOUNCES_SHIFT = 8, WB_LIGHT_SHIFT = 16 };
// counters
enum { WC_RAYS = 0 /* 3 */, WC_EVALS = 3, WC_TRACE_CUR, WC_EVAL_CUR,
       WC_LIVE,          // slots that hold a camera path after k_wshade (0: the batch is done)
       WC_NEXT_SAMPLE,   // (everything before it is reset every iteration)
       WC_LIVE2,         // phase 2: WC_LIVE and WC_LIVE2 are the lengths of the two live-chain lists (ping-pong, see k_wshade<2>)
       WC_SORT,          // phase 1: lengths of the two sorted slot lists (see WvArgs::sortList)
       WC_COUNT = WC_SORT + 2 };
#define W_RAY_ALLOWED 0x10000
#define W_RR 0x20000   // (direction 2, phase 2) the path ray also has to survive Russian roulette: see k_weval

struct WvArgs {
    DScene sc;
    RenderParams rp;
    BatchArgs b;
    float4 *pool;
    uint32_t P;
    uint32_t *qRay;      // [3][P] slots with a shadow / MIS / path ray to trace this iteration
    uint32_t *qEval;     // [3 * P] (slot << 2) | direction
    unsigned *counters;  // WC_*
    unsigned totalSamples;
    int iter;            // 0: every slot starts with a camera sample
    DeviceStats *stats;
    uint32_t *sortList;  // phase 1: [2 kinds][P] or null
    uint8_t *sortClass;  //          [P rounded up to 1024] 0: none, 1 + SORT_* : the slot's class of this iteration
    uint32_t *liveList;  // phase 2: [2][nBatch] the chains still running, compacted: iteration i reads list i & 1 (from
                         // iteration 1 on) and writes list (i + 1) & 1, so the kernels shrink with the number of live chains
};
DEV int WvLiveCounter(int list) { return list ? WC_LIVE2 : WC_LIVE; }
// Phase 1, slots sorted by what k_wshade will do with them next (null: thread i serves slot i).  Every live slot gets
// exactly one class during an iteration -- from whoever learns its fate: k_wtrace when its path ray comes back (hit, and
// not the bounce-3 vertex that is only parked: SORT_VERTEX; else SORT_OTHER), k_wshade / k_weval when they decide that no
// path ray will be traced (SORT_OTHER: the sample ends at the next pass) -- k_wcompact turns the class bytes into two slot
// lists, and k_wshade of the next iteration serves list SORT_VERTEX, then SORT_OTHER: its warps are all shading a vertex,
// or all finishing a sample and starting a camera ray, instead of a mix that ran with 7 of 32 lanes active
// (profiles/r02c_k_wshade_ncu_full.txt).  The lists keep the slots of every 1024-slot block in ascending order, so
// neighbouring threads still read neighbouring slots (lists appended to in completion order scattered every slot access:
// k_wtrace 228 -> 284 ms per step, profiles/r02f_ab_wsort.json).
enum { SORT_VERTEX = 0, SORT_OTHER = 1 };
DEV void WvSortPush(const WvArgs &a, int kind, uint32_t slot);

// `kind` must be the same for all converged lanes (a compile-time constant at the call site)
DEV void WvPushRay(const WvArgs &a, int kind, uint32_t slot) {
    const unsigned i = AggregatedAddV(a.counters + WC_RAYS + kind);
    a.qRay[(size_t)kind * a.P + i] = slot;
}
// ... any kind per lane: one atomic per (warp, kind)
DEV void WvPushRayAny(const WvArgs &a, int kind, uint32_t slot) {
    const unsigned peers = __match_any_sync(__activemask(), kind);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(a.counters + WC_RAYS + kind, (unsigned)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    a.qRay[(size_t)kind * a.P + base + __popc(peers & ((1u << lane) - 1u))] = slot;
}

DEV void WvSortPush(const WvArgs &a, int kind, uint32_t slot) {
    if (a.sortList) a.sortClass[slot] = (uint8_t)(1 + kind);
}

// Class bytes -> slot lists.  One block per 1024 consecutive slots (a thread reads its four class bytes as one word), a
// block-wide exclusive scan of the per-thread counts, one atomicAdd per block and list for the block's range, the slots
// written in ascending order inside it; the class bytes are cleared for the next iteration.
#define WCOMPACT_THREADS 256
__global__ void __launch_bounds__(WCOMPACT_THREADS) k_wcompact(const __grid_constant__ WvArgs a) {
#ifdef __CUDACC__
    __shared__ unsigned warpV[WCOMPACT_THREADS / 32], warpO[WCOMPACT_THREADS / 32], baseV, baseO;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t first = (blockIdx.x * WCOMPACT_THREADS + threadIdx.x) * 4u;
    uint32_t *word = reinterpret_cast<uint32_t *>(a.sortClass) + (first >> 2);
    const uint32_t w = first < a.P ? *word : 0u;
    if (w) *word = 0u;
    unsigned cV = 0, cO = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t c = (w >> (8 * k)) & 0xffu;
        cV += c == 1u + SORT_VERTEX; cO += c == 1u + SORT_OTHER;
    }
    unsigned sV = cV, sO = cO;   // inclusive warp scans
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(FULL, sV, o), q = __shfl_up_sync(FULL, sO, o);
        if (lane >= o) { sV += v; sO += q; }
    }
    if (lane == 31) { warpV[warp] = sV; warpO[warp] = sO; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tV = 0, tO = 0;
        for (int i = 0; i < WCOMPACT_THREADS / 32; ++i) { const unsigned v = warpV[i], q = warpO[i]; warpV[i] = tV; warpO[i] = tO; tV += v; tO += q; }
        baseV = tV ? atomicAdd(a.counters + WC_SORT + SORT_VERTEX, tV) : 0u;
        baseO = tO ? atomicAdd(a.counters + WC_SORT + SORT_OTHER, tO) : 0u;
    }
    __syncthreads();
    unsigned oV = baseV + warpV[warp] + sV - cV, oO = baseO + warpO[warp] + sO - cO;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t c = (w >> (8 * k)) & 0xffu;
        if (c == 1u + SORT_VERTEX) a.sortList[(size_t)SORT_VERTEX * a.P + oV++] = first + k;
        else if (c == 1u + SORT_OTHER) a.sortList[(size_t)SORT_OTHER * a.P + oO++] = first + k;
    }
#else   // (host-compiled harness: one thread does it all)
    if (blockIdx.x || threadIdx.x) return;
    for (uint32_t slot = 0; slot < a.P; ++slot) {
        const uint8_t c = a.sortClass[slot];
        a.sortClass[slot] = 0;
        if (c) a.sortList[(size_t)(c - 1) * a.P + a.counters[WC_SORT + (c - 1)]++] = slot;
    }
#endif
}

// ----------------------------------------------------------------------------------------------------------- k_wshade
// PHASE 1: slots are camera samples (bounces 0..2, random numbers from the tables), handed out by a global counter.
// PHASE 2: slot c is task chain c -- the parked samples of one task-pixel in sample order, each resumed with the task's RNG
//          exactly where the previous one stopped (7 + 3 + [1] draws per vertex, path.cpp:86-108) -- so the chains advance
//          one vertex per iteration side by side and the iteration count is the longest chain's vertex count; what
//          k_vertex<2> did with one lane per chain and 5 of 32 lanes active (profiles/r02a_k_vertex_ncu_full.txt).
template <int PHASE, bool STATS, bool ANIM>
__global__ void __launch_bounds__(128, K_WSHADE_MIN_BLOCKS) k_wshade(const __grid_constant__ WvArgs a) {
    __shared__ uint32_t smt[PHASE == 2 ? 4 : 1][PHASE == 2 ? MT_N : 1];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t slot = tid;
    bool live;
    if (PHASE == 1) {
        live = tid < a.P;
        if (a.sortList && a.iter > 0) {   // the sorted lists k_wcompact left: vertices to shade first, then the rest
            const unsigned nV = a.counters[WC_SORT + SORT_VERTEX], nO = a.counters[WC_SORT + SORT_OTHER];
            live = tid < nV + nO;
            if (live) slot = tid < nV ? a.sortList[(size_t)SORT_VERTEX * a.P + tid] : a.sortList[(size_t)SORT_OTHER * a.P + (tid - nV)];
        }
    } else if (a.iter == 0) live = tid < (uint32_t)a.b.nBatch;
    else {   // the chains that were still running after the previous iteration
        live = tid < a.counters[WvLiveCounter(a.iter & 1)];
        slot = live ? a.liveList[(size_t)(a.iter & 1) * a.b.nBatch + tid] : 0u;
    }
    const DScene &sc = a.sc;
    const RenderParams &rp = a.rp;
    const int spp = rp.spp;
    const size_t P = a.P;
    float4 *pool = a.pool;
    const int nLights = (int)sc.nLights;
    const int k0 = nLights > 0 ? 7 : 0;   // PHASE 2: draws of UniformSampleOneLight before the path's BSDFSample
#define WF(f) pool[(size_t)(f) * P + slot]
    if (PHASE == 1 && !live) return;

    Ray ray;
    ray.o = V3(0, 0, 0); ray.d = V3(0, 0, 1); ray.mint = 0.f; ray.maxt = 0.f; ray.time = 0.f;
    if (ANIM) { ray.xfCache = pool + (size_t)W_AX * P + slot; ray.xfStride = a.P; }
    Hit hit;
    hit.prim = -1; hit.t = 0.f; hit.b1 = hit.b2 = 0.f; hit.maxtIn = 0.f;
    Spec T(1.f), L(0.f);
    int bounces = 0, s = 0, bi = 0, code = 0;
    bool afterSpecular = false;
    bool have = false;       // the slot holds a camera path whose path ray has been traced
    bool atVertex = false;   // ... and hit something
    int mti = 0, mw = 0;     // PHASE 2: the chain's RNG position, the tail-mask word being walked
    uint32_t mbits = 0;      //          ... and its remaining bits
    unsigned long long nShadedN = 0;

    if (a.iter == 0) {
        if (PHASE == 2 && live) {
            bi = a.b.order ? (int)a.b.order[slot] : (int)slot;
            const TaskRec tr = a.b.tasks[bi];
            if (a.b.round >= (tr.x1 - tr.x0) * (tr.y1 - tr.y0) || a.b.tailCount[bi] == 0) {
                WF(W_T) = make_float4(0.f, 0.f, 0.f, __uint_as_float(WB_RETIRED));
                live = false;
            } else {
                mti = a.b.mti[bi];
                mbits = __ldcg(a.b.tailMask + (size_t)bi * a.b.maskWords);
            }
        }
    } else if (live) {
        const float4 fT = WF(W_T);
        const uint32_t bits = __float_as_uint(fT.w);
        if (bits & WB_RETIRED) {
            if (PHASE == 1) return;
            live = false;
        } else {
        const float4 fL = WF(W_L), fId = WF(W_ID), fP = WF(W_P), fD2 = WF(W_D2), fH = WF(W_HIT);
        T = Spec(fT.x, fT.y, fT.z);
        L = Spec(fL.x, fL.y, fL.z); code = __float_as_int(fL.w);
        bi = __float_as_int(fId.x); s = __float_as_int(fId.y); ray.time = fId.z;
        bounces = (int)((bits >> WB_BOUNCES_SHIFT) & 0xffu);
        afterSpecular = (bits & WB_AFTER_SPEC) != 0;
        float rrU = 0.f;
        if (PHASE == 2) {
            const float4 ch = WF(W_CH);
            mti = __float_as_int(ch.x); mw = __float_as_int(ch.y); mbits = __float_as_uint(ch.z); rrU = ch.w;
        }
        bool pathTraced = true;
        if (bits & WB_VERTEX) {
            // ---- the second half of the previous vertex (dvertex.cuh VertexB) with the BSDF values and the ray results
            const float4 f0 = WF(W_F0), f1 = WF(W_F1), f2 = WF(W_F2), v0 = WF(W_V0), v1 = WF(W_V1);
            Vertex v;
            v.need[0] = (bits & WB_NEED0) != 0; v.need[1] = (bits & WB_NEED1) != 0; v.need[2] = (bits & WB_NEED2) != 0;
            v.Li = Spec(v0.x, v0.y, v0.z); v.scalarL = v0.w;
            v.absM = v1.x; v.weightM = v1.y; v.pdfM = v1.z; v.absP = v1.w; v.pdfP = fId.w;
            const int lightNum = (int)(bits >> WB_LIGHT_SHIFT);
            v.Lemit = Spec(0.f);
            if (v.need[1]) {
                const PbrtGpuLight &light = sc.lights[lightNum];
                v.Lemit = Spec(light.L[0], light.L[1], light.L[2]);
#if PBRT_INFINITE
                if (light.type == PBRTGPU_LIGHT_INFINITE) { const float4 d1 = WF(W_D1); v.Lemit = LightLe(sc, light, V3(d1.x, d1.y, d1.z)); }
#endif
            }
#if PBRT_SPECULAR
            v.specSampled = (bits & WB_SPEC_SAMPLED) != 0;
            v.specF = Spec(0.f);
            if (v.specSampled) { const float4 v2 = WF(W_V2); v.specF = Spec(v2.x, v2.y, v2.z); }
#endif
            const Spec f[3] = {Spec(f0.x, f0.y, f0.z), Spec(f1.x, f1.y, f1.z), Spec(f2.x, f2.y, f2.z)};
            bool rrDrawn = false;
            const bool cont = VertexB(f, bounces, rp.maxDepth, T, rrU, &v, &rrDrawn);   // (PHASE 1, bounces <= 2: no roulette)
            if (PHASE == 2) mti += k0 + 3 + (rrDrawn ? 1 : 0);
            Spec Ld(0.f);
            if (v.ray[VX_LIGHT] && WF(W_D0).w == 0.f) Ld = Ld + v.pendS;    // visibility.Unoccluded(scene)
            if (v.ray[VX_MIS] && WF(W_D1).w == 1.f) Ld = Ld + v.pendM;      // the MIS ray landed on the light
            L = L + T * ((float)nLights * Ld);                              // path.cpp:71-80
            T = v.Tnew;
            if (cont) {
#if PBRT_SPECULAR
                afterSpecular = v.specSampled;
#endif
                ++bounces;
            } else
                pathTraced = false;   // PathIntegrator::Li returned at this vertex
        }
        have = true;
        if (pathTraced) {
            ray.o = V3(fP.x, fP.y, fP.z); ray.mint = fP.w;
            ray.d = V3(fD2.x, fD2.y, fD2.z); ray.maxt = fD2.w;
            hit.prim = __float_as_int(fH.x); hit.b1 = fH.y; hit.b2 = fH.z; hit.maxtIn = fH.w; hit.t = fD2.w;
            atVertex = hit.prim >= 0;
        }
        if (!atVertex) {
            // the path is over: no continuation, or the ray left the scene (which adds nothing for point / area lights:
            // path.cpp:113-118, samplerrenderer.cpp:238-242; infinite lights: EscapedRadiance); samplerrenderer.cpp:119-133 checks
#if PBRT_INFINITE
            if (pathTraced) EscapedRadiance(sc, ray.d, !(bits & WB_VERTEX), afterSpecular, T, L);
#endif
            const Spec Ls = SanitizeRadiance(L);
            a.b.ls[(size_t)bi * spp + s] = make_float4(Ls.r, Ls.g, Ls.b, __int_as_float(code));
            have = false;
        } else if (PHASE == 1 && bounces == 3) {
            // ---- park the sample: from here on the path draws from the task's RNG stream (phase 2)
            float4 *r = a.b.rec + ((size_t)bi * spp + s) * 4;
            r[0] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint);
            r[1] = make_float4(ray.d.x, ray.d.y, ray.d.z, hit.t);
            r[2] = make_float4(T.r, T.g, T.b, __int_as_float(hit.prim));
            r[3] = make_float4(hit.b1, hit.b2, hit.maxtIn, ray.time);
            a.b.ls[(size_t)bi * spp + s] = make_float4(L.r, L.g, L.b, __int_as_float(code | LS_TAIL | (afterSpecular ? LS_SPEC : 0)));
            atomicAdd(a.b.tailCount + bi, 1u);
            atomicOr(a.b.tailMask + (size_t)bi * a.b.maskWords + (s >> 5), 1u << (s & 31));
            have = false;
        }
        }
    }

    if (PHASE == 1 && !have) {
        // ---- the next camera sample of the batch: (task, sample) = (g / spp, g % spp)
        bool got = false;
        for (;;) {
            const unsigned g = AggregatedAddV(a.counters + WC_NEXT_SAMPLE);
            if (g >= a.totalSamples) break;
            bi = (int)(g / (unsigned)spp);
            s = (int)(g - (unsigned)bi * (unsigned)spp);
            const TaskRec tr = a.b.tasks[bi];
            const int tw = tr.x1 - tr.x0;
            if (a.b.round >= tw * (tr.y1 - tr.y0)) continue;   // (this task has no pixel in this round)
            const int xPos = tr.x0 + a.b.round % tw, yPos = tr.y0 + a.b.round / tw;
            const uint16_t *row = a.b.perm + ((size_t)bi * spp + s) * PERM_STRIDE;
            code = CameraRay(rp, row, a.b.scr + (size_t)bi * SCR_STRIDE, xPos, yPos, &ray);
            got = true;
            break;
        }
        if (!got) {
            WF(W_T) = make_float4(0.f, 0.f, 0.f, __uint_as_float(WB_RETIRED));
            return;
        }
        AggregatedAddV(a.counters + WC_LIVE);
        if (ANIM) FillXfCache(sc, ray.time, &WF(W_AX), a.P);
        WF(W_P) = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint);
        WF(W_D2) = make_float4(ray.d.x, ray.d.y, ray.d.z, ray.maxt);
        WF(W_T) = make_float4(1.f, 1.f, 1.f, __uint_as_float(0u));
        WF(W_L) = make_float4(0.f, 0.f, 0.f, __int_as_float(code));
        WF(W_ID) = make_float4(__int_as_float(bi), __int_as_float(s), ray.time, 0.f);
        WvPushRay(a, VX_PATH, slot);
        return;
    }
    if (PHASE == 2) {
        if (live && !have) {
            // ---- the chain's next parked sample (a bounce-3 vertex: nothing to trace first), or the end of the chain
            while (mbits == 0 && mw + 1 < a.b.maskWords) { ++mw; mbits = __ldcg(a.b.tailMask + (size_t)bi * a.b.maskWords + mw); }
            if (mbits == 0) {
                a.b.mti[bi] = mti;   // the task's stream position for its next pixel
                WF(W_T) = make_float4(0.f, 0.f, 0.f, __uint_as_float(WB_RETIRED));
                live = false;
            } else {
                s = mw * 32 + (__ffs(mbits) - 1);
                mbits &= mbits - 1;
                const float4 lv = __ldcs(a.b.ls + (size_t)bi * spp + s);
                const float4 *r = a.b.rec + ((size_t)bi * spp + s) * 4;
                const float4 r0 = __ldcs(r), r1 = __ldcs(r + 1), r2 = __ldcs(r + 2), r3 = __ldcs(r + 3);
                ray.o = V3(r0.x, r0.y, r0.z); ray.mint = r0.w;
                ray.d = V3(r1.x, r1.y, r1.z); ray.maxt = r1.w;
                hit.t = r1.w;
                T = Spec(r2.x, r2.y, r2.z);
                hit.prim = __float_as_int(r2.w);
                hit.b1 = r3.x; hit.b2 = r3.y; hit.maxtIn = r3.z;
                ray.time = r3.w;
                if (ANIM) FillXfCache(sc, ray.time, &WF(W_AX), a.P);
                L = Spec(lv.x, lv.y, lv.z);
                code = __float_as_int(lv.w);
                afterSpecular = (code & LS_SPEC) != 0;
                code &= ~(LS_SPEC | LS_TAIL);
                bounces = 3;
                have = true; atVertex = true;
            }
        }
        // ---- cooperative RNG top-up: the <= 11 draws of the vertex must be addressable in the two-block window
        const unsigned FULL = 0xffffffffu;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        unsigned m = __ballot_sync(FULL, live && mti >= MT_N);
        while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            const int tb = __shfl_sync(FULL, bi, l);
            WarpAdvanceWindow(a.b.rngCur + (size_t)tb * MT_N, a.b.rngNext + (size_t)tb * MT_N, smt[warp], lane);
            if (lane == l) mti -= MT_N;
        }
        if (!live) return;
    }

    if (PHASE == 1) AggregatedAddV(a.counters + WC_LIVE);
    else a.liveList[(size_t)((a.iter + 1) & 1) * a.b.nBatch + AggregatedAddV(a.counters + WvLiveCounter((a.iter + 1) & 1))] = slot;
    // ---- a vertex to shade: VertexA, the record, the requests
    VertexRandoms u;
    float rrU = 0.f;
    if (PHASE == 1)
        VertexRandomsFromTables(a.b.perm + ((size_t)bi * spp + s) * PERM_STRIDE, a.b.scr + (size_t)bi * SCR_STRIDE, bounces, &u);
    else {
        const uint32_t *rc = a.b.rngCur + (size_t)bi * MT_N, *rn = a.b.rngNext + (size_t)bi * MT_N;
        auto rngF = [&](int k) -> float {
            const int j = mti + k;
            return UIntToFloat(Temper(j < MT_N ? __ldcg(rc + j) : __ldcg(rn + (j - MT_N))));
        };
        if (nLights > 0) {   // integrator.cpp:95,107-108: lightNum, LightSample(rng), BSDFSample(rng)
            u.lnU = rngF(0); u.lsU0 = rngF(1); u.lsU1 = rngF(2);
            u.bsU0 = rngF(4); u.bsU1 = rngF(5); u.bsUc = rngF(6);
        } else u.lnU = u.lsU0 = u.lsU1 = u.bsU0 = u.bsU1 = u.bsUc = 0.f;
        u.pU0 = rngF(k0); u.pU1 = rngF(k0 + 1); u.pUc = rngF(k0 + 2);   // path.cpp:90
        if (bounces > 3) rrU = rngF(k0 + 3);                            // path.cpp:104 (looked at only if the sample's f is not black)
    }
    Vertex v;
    {
        ShadeCtx ctx;
        VertexA<ANIM>(sc, ray, hit, bounces, afterSpecular, u, T, L, ctx, &v, STATS ? &nShadedN : nullptr);
    }
    const bool pathAllowed = bounces != rp.maxDepth;
    const bool roulette = PHASE == 2 && bounces > 3;
    uint32_t bits = ((uint32_t)bounces << WB_BOUNCES_SHIFT) | WB_VERTEX | (afterSpecular ? WB_AFTER_SPEC : 0u) |
                    (v.need[0] ? WB_NEED0 : 0u) | (v.need[1] ? WB_NEED1 : 0u) | (v.need[2] ? WB_NEED2 : 0u) |
                    ((uint32_t)v.lightNum << WB_LIGHT_SHIFT);
#if PBRT_SPECULAR
    if (v.specSampled) bits |= WB_SPEC_SAMPLED;
#endif
    WF(W_P) = make_float4(v.p.x, v.p.y, v.p.z, v.eps);
    WF(W_T) = make_float4(T.r, T.g, T.b, __uint_as_float(bits));
    WF(W_L) = make_float4(L.r, L.g, L.b, __int_as_float(code));
    WF(W_ID) = make_float4(__int_as_float(bi), __int_as_float(s), ray.time, v.pdfP);
    WF(W_V0) = make_float4(v.Li.r, v.Li.g, v.Li.b, v.scalarL);
    WF(W_V1) = make_float4(v.absM, v.weightM, v.pdfM, v.absP);
    if (PHASE == 2) WF(W_CH) = make_float4(__int_as_float(mti), __int_as_float(mw), __uint_as_float(mbits), rrU);
    if (v.need[VX_LIGHT]) WF(W_D0) = make_float4(v.wiW[VX_LIGHT].x, v.wiW[VX_LIGHT].y, v.wiW[VX_LIGHT].z, v.visMaxt);
    if (v.need[VX_MIS]) WF(W_D1) = make_float4(v.wiW[VX_MIS].x, v.wiW[VX_MIS].y, v.wiW[VX_MIS].z, __int_as_float(v.lightNum));
    if (v.need[VX_PATH]) WF(W_D2) = make_float4(v.wiW[VX_PATH].x, v.wiW[VX_PATH].y, v.wiW[VX_PATH].z, INF_F);
#if PBRT_SPECULAR
    if (v.specSampled) WF(W_V2) = make_float4(v.specF.r, v.specF.g, v.specF.b, 0.f);
#endif
    const bool measured = sc.materials[v.material].type == PBRTGPU_MAT_MEASURED;
    bool wroteWo = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (!v.need[k]) continue;
        const bool allowed = k != VX_PATH || pathAllowed;
#if PBRT_SPECULAR
        if (k == VX_PATH && v.specSampled) {   // the value is specF (BSDF::Sample_f, reflection.cpp:572): nothing to evaluate
            if (allowed && PathSampleAlive(v.specF, T, v.absP, v.pdfP, roulette, rrU)) WvPushRay(a, k, slot);
            else if (PHASE == 1) WvSortPush(a, SORT_OTHER, slot);
            continue;
        }
#endif
        if (measured) {
            if (!wroteWo) { WF(W_EWO) = make_float4(v.evalWo.x, v.evalWo.y, v.evalWo.z, __int_as_float(v.material)); wroteWo = true; }
            WF(W_EW0 + k) = make_float4(v.evalWi[k].x, v.evalWi[k].y, v.evalWi[k].z,
                                        __int_as_float(v.evalFlags[k] | (allowed ? W_RAY_ALLOWED : 0) |
                                                       ((k == VX_PATH && roulette) ? W_RR : 0)));
            const unsigned i = AggregatedAddV(a.counters + WC_EVALS);
            a.qEval[i] = (slot << 2) | (uint32_t)k;
        } else {
            const Spec f = VertexEvalAnalytic(sc, v.material, v.evalWo, v.evalWi[k], v.evalFlags[k]);
            WF(W_F0 + k) = make_float4(f.r, f.g, f.b, 0.f);
            if (allowed && (k == VX_PATH ? PathSampleAlive(f, T, v.absP, v.pdfP, roulette, rrU) : !f.IsBlack())) WvPushRay(a, k, slot);
            else if (PHASE == 1 && k == VX_PATH) WvSortPush(a, SORT_OTHER, slot);
        }
    }
    if (PHASE == 1 && !v.need[VX_PATH]) WvSortPush(a, SORT_OTHER, slot);   // no continuation direction at all
    if (STATS && nShadedN) atomicAdd(&a.stats->shadedN, nShadedN);
#undef WF
}

// k_weval's "does direction k get its ray" test: the value is not black and -- the path sample from bounce 4 on -- the path
// survives VertexB's Russian roulette (the same expressions on the same operands)
DEV bool WvRayWanted(const WvArgs &a, uint32_t slot, int k, int word, const Spec &f) {
    if (!(word & W_RAY_ALLOWED)) return false;
    if (!(word & W_RR)) return !f.IsBlack();
    const size_t P = a.P;
    const float4 fT = a.pool[(size_t)W_T * P + slot];
    return PathSampleAlive(f, Spec(fT.x, fT.y, fT.z), a.pool[(size_t)W_V1 * P + slot].w, a.pool[(size_t)W_ID * P + slot].w, true,
                           a.pool[(size_t)W_CH * P + slot].w);
}

// ------------------------------------------------------------------------------------------------------------ k_weval
// measured-BRDF values (IrregIsotropicBRDF::f, reflection.cpp:259-272) of the queued (slot, direction) requests.
//
// dshade.cuh's BxF does it per thread in two kd-tree walks of data-dependent length -- the exact third-nearest distance
// (which fixes the radius the reference's doubling search ends on), then the gather at that radius in the reference's
// visiting order -- and a warp of such threads ran with 8.8 of 32 lanes active (profiles/r01_k_eval_ncu_full.txt).  Here
// the lanes are persistent and both walks are ONE node step per lane and warp iteration: a lane's state is the entry it
// processes next, a step loads that node's 32-byte record and yields the next entry (a child entered directly, or the
// stack top), a lane whose gather ends writes its value and takes the next request at the following refill.
//   entry = node << 1 | flag.  third-nearest walk: flag = "far child of this node, if the splitting plane is still closer
//   than the third-best distance" (the plane distance is recomputed from the parent's record when the entry is popped: no
//   float stack).  gather: flag = "children done: process the node itself" (KdTree::privateLookup's order, kdtree.h:150-183:
//   near child, far child, node).
#if PBRT_KD_INTERLEAVED
#define WEVAL_STACK 48
__global__ void __launch_bounds__(128, K_WEVAL_MIN_BLOCKS) k_weval(const __grid_constant__ WvArgs a) {
    const DScene &sc = a.sc;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned n = a.counters[WC_EVALS];
    const size_t P = a.P;
    int mode = 0;             // 0: idle, 1: third-nearest walk, 2: gather
    uint32_t slot = 0, cur = 0, first_ = 0, nNodes = 0;   // first_: the tree's first node in kdRec
    int k = 0, word = 0, sp = 0;
    V3 p(0, 0, 0);
    float b0 = 0.f, b1 = 0.f, b2 = 0.f, maxDist2 = 0.f, sumW = 0.f;
    Spec v(0.f);
    uint32_t stk[WEVAL_STACK];
    unsigned cPos = 0, cEnd = 0;
    bool exhausted = false;
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, mode == 0);
        if (idle && !(exhausted && cPos >= cEnd)) {
            if (cPos >= cEnd) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(a.counters + WC_EVAL_CUR, 32u);
                base = __shfl_sync(FULL, base, 0);
                if (base >= n) { exhausted = true; cPos = cEnd = 0; }
                else { cPos = base; cEnd = min(base + 32u, n); }
            }
            if (mode == 0) {
                const unsigned idx = cPos + __popc(idle & ((1u << lane) - 1u));
                if (idx < cEnd) {
                    const uint32_t e = a.qEval[idx];
                    slot = e >> 2;
                    k = (int)(e & 3u);
                    const float4 wo = a.pool[(size_t)W_EWO * P + slot], wi = a.pool[(size_t)(W_EW0 + k) * P + slot];
                    word = __float_as_int(wi.w);
                    const PbrtGpuMaterial &m = sc.materials[__float_as_int(wo.w)];
                    bool walk = false;
                    Spec f(0.f);
                    if (m.type == PBRTGPU_MAT_MEASURED) {
                        // the BSDF's one component (measured.cpp:194-210) takes part if its type matches the flags
                        if (Matches(BSDF_REFLECTION | BSDF_GLOSSY, word & 0xffff)) {
                            const PbrtGpuMeasured md = sc.measured[m.measured];
                            first_ = md.firstNode; nNodes = md.nNodes;
                            p = BRDFRemap(V3(wo.x, wo.y, wo.z), V3(wi.x, wi.y, wi.z));
                            b0 = b1 = b2 = INF_F;
                            cur = 0u; sp = 0;
                            walk = true;
                        }
                    } else   // (k_wshade evaluates analytic materials itself; kept for completeness)
                        f = VertexEvalAnalytic(sc, __float_as_int(wo.w), V3(wo.x, wo.y, wo.z), V3(wi.x, wi.y, wi.z), word & 0xffff);
                    if (walk) mode = 1;
                    else {
                        a.pool[(size_t)(W_F0 + k) * P + slot] = make_float4(f.r, f.g, f.b, 0.f);
                        if (WvRayWanted(a, slot, k, word, f)) WvPushRayAny(a, k, slot);
                        else if (k == VX_PATH) WvSortPush(a, SORT_OTHER, slot);
                    }
                }
            }
            cPos = min(cEnd, cPos + (unsigned)__popc(idle));
        }
        const bool more = !(exhausted && cPos >= cEnd);
        if (!__any_sync(FULL, mode != 0)) {
            if (!more) break;
            continue;
        }
        for (int it = 0; it < 64; ++it) {
            if (mode != 0) {
                // ---- one node step, the two walks on one instruction stream (selects, not branches: a branchy form of this
                // body ran with 10 of 32 lanes active, profiles/r02b_k_weval_ncu_full.txt)
                const uint32_t node = cur >> 1;
                const bool first = (cur & 1u) == 0;   // first arrival at the node (else: its flagged entry came off the stack)
                const bool knn = mode == 1;
                const float4 *rec = sc.kdRec + 2 * (size_t)(first_ + node);
                const float4 k0 = __ldg(rec), k1 = __ldg(rec + 1);
                const float splitPos = k0.x;
                const uint32_t bits = __float_as_uint(k0.y);
                const int axis = (int)(bits & 3u);
                const bool interior = axis != 3;
                const bool hasLeft = ((bits >> 2) & 1u) != 0;
                const uint32_t rightChild = bits >> 3;
                const bool hasRight = rightChild < nNodes;
                const float pa = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);   // (a leaf's value is not looked at)
                const float plane2 = (pa - splitPos) * (pa - splitPos);
                const bool leftFirst = pa <= splitPos;
                const bool nearExists = interior && (leftFirst ? hasLeft : hasRight);
                const bool farExists = interior && (leftFirst ? hasRight : hasLeft);
                const uint32_t nearChild = leftFirst ? node + 1 : rightChild, farChild = leftFirst ? rightChild : node + 1;
                const float d2 = DistanceSquared(V3(k0.z, k0.w, k1.x), p);
                // third-nearest walk, first arrival: the node's own point enters the three best (d2 < b0 implies d2 < b1 < b2)
                const bool c2 = knn && first && d2 < b2, c1 = c2 && d2 < b1, c0 = c1 && d2 < b0;
                b2 = c2 ? (c1 ? b1 : d2) : b2;
                b1 = c1 ? (c0 ? b0 : d2) : b1;
                b0 = c0 ? d2 : b0;
                // gather, second arrival (or a leaf): the node's own point is processed (IrregIsoProc, reflection.cpp:40-55)
                if (!knn && (!first || !interior) && d2 < maxDist2) {
                    const float weight = ref_expf(-100.f * d2);
                    v = v + weight * Spec(k1.y, k1.z, k1.w);
                    sumW += weight;
                }
                // what goes on the stack: the entry to come back to (third-nearest: "my far child, if its plane is still
                // close enough then"; gather: "my own point"), and -- gather only -- the far child itself
                if (first && interior && (!knn || farExists)) stk[sp++] = cur | 1u;
                if (!knn && first && farExists && plane2 < maxDist2) stk[sp++] = farChild << 1;
                // where to go: the near child; the far child of a popped third-nearest entry; else whatever the stack holds
                uint32_t next = 0u;
                bool have = false;
                if (first && nearExists) { next = nearChild << 1; have = true; }
                if (knn && !first && plane2 < b2) { next = farChild << 1; have = true; }
                if (!have && sp > 0) { next = stk[--sp]; have = true; }
                cur = next;
                if (!have) {
                    if (knn) {
                        // reflection.cpp:261-271: the radius doubles until a lookup finds more than two samples
                        float r2 = .001f;
                        while (!(b2 < r2) && !(r2 > 1.5f)) r2 *= 2.f;
                        maxDist2 = r2;
                        v = Spec(0.f); sumW = 0.f;
                        mode = 2; cur = 0u; sp = 0;
                    } else {
                        const Spec f = v.Clamp() / sumW;
                        a.pool[(size_t)(W_F0 + k) * P + slot] = make_float4(f.r, f.g, f.b, 0.f);
                        if (WvRayWanted(a, slot, k, word, f)) WvPushRayAny(a, k, slot);
                        else if (k == VX_PATH) WvSortPush(a, SORT_OTHER, slot);
                        mode = 0;
                    }
                }
            }
            if ((it & 1) == 0) continue;   // a vote every second step
            const int nActive = __popc(__ballot_sync(FULL, mode != 0));
            if (nActive == 0 || (more && nActive < 24)) break;
        }
    }
}
#else
__global__ void __launch_bounds__(128, K_WEVAL_MIN_BLOCKS) k_weval(const __grid_constant__ WvArgs a) {
    const unsigned n = a.counters[WC_EVALS];
    const size_t P = a.P;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const uint32_t e = a.qEval[idx];
        const uint32_t slot = e >> 2;
        const int k = (int)(e & 3u);
        const float4 wo = a.pool[(size_t)W_EWO * P + slot], wi = a.pool[(size_t)(W_EW0 + k) * P + slot];
        const int word = __float_as_int(wi.w);
        const Spec f = VertexEval(a.sc, __float_as_int(wo.w), V3(wo.x, wo.y, wo.z), V3(wi.x, wi.y, wi.z), word & 0xffff);
        a.pool[(size_t)(W_F0 + k) * P + slot] = make_float4(f.r, f.g, f.b, 0.f);
        if (WvRayWanted(a, slot, k, word, f)) WvPushRayAny(a, k, slot);
                        else if (k == VX_PATH) WvSortPush(a, SORT_OTHER, slot);
    }
}
#endif

// ----------------------------------------------------------------------------------------------------------- k_wtrace
// persistent ray-queue consumer: k_trace's loop (chunked hand-out, while-while walk) over the vertex wavefront's records
template <bool STATS, bool ANIM>
__global__ void __launch_bounds__(128, K_TRACE_MIN_BLOCKS) k_wtrace(const __grid_constant__ WvArgs a) {
    const DScene &sc = a.sc;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // this iteration's rays: the shadow, MIS and path queues back to back
    const unsigned n0 = a.counters[WC_RAYS + VX_LIGHT], n1 = n0 + a.counters[WC_RAYS + VX_MIS], n = n1 + a.counters[WC_RAYS + VX_PATH];
    float4 *pool = a.pool;
    const size_t P = a.P;
    const uint32_t *q0 = a.qRay, *q1 = a.qRay + P, *q2 = a.qRay + 2 * P;
    Trav t;
#if PBRT_SMEM_STACK
    __shared__ uint32_t sStack[PBRT_SMEM_STACK * 128];
    t.sstk = sStack + threadIdx.x;
#endif
#if PBRT_TOP_LEVELS && defined(__CUDACC__)
    __shared__ __align__(128) float4 sTop[2 * PBRT_TOP_SLOTS];
    __shared__ __align__(8) uint64_t sTopBar;
    StageTopNodes(sTop, &sTopBar, sc.topNodes);
    const float4 *smemTop = sTop;
#elif PBRT_TOP_LEVELS
    const float4 *smemTop = sc.topNodes;   // (host-compiled harness: the table itself)
#else
    const float4 *smemTop = nullptr;
#endif
    bool active = false;
    bool leaf = false;        // the lane stands before a leaf (TravDescend returned 1)
    uint32_t slot = 0;
    int kind = 0, lightNum = 0;
    Ray wray;                 // the ray as queued (world space): MisHitIsLight needs it after the walk
    wray.o = V3(0, 0, 0); wray.d = V3(0, 0, 1); wray.mint = 0.f; wray.maxt = 0.f; wray.time = 0.f;
    TraceStats st;
    st.nodes = st.prims = 0;
    unsigned long long nClosest = 0, nShadow = 0;
    unsigned cPos = 0, cEnd = 0;   // the warp's reserved chunk of queue entries (warp-uniform)
    bool exhausted = false;        // the cursor has passed the end of the queue
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !active);
        if (idle && !(exhausted && cPos >= cEnd)) {
            if (cPos >= cEnd) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(a.counters + WC_TRACE_CUR, (unsigned)TRACE_CHUNK);
                base = __shfl_sync(FULL, base, 0);
                if (base >= n) { exhausted = true; cPos = cEnd = 0; }
                else { cPos = base; cEnd = min(base + (unsigned)TRACE_CHUNK, n); }
            }
            if (!active) {
                const unsigned idx = cPos + __popc(idle & ((1u << lane) - 1u));
                if (idx < cEnd) {
                    kind = idx < n0 ? VX_LIGHT : (idx < n1 ? VX_MIS : VX_PATH);
                    slot = kind == VX_LIGHT ? q0[idx] : (kind == VX_MIS ? q1[idx - n0] : q2[idx - n1]);
                    const float4 r0 = pool[(size_t)W_P * P + slot], r1 = pool[(size_t)(W_D0 + kind) * P + slot];
                    wray.o = V3(r0.x, r0.y, r0.z); wray.mint = r0.w;
                    wray.d = V3(r1.x, r1.y, r1.z);
                    wray.maxt = kind == VX_LIGHT ? r1.w : INF_F;
                    lightNum = kind == VX_MIS ? __float_as_int(r1.w) : 0;
                    wray.time = ANIM ? pool[(size_t)W_ID * P + slot].z : 0.f;
                    if (ANIM) { wray.xfCache = pool + (size_t)W_AX * P + slot; wray.xfStride = a.P; }
                    TravInit(t, wray, kind == VX_LIGHT);
                    if (STATS) { if (kind == VX_LIGHT) ++nShadow; else ++nClosest; }
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
                if (kind == VX_LIGHT)
                    pool[(size_t)W_D0 * P + slot].w = t.found ? 1.f : 0.f;
                else if (kind == VX_MIS) {
#if PBRT_INFINITE
                    Ray q = wray;
                    q.maxt = t.ray.maxt;
                    Hit h = t.hit;
                    h.t = t.ray.maxt;
                    const bool counts = MisRayContributes<ANIM>(sc, q, h, t.found, lightNum);   // landed on the light / escaped to it
                    pool[(size_t)W_D1 * P + slot].w = counts ? 1.f : 0.f;
#else
                    bool onLight = false;
                    if (t.found) {
                        Ray q = wray;
                        q.maxt = t.ray.maxt;
                        Hit h = t.hit;
                        h.t = t.ray.maxt;
                        onLight = MisHitIsLight<ANIM>(sc, q, h, lightNum);
                    }
                    pool[(size_t)W_D1 * P + slot].w = onLight ? 1.f : 0.f;
#endif
                } else {
                    pool[(size_t)W_HIT * P + slot] = make_float4(__int_as_float(t.found ? t.hit.prim : -1), t.hit.b1, t.hit.b2, t.hit.maxtIn);
                    pool[(size_t)W_D2 * P + slot].w = t.ray.maxt;
                    if (a.sortList) {   // a vertex to shade next (not the bounce-3 vertex, which is only parked)?
                        const uint32_t bits = __float_as_uint(pool[(size_t)W_T * P + slot].w);
                        const bool park = (bits & WB_VERTEX) && ((bits >> WB_BOUNCES_SHIFT) & 0xffu) == 2u;
                        WvSortPush(a, (t.found && !park) ? SORT_VERTEX : SORT_OTHER, slot);
                    }
                }
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
