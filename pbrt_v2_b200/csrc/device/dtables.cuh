// dtables.cuh -- the LD sample-table kernels (LDSampler::GetMoreSamples / LDPixelSample) and the per-batch argument block.
// Kept in a header so that the CPU test stage can run the very same kernel source under a 32-thread warp emulation
// (tests/native/tables_on_host.cpp).
#pragma once
#include <algorithm>
#include <cstring>
#include "../../../include/pbrtgpu.h"
#include "dmath.cuh"
#include "drng.cuh"
#include "kcommon.cuh"

using namespace dev;   // (the kernels live in the global namespace, like the rest of pbrtgpu.cu)

// Round-2 experiment, OFF in the shipped kernels: k_tables' Fisher-Yates chains two steps at a time (see stage 2).
#ifndef PBRT_TABLES_PAIRED
#define PBRT_TABLES_PAIRED 0
#endif

// --------------------------------------------------------------------- k_tables
// LDSampler::GetMoreSamples / LDPixelSample for one pixel of each task (one warp per task).
//
// The pixel's 37 + 52*spp draws are consumed array by array (drng.cuh): scramble word(s), spp
// draws that the reference spends on 1-element shuffles (skipped: only the stream position moves),
// spp Fisher-Yates draws.  Stage 1 turns the latter into swap targets `i + y % (spp - i)` and parks
// them in the task's (not yet used) `rec` scratch; stage 2 runs the 26 inherently sequential
// shuffles, one array per lane, on u16 index arrays in shared memory, with the targets prefetched
// from global memory; stage 3 writes one 64-byte row of shuffled indices per camera sample.
//
// `stages` selects what a launch does: 1 = stage 1 only (needs one MT19937 block of shared memory per warp: many warps
// per SM, which hides the latencies of the block regeneration and of the modulo), 2 = stages 2 + 3 only (needs the index
// arrays: 2 KB of shared memory per chain at 1024 spp, so few warps per SM whatever the rest does), 3 = all in one launch.
// The render path launches 1 then 2 (ncu, profiles/r02_k_tables_*: in one launch the whole kernel ran at 4 warps per SM).
template <int WPB>
__global__ void __launch_bounds__(32 * WPB) k_tables(BatchArgs a, int spp, int G, int stages = 3) {
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = spp + 2;   // +2 u16 (one bank) of padding per array
    const size_t mtWords = (stages & 1) ? MT_N : 0;
    const size_t warpWords = mtWords + ((stages & 2) ? ((size_t)G * stride + 1) / 2 : 0);
    uint32_t *mt = smem + warp * warpWords;
    uint16_t *P = reinterpret_cast<uint16_t *>(mt + mtWords);
    const int bi = blockIdx.x * WPB + warp;
    if (bi >= a.nBatch) return;
    const TaskRec tr = a.tasks[bi];
    const int npx = (tr.x1 - tr.x0) * (tr.y1 - tr.y0);
    if (a.round >= npx) return;

    uint16_t *permOut = a.perm + (size_t)bi * spp * PERM_STRIDE;
    uint32_t *scrOut = a.scr + (size_t)bi * SCR_STRIDE;
    uint16_t *Tg = reinterpret_cast<uint16_t *>(a.rec + (size_t)bi * spp * 4);   // [N_ARRAYS][spp] swap targets

    if (stages & 1) {
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

    // ---- stage 1: walk the draw stream
    for (int arr = 0; arr < N_ARRAYS; ++arr) {
        const int ns = (arr <= 1 || arr >= 17) ? 2 : 1;   // 2-D arrays draw two scramble words
        for (int k = 0; k < ns; ++k) {
            if (mti >= MT_N) { WarpRegenMT(mt, lane); mti = 0; }
            if (lane == 0) scrOut[ScrambleSlot(arr) + k] = Temper(mt[mti]);
            ++mti;
        }
        // Shuffle(samples + i, 1, dims, rng) for every pixel sample: `RandomUInt() % 1`, no effect
        for (int n = spp; n > 0;) {
            if (mti >= MT_N) { WarpRegenMT(mt, lane); mti = 0; }
            const int take = min(n, MT_N - mti);
            mti += take;
            n -= take;
        }
        // Shuffle(samples, nPixel, dims, rng): other = i + RandomUInt() % (count - i), montecarlo.h:177
        uint16_t *Ta = Tg + (size_t)arr * spp;
        for (int i0 = 0; i0 < spp;) {
            if (mti >= MT_N) { WarpRegenMT(mt, lane); mti = 0; }
            const int take = min(spp - i0, MT_N - mti);
            for (int j = lane; j < take; j += 32) {
                const uint32_t i = (uint32_t)(i0 + j);
                const uint32_t y = Temper(mt[mti + j]);
                Ta[i] = (uint16_t)(i + y % ((uint32_t)spp - i));
            }
            mti += take;
            i0 += take;
        }
    }
    __syncwarp();
    // the stream position after the tables: cur = block in use, next = the one after it
    for (int j = lane; j < MT_N; j += 32) cur[j] = mt[j];
    __syncwarp();
    WarpRegenMT(mt, lane);
    for (int j = lane; j < MT_N; j += 32) next[j] = mt[j];
    if (lane == 0) a.mti[bi] = mti;
    __threadfence_block();
    __syncwarp();
    }   // stage 1
    if (!(stages & 2)) return;

    // ---- stages 2+3, G arrays at a time
    for (int a0 = 0; a0 < N_ARRAYS; a0 += G) {
        const int ng = min(N_ARRAYS, a0 + G) - a0;
        for (int l = 0; l < ng; ++l)
            for (int i = lane; i < spp; i += 32) P[l * stride + i] = (uint16_t)i;
        __syncwarp();
        if (lane < ng) {   // Fisher-Yates of the identity (Shuffle, montecarlo.h:174-181)
            uint16_t *p = P + lane * stride;
            const uint16_t *t = Tg + (size_t)(a0 + lane) * spp;
            // spp is a power of two >= 1.  The swap targets come from global memory eight at a time (one 128-bit load)
            // and TWO groups ahead of the chain: the chain itself is ~40 cycles per step of shared-memory round trips, a
            // group's load ~600 cycles away -- fetched only when needed it tripled the time per step (profiles/r02_k_tables_*)
            const bool wide = spp >= 8;
            uint4 q1 = make_uint4(0, 0, 0, 0), q2 = make_uint4(0, 0, 0, 0);
            if (wide) {
                q1 = __ldcg(reinterpret_cast<const uint4 *>(t));
                if (8 < spp) q2 = __ldcg(reinterpret_cast<const uint4 *>(t + 8));
            }
            for (int i0 = 0; i0 < spp; i0 += 8) {
                uint16_t tj[8];
                if (wide) {
                    const uint4 q = q1;
                    q1 = q2;
                    if (i0 + 16 < spp) q2 = __ldcg(reinterpret_cast<const uint4 *>(t + i0 + 16));
                    tj[0] = (uint16_t)(q.x & 0xffffu); tj[1] = (uint16_t)(q.x >> 16);
                    tj[2] = (uint16_t)(q.y & 0xffffu); tj[3] = (uint16_t)(q.y >> 16);
                    tj[4] = (uint16_t)(q.z & 0xffffu); tj[5] = (uint16_t)(q.z >> 16);
                    tj[6] = (uint16_t)(q.w & 0xffffu); tj[7] = (uint16_t)(q.w >> 16);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) tj[k] = (i0 + k < spp) ? __ldcg(t + i0 + k) : (uint16_t)0;
                }
#if PBRT_TABLES_PAIRED
                // two steps per shared-memory round trip: all four loads first, then the values step k+1 would have seen
                // after step k's stores are patched in registers (step k puts `a` at j0: that is position i+1 if
                // j0 == i+1, and it is what step k+1 finds at j1 if j1 == j0), then the four stores in program order
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    const int i = i0 + k;
                    if (i + 1 < spp) {
                        const int j0 = tj[k], j1 = tj[k + 1];
                        const uint16_t a = p[i], b = p[j0];
                        uint16_t c = p[i + 1], d = p[j1];
                        if (j0 == i + 1) c = a;
                        if (j1 == j0) d = a;
                        if (j1 == i + 1) d = c;
                        p[i] = b;
                        p[j0] = a;
                        p[i + 1] = d;
                        p[j1] = c;
                    } else if (i < spp) {   // spp == 1
                        const int j = tj[k];
                        const uint16_t tmp = p[i];
                        p[i] = p[j];
                        p[j] = tmp;
                    }
                }
#else
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = i0 + k;
                    if (i < spp) {
                        const int j = tj[k];
                        const uint16_t tmp = p[i];
                        p[i] = p[j];
                        p[j] = tmp;
                    }
                }
#endif
            }
        }
        __syncwarp();
        if (ng == N_ARRAYS) {
            // one 64-byte row per camera sample, assembled in registers: coalesced 128-bit stores
            for (int s = lane; s < spp; s += 32) {
                uint32_t w[16];
#pragma unroll
                for (int l = 0; l < 13; ++l) w[l] = (uint32_t)P[(2 * l) * stride + s] | ((uint32_t)P[(2 * l + 1) * stride + s] << 16);
                w[13] = w[14] = w[15] = 0u;
                uint4 *dst = reinterpret_cast<uint4 *>(permOut + (size_t)s * PERM_STRIDE);
                dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
                dst[2] = make_uint4(w[8], w[9], w[10], w[11]);
                dst[3] = make_uint4(w[12], w[13], w[14], w[15]);
            }
        } else {
            for (int s = lane; s < spp; s += 32)
                for (int l = 0; l < ng; ++l) permOut[(size_t)s * PERM_STRIDE + a0 + l] = P[l * stride + s];
        }
        __syncwarp();
    }
}

// expand one task's tables into the reference's Sample layout (test hook)
// ------------------------------------------------------------- k_tables_general
// LDPixelSample for any Sample layout (arrays with more than one value per pixel sample: the light / BSDF
// sample arrays DirectLightingIntegrator requests, directlighting.cpp:53-78).  Per array (montecarlo.h:300-325):
// scramble word(s); for every pixel sample a Fisher-Yates shuffle of its n values (n draws each: skipped when
// n == 1); one Fisher-Yates shuffle of the nPixelSamples blocks.  Stage 1 walks the draw stream and turns every
// draw into its swap target (elementwise, so MT19937 block boundaries need no care); stage 2 runs the block
// shuffles one array per lane (shared memory) and the in-block shuffles one block per lane (registers);
// stage 3 writes, for camera sample s and value k, the original sequence index  b*n + inner_b[k]  (b = block
// that ended up at position s).
template <int WPB>
__global__ void __launch_bounds__(32 * WPB) k_tables_general(BatchArgs a, int spp, int G, TableArrays L, uint8_t *tgtAll) {
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = spp + 2;
    const size_t warpWords = MT_N + ((size_t)G * stride + 1) / 2;
    uint32_t *mt = smem + warp * warpWords;
    uint16_t *P = reinterpret_cast<uint16_t *>(mt + MT_N);
    const int bi = blockIdx.x * WPB + warp;
    if (bi >= a.nBatch) return;
    const TaskRec tr = a.tasks[bi];
    const int npx = (tr.x1 - tr.x0) * (tr.y1 - tr.y0);
    if (a.round >= npx) return;

    uint32_t *cur = a.rngCur + (size_t)bi * MT_N, *next = a.rngNext + (size_t)bi * MT_N;
    int mti;
    if (a.round == 0) {
        if (lane == 0) SeedMT(mt, (uint32_t)tr.taskNum);
        mti = MT_N;
    } else {
        mti = a.mti[bi];
        const uint32_t *src = cur;
        if (mti >= MT_N) { src = next; mti -= MT_N; }
        for (int j = lane; j < MT_N; j += 32) mt[j] = __ldcg(src + j);
    }
    __syncwarp();

    uint16_t *permOut = a.perm + (size_t)bi * spp * L.rowStride;
    uint32_t *scrOut = a.scr + (size_t)bi * SCR_STRIDE;
    uint8_t *tgt = tgtAll + (size_t)bi * spp * L.tgtPerSample;

    // ---- stage 1: walk the draw stream
    for (int arr = 0; arr < L.nArrays; ++arr) {
        const int n = L.n[arr];
        for (int k = 0; k < L.dims[arr]; ++k) {
            if (mti >= MT_N) { WarpRegenMT(mt, lane); mti = 0; }
            if (lane == 0) scrOut[L.scr[arr] + k] = Temper(mt[mti]);
            ++mti;
        }
        if (n == 1) {   // Shuffle(samples + i, 1, dims, rng): `RandomUInt() % 1`, no effect
            for (int left = spp; left > 0;) {
                if (mti >= MT_N) { WarpRegenMT(mt, lane); mti = 0; }
                const int take = min(left, MT_N - mti);
                mti += take;
                left -= take;
            }
        } else {        // Shuffle(samples + i * n, n, dims, rng) for pixel sample i: draw d = i*n + k swaps k with k + y % (n - k)
            uint8_t *tin = tgt + (size_t)spp * L.tgtInner[arr];
            const int total = spp * n;
            for (int d0 = 0; d0 < total;) {
                if (mti >= MT_N) { WarpRegenMT(mt, lane); mti = 0; }
                const int take = min(total - d0, MT_N - mti);
                for (int j = lane; j < take; j += 32) {
                    const uint32_t d = (uint32_t)(d0 + j), k = d % (uint32_t)n;
                    const uint32_t y = Temper(mt[mti + j]);
                    tin[d] = (uint8_t)(k + y % ((uint32_t)n - k));
                }
                mti += take;
                d0 += take;
            }
        }
        uint16_t *Ta = reinterpret_cast<uint16_t *>(tgt + (size_t)spp * L.tgtBlock[arr]);
        for (int i0 = 0; i0 < spp;) {   // Shuffle(samples, nPixel, dims * n, rng)
            if (mti >= MT_N) { WarpRegenMT(mt, lane); mti = 0; }
            const int take = min(spp - i0, MT_N - mti);
            for (int j = lane; j < take; j += 32) {
                const uint32_t i = (uint32_t)(i0 + j);
                const uint32_t y = Temper(mt[mti + j]);
                Ta[i] = (uint16_t)(i + y % ((uint32_t)spp - i));
            }
            mti += take;
            i0 += take;
        }
    }
    __syncwarp();
    for (int j = lane; j < MT_N; j += 32) cur[j] = mt[j];
    __syncwarp();
    WarpRegenMT(mt, lane);
    for (int j = lane; j < MT_N; j += 32) next[j] = mt[j];
    if (lane == 0) a.mti[bi] = mti;
    __threadfence_block();
    __syncwarp();

    // ---- stages 2+3, G arrays at a time
    for (int a0 = 0; a0 < L.nArrays; a0 += G) {
        const int ng = min(L.nArrays, a0 + G) - a0;
        for (int l = 0; l < ng; ++l)
            for (int i = lane; i < spp; i += 32) P[l * stride + i] = (uint16_t)i;
        __syncwarp();
        if (lane < ng) {   // block shuffle: Fisher-Yates of the identity
            uint16_t *p = P + lane * stride;
            const uint16_t *t = reinterpret_cast<const uint16_t *>(tgt + (size_t)spp * L.tgtBlock[a0 + lane]);
            for (int i = 0; i < spp; ++i) {
                const int j = __ldcg(t + i);
                const uint16_t tmp = p[i];
                p[i] = p[j];
                p[j] = tmp;
            }
        }
        for (int l = 0; l < ng; ++l) {   // in-block shuffles, one pixel sample per lane, in place over the targets
            const int n = L.n[a0 + l];
            if (n == 1) continue;
            uint8_t *tin = tgt + (size_t)spp * L.tgtInner[a0 + l];
            for (int i = lane; i < spp; i += 32) {
                uint8_t p[MAX_ARRAY_SAMPLES];
                for (int k = 0; k < n; ++k) p[k] = (uint8_t)k;
                for (int k = 0; k < n; ++k) {
                    const int j = __ldcg(tin + (size_t)i * n + k);
                    const uint8_t tmp = p[k];
                    p[k] = p[j];
                    p[j] = tmp;
                }
                for (int k = 0; k < n; ++k) tin[(size_t)i * n + k] = p[k];
            }
        }
        __threadfence_block();
        __syncwarp();
        for (int s = lane; s < spp; s += 32) {
            uint16_t *row = permOut + (size_t)s * L.rowStride;
            for (int l = 0; l < ng; ++l) {
                const int n = L.n[a0 + l];
                const uint32_t b = P[l * stride + s];
                if (n == 1) row[L.row[a0 + l]] = (uint16_t)b;
                else {
                    const uint8_t *tin = tgt + (size_t)spp * L.tgtInner[a0 + l] + (size_t)b * n;
                    for (int k = 0; k < n; ++k) row[L.row[a0 + l] + k] = (uint16_t)(b * (uint32_t)n + __ldcg(tin + k));
                }
            }
        }
        __syncwarp();
    }
}

// One camera sample's values out of its table row (LDPixelSample's final loops, montecarlo.cpp:238-261): the image, lens and
// time values and every 1-D / 2-D array in Sample order.  pbrtgpu_ld_samples' kernel.
DEV void ExpandSample(const BatchArgs &a, int s, float shutterOpen, float shutterClose, const SampleLayoutDev &L,
                      int floatsPerSample, float *out) {
    const TaskRec tr = a.tasks[0];
    const uint16_t *row = a.perm + (size_t)s * L.rowStride;
    const uint32_t *scr = a.scr;
    float *o = out + (size_t)s * floatsPerSample;
    o[0] = tr.x0 + VanDerCorput(row[0], scr[0]);
    o[1] = tr.y0 + Sobol2(row[0], scr[1]);
    o[2] = VanDerCorput(row[1], scr[2]);
    o[3] = Sobol2(row[1], scr[3]);
    o[4] = Lerp(VanDerCorput(row[2], scr[4]), shutterOpen, shutterClose);
    o += 5;
    for (int j = 0; j < L.n1D; ++j)
        for (int k = 0; k < L.size1D[j]; ++k) *o++ = VanDerCorput(row[L.row1D[j] + k], scr[L.scr1D[j]]);
    for (int j = 0; j < L.n2D; ++j)
        for (int k = 0; k < L.size2D[j]; ++k) {
            *o++ = VanDerCorput(row[L.row2D[j] + k], scr[L.scr2D[j]]);
            *o++ = Sobol2(row[L.row2D[j] + k], scr[L.scr2D[j] + 1]);
        }
}

// ------------------------------------------------------------------ host side: layouts and launch plans
// Sample layout of a render descriptor -> row / scramble positions (drng.cuh)
inline bool IsPathLayout(const PbrtGpuRenderDesc &rd) {
    if (rd.n1D != N_1D || rd.n2D != N_2D) return false;
    for (int i = 0; i < rd.n1D; ++i) if (rd.size1D[i] != 1) return false;
    for (int i = 0; i < rd.n2D; ++i) if (rd.size2D[i] != 1) return false;
    return true;
}
inline void BuildLayout(const PbrtGpuRenderDesc &rd, SampleLayoutDev *L, TableArrays *T) {
    memset(L, 0, sizeof(*L));
    memset(T, 0, sizeof(*T));
    L->n1D = rd.n1D; L->n2D = rd.n2D;
    int row = 3, scr = 5, arr = 0, tgt = 0;
    auto add = [&](int n, int dims, int rowAt, int scrAt) {
        T->n[arr] = (uint16_t)n; T->dims[arr] = (uint8_t)dims; T->row[arr] = (uint16_t)rowAt; T->scr[arr] = (uint16_t)scrAt;
        T->tgtInner[arr] = (uint16_t)tgt;
        if (n > 1) tgt += (n + 1) & ~1;
        T->tgtBlock[arr] = (uint16_t)tgt;
        tgt += 2;
        ++arr;
    };
    add(1, 2, 0, 0); add(1, 2, 1, 2); add(1, 1, 2, 4);   // image, lens, time (montecarlo.cpp:228-230)
    for (int j = 0; j < rd.n1D; ++j) {
        L->size1D[j] = rd.size1D[j]; L->row1D[j] = (uint16_t)row; L->scr1D[j] = (uint16_t)scr;
        add(rd.size1D[j], 1, row, scr);
        row += rd.size1D[j]; scr += 1;
    }
    for (int j = 0; j < rd.n2D; ++j) {
        L->size2D[j] = rd.size2D[j]; L->row2D[j] = (uint16_t)row; L->scr2D[j] = (uint16_t)scr;
        add(rd.size2D[j], 2, row, scr);
        row += rd.size2D[j]; scr += 2;
    }
    L->rowStride = (row + 7) / 8 * 8;
    T->nArrays = arr; T->rowStride = L->rowStride; T->tgtPerSample = tgt;
}
inline int FloatsPerSample(const PbrtGpuRenderDesc &rd) {
    int n = 5;
    for (int j = 0; j < rd.n1D; ++j) n += rd.size1D[j];
    for (int j = 0; j < rd.n2D; ++j) n += 2 * rd.size2D[j];
    return n;
}


struct TableLaunch {
    int G, wpb;
    size_t smemBytes;
};
// stages: 3 = the one-launch form (an MT19937 block + the index arrays per warp), 2 = stages 2 + 3 only (index arrays only)
inline TableLaunch PlanTables(int spp, int stages = 3) {
    TableLaunch t;
    const size_t perArray = (size_t)2 * (spp + 2);   // one u16 index array per (0,2)-sequence array
    const size_t budget = 56 * 1024;
    const size_t mtBytes = (stages & 1) ? MT_N * 4 : 0;
    size_t g = (budget - mtBytes) / perArray;
    t.G = g >= N_ARRAYS ? N_ARRAYS : (int)std::max<size_t>(1, g);
    const size_t perWarp = mtBytes + ((size_t)t.G * perArray + 3) / 4 * 4;
    t.wpb = (int)std::max<size_t>(1, std::min<size_t>(4, (226 * 1024) / perWarp));
    t.smemBytes = perWarp * t.wpb;
    return t;
}

inline TableLaunch PlanTablesGeneral(int spp, int nArrays) {
    TableLaunch t;
    const size_t perArray = (size_t)2 * (spp + 2);
    const size_t budget = 56 * 1024;
    size_t g = (budget - MT_N * 4) / perArray;
    t.G = (int)std::max<size_t>(1, std::min<size_t>(g, (size_t)std::min(32, nArrays)));   // one array per lane in the block shuffle
    const size_t perWarp = MT_N * 4 + ((size_t)t.G * perArray + 3) / 4 * 4;
    t.wpb = (int)std::max<size_t>(1, std::min<size_t>(4, (226 * 1024) / perWarp));
    t.smemBytes = perWarp * t.wpb;
    return t;
}
