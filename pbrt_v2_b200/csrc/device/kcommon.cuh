// kcommon.cuh -- what the render kernels of pbrtgpu.cu / kvertex.cuh / kwave.cuh share: the per-batch argument block, the
// device-side counters, the warp-aggregated counter increment and the trace kernels' tuning constants.  A header of its own
// so that the CPU test stage can compile the kernels' source for the host (tests/native/device_on_host.cpp).
#pragma once
#include <cstdint>
#include "../../../include/pbrtgpu.h"
#include "dmath.cuh"

using namespace dev;   // (the kernels live in the global namespace)

#ifndef __CUDACC__
#ifndef __grid_constant__
#define __grid_constant__
#endif
#endif

struct TaskRec {
    int taskNum;
    int x0, x1, y0, y1;   // Sampler::ComputeSubWindow tile
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
    unsigned *tailCount; // [nBatch] number of samples the task parked in phase 1 (sort key of phase 2)
    uint32_t *tailMask;  // [nBatch][maskWords] bit s: sample s of the task's current pixel is parked (null: not kept)
    int maskWords;       // max(1, spp / 32)
    const unsigned *order; // [nBatch] phase 2: thread i serves task order[i] (tasks with parked samples first) or null
};

struct DeviceStats {
    unsigned long long closestRays, shadowRays, nodes, prims, shadedN, cameraSamples;
};

// persistent ray-queue consumers (k_trace, k_wtrace)
#ifndef K_TRACE_MIN_BLOCKS
#define K_TRACE_MIN_BLOCKS 7
#endif
#define TRACE_CHUNK 64           // queue entries a warp reserves with one atomic
// How a warp of the trace kernels walks between two refills.
//   0  one node step (TravStep: a box test, or a leaf's primitives) per lane and warp iteration, a ballot after each
//      (round 1: 17.7 of 32 lanes active, profiles/r01_k_trace_ncu_full.txt)
//   1  "while-while", warp-uniform: up to TRACE_DESCEND_STEPS rounds in which every lane that is not yet before a leaf does
//      one box test, then the lanes that stand before a leaf test its primitives together.  (The first form of this,
//      with the descend loop inside a per-lane function, let the lanes drift apart: 4.1 of 32 active and 5.3 ms instead of
//      0.87 ms per launch, profiles/r02a_k_trace_ncu_full.txt -- the loop must be the WARP's, with a vote per round.)
//   2  one box test per walking lane and iteration as in 0, but leaves are tested in batches: see the kernels
#ifndef PBRT_TRACE_LOOP
#define PBRT_TRACE_LOOP 0
#endif
#ifndef PBRT_LEAF_BATCH
#define PBRT_LEAF_BATCH 12
#endif
#ifndef TRACE_DESCEND_STEPS
#define TRACE_DESCEND_STEPS 12   // node tests per lane before the warp looks at its leaves / refills
#endif
#ifndef TRACE_REFILL_BELOW
#define TRACE_REFILL_BELOW 24    // refill the warp from the queue when fewer lanes than this are still walking
#endif
