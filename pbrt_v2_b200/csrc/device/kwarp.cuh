// kwarp.cuh -- warp-level helpers of the render kernels (kvertex.cuh, kwave.cuh).  Not part of kcommon.cuh: the LD table
// kernels' host harness (tests/native/tables_on_host.cpp) includes that one and shims no warp intrinsics.
#pragma once
#include "kcommon.cuh"

// warp-aggregated atomicAdd(ctr, 1) for the currently converged lanes
DEV unsigned AggregatedAddV(unsigned *ctr) {
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(ctr, (unsigned)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

