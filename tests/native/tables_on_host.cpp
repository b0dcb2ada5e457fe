// tables_on_host.cpp -- TEST INFRASTRUCTURE.  The LD sample-table kernels (device/dtables.cuh: k_tables, k_tables_general)
// run on the host under a warp emulation: 32 host threads play the lanes of one warp, __syncwarp() is a barrier among
// them, shared memory is a buffer they share.  The kernels use no other warp primitive (no shuffles, no ballots), so
// this executes the kernel source as it is, for any samples-per-pixel and any Sample layout, and the CPU test stage can
// compare the tables with the oracle bit for bit.  Build: g++ -fvisibility=hidden -Wl,-Bsymbolic -pthread (see
// tests/test_device_on_host.py); never linked into the product.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <pthread.h>
#include <cuda_runtime.h>

#define DOH_EXPORT extern "C" __attribute__((visibility("default")))
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#undef __launch_bounds__
#define __launch_bounds__(...)
using std::isnan;
using std::isinf;
using std::min;
using std::max;

// ---- the emulated execution context
static thread_local uint3 threadIdx, blockIdx;
static pthread_barrier_t g_warpBarrier;
alignas(16) uint32_t smem[64 * 1024];          // `extern __shared__ uint32_t smem[]` of the kernels (256 KB)
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&g_warpBarrier); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { T v; memcpy(&v, p, sizeof(T)); std::atomic_signal_fence(std::memory_order_seq_cst); return v; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
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

#include "../../pbrt_v2_b200/csrc/device/dtables.cuh"

namespace {

template <typename K> void LaunchWarp(K kernel) {   // one block of one warp: <<<1, 32>>>
    pthread_barrier_init(&g_warpBarrier, nullptr, 32);
    std::vector<std::thread> lanes;
    for (unsigned l = 0; l < 32; ++l)
        lanes.emplace_back([=]() {
            threadIdx.x = l; threadIdx.y = threadIdx.z = 0;
            blockIdx.x = blockIdx.y = blockIdx.z = 0;
            kernel();
        });
    for (std::thread &t : lanes) t.join();
    pthread_barrier_destroy(&g_warpBarrier);
}

}  // namespace

DOH_EXPORT int toh_paired_build(void) { return PBRT_TABLES_PAIRED; }
static int g_split = 0;
// 1: run k_tables the way the render path launches it -- the stream walk (stages 1) and the shuffles (stages 2) as two launches
DOH_EXPORT void toh_set_split(int on) { g_split = on; }   // which variant of the kernel source this harness holds

// pbrtgpu_ld_samples with the table kernel executed by the warp emulation.  general != 0 forces k_tables_general even for
// the path integrator's layout; rounds > 0 additionally runs that many further pixels of the task first (the RNG state
// handed from pixel to pixel through rngCur / rngNext / mti) and returns the LAST pixel's samples -- without path tails
// in between, which is what the oracle's LDPixelSample gives when called back to back.
DOH_EXPORT int toh_ld_samples(const PbrtGpuRenderDesc *rd, int taskNum, int x0, int y0, int x1, int y1, int general, int rounds,
                              float *out) {
    const int spp = rd->samplesPerPixel;
    SampleLayoutDev L;
    TableArrays T;
    BuildLayout(*rd, &L, &T);
    const bool useGeneral = general || !IsPathLayout(*rd);
    TaskRec tr;
    tr.taskNum = taskNum; tr.x0 = x0; tr.x1 = x1; tr.y0 = y0; tr.y1 = y1;
    std::vector<uint16_t> perm((size_t)spp * L.rowStride + 64);
    std::vector<uint32_t> scr(SCR_STRIDE), cur(MT_N), next(MT_N);
    std::vector<float4> rec((size_t)spp * 4 + 4);
    std::vector<uint8_t> tgt((size_t)spp * T.tgtPerSample + 64);
    int mti = 0;
    BatchArgs b;
    memset(&b, 0, sizeof(b));
    b.tasks = &tr; b.nBatch = 1; b.perm = perm.data(); b.scr = scr.data(); b.rngCur = cur.data(); b.rngNext = next.data();
    b.mti = &mti; b.rec = rec.data();
    const TableLaunch tl = useGeneral ? PlanTablesGeneral(spp, T.nArrays) : PlanTables(spp);
    const int G = tl.G;
    if (tl.smemBytes / tl.wpb > sizeof(smem)) return 2;
    for (int r = 0; r <= rounds; ++r) {
        b.round = r;
        if (useGeneral) LaunchWarp([&]() { k_tables_general<1>(b, spp, G, T, tgt.data()); });
        else if (g_split) {
            const int G2 = PlanTables(spp, 2).G;
            LaunchWarp([&]() { k_tables<1>(b, spp, 0, 1); });
            LaunchWarp([&]() { k_tables<1>(b, spp, G2, 2); });
        } else LaunchWarp([&]() { k_tables<1>(b, spp, G); });
    }
    const int tw = x1 - x0;
    tr.x0 = x0 + rounds % tw; tr.y0 = y0 + rounds / tw;   // ExpandSample offsets by the pixel in tasks[0]
    const int nf = FloatsPerSample(*rd);
    for (int s = 0; s < spp; ++s) ExpandSample(b, s, rd->shutterOpen, rd->shutterClose, L, nf, out);
    return 0;
}
