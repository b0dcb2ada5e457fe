// drng.cuh -- MT19937 streams and the (0,2)-sequence sample tables on the device.
//
// Reference: RNG (core/rng.cpp:43-107), LDPixelSample (core/montecarlo.cpp:200-252),
// LDShuffleScrambled1D/2D, Shuffle, VanDerCorput, Sobol2 (core/montecarlo.h:174-181,
// 281-327).  SURVEY.md Appendix B gives the draw accounting this file relies on:
//
//   per pixel, 26 arrays in this order: image(2-D) lens(2-D) time(1-D),
//   14 x 1-D, 9 x 2-D; a 1-D array consumes 1 + 2*spp draws, a 2-D array
//   2 + 2*spp:  [scramble(s)] [spp draws of `RandomUInt() % 1` = no-ops]
//   [spp Fisher-Yates draws: other = i + RandomUInt() % (spp - i)]
//
// Because every (0,2) value is a pure function of (index, scramble), the tables
// are stored as the shuffled INDEX permutation (u16 per array per sample, 64 B
// per camera sample) plus 37 scramble words per pixel; the consumer rebuilds
// VanDerCorput/Sobol2 values on the fly.  The Fisher-Yates pass itself is
// inherently sequential per array, so the 26 arrays of a pixel are shuffled by
// 26 lanes of one warp, in shared memory.
//
// Per-task RNG state in HBM: cur[624] (the MT block being consumed), next[624]
// (the block after it) and mti (position in cur; may run past 624 into next).
// Keeping two blocks lets the path kernel read "draw mti+k" for the <= 11 draws
// of one bounce without ever regenerating in divergent code.
#pragma once
#include "dmath.cuh"

namespace dev {

constexpr int MT_N = 624;
constexpr int MT_M = 397;
constexpr int N_ARRAYS = 26;      // 2 + 1 + 14 + 9
constexpr int N_1D = 14, N_2D = 9;
constexpr int N_SCRAMBLES = 37;   // 2+2+1 + 14 + 18
constexpr int PERM_STRIDE = 32;   // u16 per sample row (26 used) = 64 bytes
constexpr int SCR_STRIDE = 104;   // scramble words per task: 5 + n1D + 2*n2D <= 5 + 32 + 64

// Sample::n1D / n2D (core/sampler.h:93-116) and where each requested array sits in a camera sample's table
// row and in the task's scramble block.  Rows: [0] image [1] lens [2] time, then the 1-D arrays' entries,
// then the 2-D arrays'; every entry is a u16 index into the array's (0,2)-sequence (value = VanDerCorput /
// Sobol2 of that index with the array's scramble words).  For the path integrator's 14 + 9 single-value
// arrays this is the fixed layout the specialised k_tables writes (row1D[j] = 3+j, row2D[j] = 17+j).
constexpr int MAX_SAMPLE_ARRAYS = 32;
constexpr int MAX_TABLE_ARRAYS = 3 + 2 * MAX_SAMPLE_ARRAYS;
constexpr int MAX_ARRAY_SAMPLES = 64;   // values per pixel sample in one array (a light's RoundSize(nsamples))
struct SampleLayoutDev {
    int rowStride;   // u16 entries per row, a multiple of 8
    int n1D, n2D;
    uint16_t size1D[MAX_SAMPLE_ARRAYS], size2D[MAX_SAMPLE_ARRAYS];
    uint16_t row1D[MAX_SAMPLE_ARRAYS], row2D[MAX_SAMPLE_ARRAYS];
    uint16_t scr1D[MAX_SAMPLE_ARRAYS], scr2D[MAX_SAMPLE_ARRAYS];
};
// the same, array by array in LDPixelSample's draw order (montecarlo.cpp:228-236), for the general table kernel
struct TableArrays {
    int nArrays;
    int rowStride;
    int tgtPerSample;                       // bytes of shuffle-target scratch per camera sample
    uint16_t n[MAX_TABLE_ARRAYS];           // values per pixel sample
    uint16_t row[MAX_TABLE_ARRAYS];         // first row entry
    uint16_t scr[MAX_TABLE_ARRAYS];         // first scramble word
    uint16_t tgtInner[MAX_TABLE_ARRAYS];    // per-sample byte offset of the in-block shuffle targets (u8 each, n > 1 only)
    uint16_t tgtBlock[MAX_TABLE_ARRAYS];    // ... of the block shuffle targets (one u16)
    uint8_t dims[MAX_TABLE_ARRAYS];         // 1 or 2 scramble words
};

DEV uint32_t Temper(uint32_t y) {   // rng.cpp:101-105
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
DEV float UIntToFloat(uint32_t u) { return (u & 0xffffff) / float(1 << 24); }   // rng.cpp:61

DEV float VanDerCorput(uint32_t n, uint32_t scramble) {   // montecarlo.h:281-290
    n = __brev(n);
    n ^= scramble;
    return minf(((n >> 8) & 0xffffff) / float(1 << 24), ONE_MINUS_EPS);
}
// montecarlo.h:293-297 XORs v_i over the set bits i of n, v_0 = 1 << 31, v_(i+1) = v_i ^ (v_i >> 1): with a right shift read
// as multiplication by x, v_i = (1 + x)^i, so the sum is the polynomial n(y) over GF(2) evaluated at y = 1 + x -- a Taylor
// shift, five butterfly steps because (1 + x)^(2^k) = 1 + x^(2^k) -- written most significant bit first.  Same word for
// every n (tests/test_device_on_host.py::test_sobol2_closed_form checks it against the loop); no data-dependent trip count.
DEV uint32_t Sobol2Word(uint32_t n) {
    n ^= (n >> 1) & 0x55555555u;
    n ^= (n >> 2) & 0x33333333u;
    n ^= (n >> 4) & 0x0f0f0f0fu;
    n ^= (n >> 8) & 0x00ff00ffu;
    n ^= (n >> 16) & 0x0000ffffu;
    return __brev(n);
}
DEV float Sobol2(uint32_t n, uint32_t scramble) {
    scramble ^= Sobol2Word(n);
    return minf(((scramble >> 8) & 0xffffff) / float(1 << 24), ONE_MINUS_EPS);
}

// Warp-cooperative regeneration of one MT19937 block in shared memory
// (rng.cpp:76-96).  new[k] needs old[k], old[k+1] and old[k+397] (k < 227) or
// new[k-227] (k >= 227); batches of 32 consecutive k respect that dependency.
DEV void WarpRegenMT(uint32_t *mt, int lane) {
    for (int base = 0; base < MT_N; base += 32) {
        const int k = base + lane;
        uint32_t y = 0, src = 0;
        if (k < MT_N) {
            const uint32_t a = mt[k];
            const uint32_t b = mt[(k + 1 == MT_N) ? 0 : k + 1];   // k = 623 reads the NEW mt[0]
            y = (a & 0x80000000u) | (b & 0x7fffffffu);
            src = mt[(k < MT_N - MT_M) ? k + MT_M : k + (MT_M - MT_N)];
        }
        __syncwarp();
        if (k < MT_N) mt[k] = src ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        __syncwarp();
    }
}

// A task's two-block MT19937 window (cur = the block in use, next = the one after it) one block forward, by the lanes of
// one warp; `mt` is a block of shared memory of theirs.  (A host compiler sees the one-thread form: the CPU test stage
// runs the kernels one thread at a time, tests/native/device_on_host.cpp.)
DEV void WarpAdvanceWindow(uint32_t *cur, uint32_t *next, uint32_t *mt, int lane) {
#ifdef __CUDACC__
    for (int j = lane; j < MT_N; j += 32) {
        const uint32_t v = __ldcg(next + j);
        __stcg(cur + j, v);
        mt[j] = v;
    }
    __syncwarp();
    WarpRegenMT(mt, lane);
    for (int j = lane; j < MT_N; j += 32) __stcg(next + j, mt[j]);
    __syncwarp();
#else
    (void)lane;
    for (int j = 0; j < MT_N; ++j) { cur[j] = next[j]; mt[j] = next[j]; }
    for (int k = 0; k < MT_N; ++k) {
        const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % MT_N] & 0x7fffffffu);
        mt[k] = mt[(k + MT_M) % MT_N] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (int j = 0; j < MT_N; ++j) next[j] = mt[j];
#endif
}

// RNG::Seed, rng.cpp:43-55 (sequential by nature; one lane)
DEV void SeedMT(uint32_t *mt, uint32_t seed) {
    mt[0] = seed;
    for (int i = 1; i < MT_N; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
}

// draw index d in [0, 37 + 52*spp) of a pixel -> (array, role, i)
//   role 0: scramble word `i` of that array; 1: no-op draw; 2: shuffle draw for position i
DEV void ClassifyDraw(uint32_t d, uint32_t spp, int *array, int *role, uint32_t *i) {
    const uint32_t L1 = 1 + 2 * spp, L2 = 2 + 2 * spp;
    uint32_t a, off, ns;
    if (d < 2 * L2) { a = d / L2; off = d - a * L2; ns = 2; }
    else if (d < 2 * L2 + 15 * L1) { uint32_t e = d - 2 * L2; a = e / L1; off = e - a * L1; a += 2; ns = 1; }
    else { uint32_t e = d - 2 * L2 - 15 * L1; a = e / L2; off = e - a * L2; a += 17; ns = 2; }
    *array = (int)a;
    if (off < ns) { *role = 0; *i = off; }
    else if (off < ns + spp) { *role = 1; *i = off - ns; }
    else { *role = 2; *i = off - ns - spp; }
}
DEV uint32_t ArrayStart(int a, uint32_t spp) {   // first draw index of array a (a may be 26 = end)
    const uint32_t L1 = 1 + 2 * spp, L2 = 2 + 2 * spp;
    if (a <= 2) return a * L2;
    if (a <= 17) return 2 * L2 + (a - 2) * L1;
    return 2 * L2 + 15 * L1 + (a - 17) * L2;
}
DEV int ScrambleSlot(int a) {   // index of array a's first scramble word in the 37-word block
    if (a <= 2) return 2 * a;            // image 0,1  lens 2,3  time 4
    if (a <= 17) return 5 + (a - 3);     // oneD j -> 5 + j   (a = 3 + j)
    return 19 + 2 * (a - 17);            // twoD j -> 19 + 2j (a = 17 + j)
}

}  // namespace dev
