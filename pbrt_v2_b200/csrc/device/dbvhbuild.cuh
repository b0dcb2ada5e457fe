// dbvhbuild.cuh -- BVHAccel construction on the device.
//
// Replaces BVHAccel::BVHAccel's recursiveBuild + flattenBVHTree (accelerators/bvh.cpp:153-202, 210-380, 383-399)
// for SPLIT_SAH, and produces the SAME tree: the same nodes in the same depth-first order with the same
// primitive order in the leaves, so that the traversal kernels resolve every hit/miss and `t` tie as on a tree
// built by the reference (DESIGN.md "BVH construction").  What makes that possible without recursion:
//
//   * every decision of recursiveBuild depends only on the node's primitive RANGE [start, end) of the build
//     array: its bounds and centroid bounds (min / max: exact and order-free), the 12 SAH bucket counts and
//     bounds (ditto), and float expressions of those -- restated below operation for operation;
//   * std::partition (libstdc++, bidirectional/random-access version) always produces the same permutation:
//     with nT elements satisfying the predicate, the k-th element of [start, start+nT) that fails it is
//     swapped with the k-th element from the right of [start+nT, end) that satisfies it; everything else
//     stays.  Ranks come from a prefix sum of the predicate, so the whole level partitions in parallel;
//   * std::nth_element is only reached for ranges of <= 4 primitives (bvh.cpp:289-294), where libstdc++'s
//     introselect is one median-of-three partition step (4 elements) and an insertion sort -- restated for
//     one thread per node;
//   * recursiveBuild appends leaf ranges to orderedPrims left to right, so orderedPrims IS the final build
//     array, and a leaf's primitivesOffset is its `start`;
//   * flattenBVHTree numbers nodes in preorder: index = parent's + 1 (first child) or parent's + 1 + size of
//     the first child's subtree (second child); subtree sizes come from a bottom-up pass over the levels.
//
// The build is level-synchronous: all nodes of one depth are processed by a handful of kernels over the n
// build-array positions (position -> node map `nodeOf`).  The kernels use nothing but global loads/stores and
// integer atomics (float min/max through the order-preserving key below), so the CPU test stage runs this very
// file with the launches replaced by loops (tests/native/bvhbuild_on_host.cpp) and compares the result with
// the host front end's recursive build node for node.
//
// Signed zeros: the reference's Union() keeps the first of -0 / +0 it meets, the key order prefers -0 for a
// minimum and +0 for a maximum.  Bounds only ever enter comparisons and differences (SAH areas, slab tests), where
// the sign of a zero cannot change an outcome; node bounds may therefore differ from the host build in the
// SIGN BIT OF A ZERO coordinate and in nothing else (the tests compare floats by value).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../../include/pbrtgpu.h"

#ifndef BVHB_HOST_EMULATION
#define BVHB_GLOBAL __global__
#define BVHB_DEV __device__ __forceinline__
#else
#define BVHB_GLOBAL static
#define BVHB_DEV static inline
#endif

namespace bvhb {

enum { N_BUCKETS = 12 };                                   // bvh.cpp:297
enum { KIND_LEAF = 0, KIND_MIDSPLIT = 1, KIND_PARTITION = 2, KIND_SAH_PENDING = 3 };
enum { SCAN_CHUNK = 256 };

// order-preserving float -> uint key (inputs are finite: checked by pbrtgpu_bvh_build)
BVHB_DEV uint32_t FloatKey(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
BVHB_DEV float KeyFloat(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
#define BVHB_KEY_POS_INF 0xff800000u   // FloatKey(+inf): the empty box's pMin
#define BVHB_KEY_NEG_INF 0x007fffffu   // FloatKey(-inf): the empty box's pMax

struct Box { float lo[3], hi[3]; };
BVHB_DEV float SurfaceArea(const Box &b) {   // BBox::SurfaceArea, core/geometry.h:533-536
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return 2.f * (dx * dy + dx * dz + dy * dz);
}

// the build state, all device arrays.  Per position of the build array: idx, nodeOf, flag, prefix, pairR.
// Per node (ids in order of creation, level by level): the rest.
struct State {
    uint32_t n;
    uint32_t maxPrimsInNode;
    const float *bounds;      // [n][6] pMin xyz, pMax xyz (BVHPrimitiveInfo::bounds, bvh.cpp:40-49)
    uint32_t *idx;            // [n] BVHPrimitiveInfo::primitiveNumber, permuted in place like buildData
    int32_t *nodeOf;          // [n] node that owns the position on the current level, -1 once it is in a leaf
    uint8_t *flag;            // [n rounded up to SCAN_CHUNK] partition predicate, 0 in the padding
    uint32_t *prefix;         // [n + 1] exclusive prefix sum of flag
    uint32_t *sum1, *sum2;    // scan partials: per SCAN_CHUNK flags, per SCAN_CHUNK sum1 entries
    uint32_t *pairR;          // [n] position of the k-th misplaced element of the right part (from the right), at [start + k]
    // nodes
    uint32_t *start, *end, *mid, *left;   // left = id of the first child (second = left + 1)
    uint32_t *bbKey, *cbKey;  // [6] each: node bounds / centroid bounds as keys (min xyz, max xyz)
    uint8_t *kind, *dim;
    uint32_t *minCostSplit;
    uint32_t *size, *pre;     // subtree size, preorder index
    // SAH buckets of the current level's nodes that reach the bucket pass (more than 4 primitives: at most n / 5 per level),
    // by sahSlot[node] (handed out in k_decide; which node gets which slot does not matter)
    uint32_t *sahSlot;
    uint32_t *bCount;         // [N_BUCKETS]
    uint32_t *bKey;           // [N_BUCKETS][6]
    uint32_t *sahCount;       // [1] slots handed out on this level
    uint32_t *nodeCount;      // [1]
};

BVHB_DEV float Centroid(const float *b, int d) { return .5f * b[d] + .5f * b[3 + d]; }   // bvh.cpp:44-46
BVHB_DEV int BucketOf(float c, float cmin, float cmax) {   // bvh.cpp:305-308 and 350-356
    int b = N_BUCKETS * ((c - cmin) / (cmax - cmin));
    if (b == N_BUCKETS) b = N_BUCKETS - 1;
    return b;
}

#define BVHB_THREAD_ID (blockIdx.x * blockDim.x + threadIdx.x)
#define BVHB_THREADS (gridDim.x * blockDim.x)

// ---- per level ---------------------------------------------------------------------------------------
BVHB_GLOBAL void k_init_level(State s, uint32_t levelBase, uint32_t levelCount) {
    for (uint32_t t = BVHB_THREAD_ID; t < levelCount; t += BVHB_THREADS) {
        const uint32_t node = levelBase + t;
        for (int k = 0; k < 3; ++k) {
            s.bbKey[6 * (size_t)node + k] = s.cbKey[6 * (size_t)node + k] = BVHB_KEY_POS_INF;
            s.bbKey[6 * (size_t)node + 3 + k] = s.cbKey[6 * (size_t)node + 3 + k] = BVHB_KEY_NEG_INF;
        }
        if (t == 0) *s.sahCount = 0;
    }
}

// bounds and centroid bounds of every node of the level (bvh.cpp:216-218, 232-234).  Positions are visited with a
// grid stride; a thread folds consecutive visits to the same node locally, so the few huge nodes of the top levels
// cost one set of atomics per thread instead of one per primitive.
BVHB_GLOBAL void k_reduce(State s) {
    int32_t cur = -1;
    float bl[3], bh[3], cl[3], ch[3];
    for (uint32_t i = BVHB_THREAD_ID;; i += BVHB_THREADS) {
        const int32_t node = i < s.n ? s.nodeOf[i] : -1;
        if (node != cur) {
            if (cur >= 0)
                for (int k = 0; k < 3; ++k) {
                    atomicMin(&s.bbKey[6 * (size_t)cur + k], FloatKey(bl[k]));
                    atomicMax(&s.bbKey[6 * (size_t)cur + 3 + k], FloatKey(bh[k]));
                    atomicMin(&s.cbKey[6 * (size_t)cur + k], FloatKey(cl[k]));
                    atomicMax(&s.cbKey[6 * (size_t)cur + 3 + k], FloatKey(ch[k]));
                }
            cur = node;
            for (int k = 0; k < 3; ++k) {
                bl[k] = cl[k] = __uint_as_float(0x7f800000u);
                bh[k] = ch[k] = __uint_as_float(0xff800000u);
            }
        }
        if (i >= s.n) break;
        if (node < 0) continue;
        const float *b = s.bounds + (size_t)s.idx[i] * 6;
        for (int k = 0; k < 3; ++k) {
            const float c = Centroid(b, k);
            bl[k] = b[k] < bl[k] ? b[k] : bl[k];
            bh[k] = b[3 + k] > bh[k] ? b[3 + k] : bh[k];
            cl[k] = c < cl[k] ? c : cl[k];
            ch[k] = c > ch[k] ? c : ch[k];
        }
    }
}

// std::nth_element(first, first + nth, first + n, centroid[dim] <) of libstdc++ for n <= 4 (n > 3: one introselect
// round = median-of-three to the front + unguarded partition, then the insertion sort of what is left)
BVHB_DEV void SmallNthElement(const State &s, uint32_t *a, int n, int nth, int dim) {
    float key[4];
    uint32_t val[4];
    for (int i = 0; i < n; ++i) { val[i] = a[i]; key[i] = Centroid(s.bounds + (size_t)a[i] * 6, dim); }
#define BVHB_SWAP(i, j) do { const float tk = key[i]; key[i] = key[j]; key[j] = tk; const uint32_t tv = val[i]; val[i] = val[j]; val[j] = tv; } while (0)
    int first = 0, last = n;
    int depthLimit = 0;
    for (int m = n; m > 1; m >>= 1) ++depthLimit;
    depthLimit *= 2;
    while (last - first > 3) {
        // (depthLimit cannot reach 0 with n <= 4: the heap-select fallback is never taken)
        --depthLimit;
        const int midp = first + (last - first) / 2;
        const int ia = first + 1, ib = midp, ic = last - 1;   // __move_median_to_first(first, first+1, mid, last-1)
        if (key[ia] < key[ib]) {
            if (key[ib] < key[ic]) BVHB_SWAP(first, ib);
            else if (key[ia] < key[ic]) BVHB_SWAP(first, ic);
            else BVHB_SWAP(first, ia);
        } else if (key[ia] < key[ic]) BVHB_SWAP(first, ia);
        else if (key[ib] < key[ic]) BVHB_SWAP(first, ic);
        else BVHB_SWAP(first, ib);
        int lo = first + 1, hi = last;                        // __unguarded_partition(first+1, last, pivot = *first)
        for (;;) {
            while (key[lo] < key[first]) ++lo;
            --hi;
            while (key[first] < key[hi]) --hi;
            if (!(lo < hi)) break;
            BVHB_SWAP(lo, hi);
            ++lo;
        }
        if (lo <= nth) first = lo; else last = lo;
    }
    for (int i = first + 1; i < last; ++i) {                  // __insertion_sort(first, last)
        const float k = key[i];
        const uint32_t v = val[i];
        int j = i;
        if (k < key[first]) {
            for (; j > first; --j) { key[j] = key[j - 1]; val[j] = val[j - 1]; }
        } else {
            for (; k < key[j - 1]; --j) { key[j] = key[j - 1]; val[j] = val[j - 1]; }
        }
        key[j] = k; val[j] = v;
    }
#undef BVHB_SWAP
    for (int i = 0; i < n; ++i) a[i] = val[i];
}

// recursiveBuild's case analysis up to the SAH bucket pass (bvh.cpp:219-294)
BVHB_GLOBAL void k_decide(State s, uint32_t levelBase, uint32_t levelCount) {
    for (uint32_t t = BVHB_THREAD_ID; t < levelCount; t += BVHB_THREADS) {
        const uint32_t node = levelBase + t;
        const uint32_t st = s.start[node], en = s.end[node], np = en - st;
        s.left[node] = 0;
        s.dim[node] = 0;
        s.mid[node] = (st + en) / 2;
        if (np <= 1) { s.kind[node] = KIND_LEAF; continue; }   // (np == 0 cannot arise: every split leaves both sides non-empty)
        float d[3];
        for (int k = 0; k < 3; ++k) d[k] = KeyFloat(s.cbKey[6 * (size_t)node + 3 + k]) - KeyFloat(s.cbKey[6 * (size_t)node + k]);
        const int dim = (d[0] > d[1] && d[0] > d[2]) ? 0 : (d[1] > d[2] ? 1 : 2);   // BBox::MaximumExtent
        s.dim[node] = (uint8_t)dim;
        if (s.cbKey[6 * (size_t)node + 3 + dim] == s.cbKey[6 * (size_t)node + dim] ||
            KeyFloat(s.cbKey[6 * (size_t)node + 3 + dim]) == KeyFloat(s.cbKey[6 * (size_t)node + dim])) {   // bvh.cpp:239-261 (-0 == +0 too)
            s.kind[node] = np <= s.maxPrimsInNode ? KIND_LEAF : KIND_MIDSPLIT;
            if (s.kind[node] == KIND_LEAF) s.dim[node] = 0;
            continue;
        }
        if (np <= 4) {   // bvh.cpp:289-294: equal counts
            SmallNthElement(s, s.idx + st, (int)np, (int)(s.mid[node] - st), dim);
            s.kind[node] = KIND_MIDSPLIT;
            continue;
        }
        s.kind[node] = KIND_SAH_PENDING;
        const uint32_t slot = atomicAdd(s.sahCount, 1u);
        s.sahSlot[node] = slot;
        for (int b = 0; b < N_BUCKETS; ++b) {
            s.bCount[(size_t)slot * N_BUCKETS + b] = 0;
            for (int k = 0; k < 3; ++k) {
                s.bKey[((size_t)slot * N_BUCKETS + b) * 6 + k] = BVHB_KEY_POS_INF;
                s.bKey[((size_t)slot * N_BUCKETS + b) * 6 + 3 + k] = BVHB_KEY_NEG_INF;
            }
        }
    }
}

// bvh.cpp:296-311: bucket counts and bounds
BVHB_GLOBAL void k_bucket(State s) {
    for (uint32_t i = BVHB_THREAD_ID; i < s.n; i += BVHB_THREADS) {
        const int32_t node = s.nodeOf[i];
        if (node < 0 || s.kind[node] != KIND_SAH_PENDING) continue;
        const int dim = s.dim[node];
        const float cmin = KeyFloat(s.cbKey[6 * (size_t)node + dim]), cmax = KeyFloat(s.cbKey[6 * (size_t)node + 3 + dim]);
        const float *b = s.bounds + (size_t)s.idx[i] * 6;
        const int bk = BucketOf(Centroid(b, dim), cmin, cmax);
        const size_t slot = (size_t)s.sahSlot[node] * N_BUCKETS + bk;
        atomicAdd(&s.bCount[slot], 1u);
        for (int k = 0; k < 3; ++k) {
            atomicMin(&s.bKey[slot * 6 + k], FloatKey(b[k]));
            atomicMax(&s.bKey[slot * 6 + 3 + k], FloatKey(b[3 + k]));
        }
    }
}

// bvh.cpp:313-361: cost of the 11 splits, the cheapest one, leaf or partition
BVHB_GLOBAL void k_sah(State s, uint32_t levelBase, uint32_t levelCount) {
    for (uint32_t t = BVHB_THREAD_ID; t < levelCount; t += BVHB_THREADS) {
        const uint32_t node = levelBase + t;
        if (s.kind[node] != KIND_SAH_PENDING) continue;
        const uint32_t np = s.end[node] - s.start[node];
        Box bbox;
        for (int k = 0; k < 3; ++k) { bbox.lo[k] = KeyFloat(s.bbKey[6 * (size_t)node + k]); bbox.hi[k] = KeyFloat(s.bbKey[6 * (size_t)node + 3 + k]); }
        const float invArea = SurfaceArea(bbox);
        float minCost = 0.f;
        uint32_t minCostSplit = 0;
        for (int i = 0; i < N_BUCKETS - 1; ++i) {
            Box b0, b1;
            for (int k = 0; k < 3; ++k) {
                b0.lo[k] = b1.lo[k] = __uint_as_float(0x7f800000u);
                b0.hi[k] = b1.hi[k] = __uint_as_float(0xff800000u);
            }
            int count0 = 0, count1 = 0;
            for (int j = 0; j < N_BUCKETS; ++j) {
                const size_t slot = (size_t)s.sahSlot[node] * N_BUCKETS + j;
                Box &dst = j <= i ? b0 : b1;
                (j <= i ? count0 : count1) += (int)s.bCount[slot];
                for (int k = 0; k < 3; ++k) {
                    const float lo = KeyFloat(s.bKey[slot * 6 + k]), hi = KeyFloat(s.bKey[slot * 6 + 3 + k]);
                    dst.lo[k] = lo < dst.lo[k] ? lo : dst.lo[k];
                    dst.hi[k] = hi > dst.hi[k] ? hi : dst.hi[k];
                }
            }
            // an empty side gives 0 * inf = NaN as in the reference (never the minimum: `<` is false for it)
            const float cost = .125f + (count0 * SurfaceArea(b0) + count1 * SurfaceArea(b1)) / invArea;
            if (i == 0 || cost < minCost) { minCost = cost; minCostSplit = (uint32_t)i; }
        }
        if (np > s.maxPrimsInNode || minCost < np) {
            s.kind[node] = KIND_PARTITION;
            s.minCostSplit[node] = minCostSplit;
        } else {
            s.kind[node] = KIND_LEAF;
            s.dim[node] = 0;
        }
    }
}

// the partition predicate of every position (CompareToBucket, bvh.cpp:126-145); 0 outside partitioned nodes
BVHB_GLOBAL void k_flags(State s, uint32_t nPadded) {
    for (uint32_t i = BVHB_THREAD_ID; i < nPadded; i += BVHB_THREADS) {   // (the padding up to a whole scan chunk stays 0)
        const int32_t node = i < s.n ? s.nodeOf[i] : -1;
        uint8_t f = 0;
        if (node >= 0 && s.kind[node] == KIND_PARTITION) {
            const int dim = s.dim[node];
            const float cmin = KeyFloat(s.cbKey[6 * (size_t)node + dim]), cmax = KeyFloat(s.cbKey[6 * (size_t)node + 3 + dim]);
            f = BucketOf(Centroid(s.bounds + (size_t)s.idx[i] * 6, dim), cmin, cmax) <= (int)s.minCostSplit[node];
        }
        s.flag[i] = f;
    }
}

// ---- exclusive prefix sum of flag[0..n) into prefix[0..n], in chunks (no intra-block cooperation needed) ----
BVHB_GLOBAL void k_scan_sum_flags(State s, uint32_t nChunks) {
    for (uint32_t c = BVHB_THREAD_ID; c < nChunks; c += BVHB_THREADS) {
        const uint4 *w = reinterpret_cast<const uint4 *>(s.flag + (size_t)c * SCAN_CHUNK);   // flags are 0 / 1 bytes
        uint32_t sum = 0;
        for (int k = 0; k < SCAN_CHUNK / 16; ++k) {
            const uint4 v = w[k];
            sum += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
        }
        s.sum1[c] = sum;
    }
}
BVHB_GLOBAL void k_scan_sum_sums(State s, uint32_t nChunks, uint32_t nSuper) {
    for (uint32_t c = BVHB_THREAD_ID; c < nSuper; c += BVHB_THREADS) {
        const uint32_t a = c * SCAN_CHUNK, b = (a + SCAN_CHUNK < nChunks) ? a + SCAN_CHUNK : nChunks;
        uint32_t sum = 0;
        for (uint32_t i = a; i < b; ++i) sum += s.sum1[i];
        s.sum2[c] = sum;
    }
}
BVHB_GLOBAL void k_scan_top(State s, uint32_t nSuper) {
    if (BVHB_THREAD_ID != 0) return;
    uint32_t run = 0;
    for (uint32_t i = 0; i < nSuper; ++i) { const uint32_t v = s.sum2[i]; s.sum2[i] = run; run += v; }
}
BVHB_GLOBAL void k_scan_down_sums(State s, uint32_t nChunks, uint32_t nSuper) {
    for (uint32_t c = BVHB_THREAD_ID; c < nSuper; c += BVHB_THREADS) {
        const uint32_t a = c * SCAN_CHUNK, b = (a + SCAN_CHUNK < nChunks) ? a + SCAN_CHUNK : nChunks;
        uint32_t run = s.sum2[c];
        for (uint32_t i = a; i < b; ++i) { const uint32_t v = s.sum1[i]; s.sum1[i] = run; run += v; }
    }
}
BVHB_GLOBAL void k_scan_down_flags(State s, uint32_t nChunks) {
    for (uint32_t c = BVHB_THREAD_ID; c < nChunks; c += BVHB_THREADS) {
        const uint32_t a = c * SCAN_CHUNK, b = (a + SCAN_CHUNK < s.n) ? a + SCAN_CHUNK : s.n;
        uint32_t run = s.sum1[c];
        for (uint32_t i = a; i < b; ++i) { s.prefix[i] = run; run += s.flag[i]; }
        if (b == s.n) s.prefix[s.n] = run;
    }
}

// mid of the partitioned nodes (std::partition's return value, bvh.cpp:357-360)
BVHB_GLOBAL void k_mid(State s, uint32_t levelBase, uint32_t levelCount) {
    for (uint32_t t = BVHB_THREAD_ID; t < levelCount; t += BVHB_THREADS) {
        const uint32_t node = levelBase + t;
        if (s.kind[node] == KIND_PARTITION) s.mid[node] = s.start[node] + (s.prefix[s.end[node]] - s.prefix[s.start[node]]);
    }
}
// the k-th misplaced element of the left part (a `false` before mid) is swapped with the k-th misplaced element of the
// right part counted from the right (a `true` at or after mid): the latter leave their positions at slot start + k ...
BVHB_GLOBAL void k_pairs(State s) {
    for (uint32_t i = BVHB_THREAD_ID; i < s.n; i += BVHB_THREADS) {
        const int32_t node = s.nodeOf[i];
        if (node < 0 || s.kind[node] != KIND_PARTITION) continue;
        const uint32_t st = s.start[node], en = s.end[node], mid = s.mid[node];
        if (i >= mid && s.flag[i]) s.pairR[st + (s.prefix[en] - s.prefix[i + 1])] = i;
    }
}
// ... where the former pick them up
BVHB_GLOBAL void k_swap(State s) {
    for (uint32_t i = BVHB_THREAD_ID; i < s.n; i += BVHB_THREADS) {
        const int32_t node = s.nodeOf[i];
        if (node < 0 || s.kind[node] != KIND_PARTITION) continue;
        const uint32_t st = s.start[node];
        if (i < s.mid[node] && !s.flag[i]) {
            const uint32_t j = s.pairR[st + ((i - st) - (s.prefix[i] - s.prefix[st]))];
            const uint32_t a = s.idx[i];
            s.idx[i] = s.idx[j];
            s.idx[j] = a;
        }
    }
}
// two children per interior node, created in pairs (ids need not be deterministic: the output order is the preorder)
BVHB_GLOBAL void k_children(State s, uint32_t levelBase, uint32_t levelCount) {
    for (uint32_t t = BVHB_THREAD_ID; t < levelCount; t += BVHB_THREADS) {
        const uint32_t node = levelBase + t;
        if (s.kind[node] == KIND_LEAF) continue;
        const uint32_t c = atomicAdd(s.nodeCount, 2u);
        s.left[node] = c;
        s.start[c] = s.start[node]; s.end[c] = s.mid[node];
        s.start[c + 1] = s.mid[node]; s.end[c + 1] = s.end[node];
    }
}
BVHB_GLOBAL void k_assign(State s) {
    for (uint32_t i = BVHB_THREAD_ID; i < s.n; i += BVHB_THREADS) {
        const int32_t node = s.nodeOf[i];
        if (node < 0) continue;
        s.nodeOf[i] = s.kind[node] == KIND_LEAF ? -1 : (int32_t)(s.left[node] + (i >= s.mid[node] ? 1u : 0u));
    }
}

// ---- flattenBVHTree (bvh.cpp:383-399) ------------------------------------------------------------------
BVHB_GLOBAL void k_sizes(State s, uint32_t levelBase, uint32_t levelCount) {
    for (uint32_t t = BVHB_THREAD_ID; t < levelCount; t += BVHB_THREADS) {
        const uint32_t node = levelBase + t;
        s.size[node] = s.kind[node] == KIND_LEAF ? 1u : 1u + s.size[s.left[node]] + s.size[s.left[node] + 1];
    }
}
BVHB_GLOBAL void k_preorder(State s, uint32_t levelBase, uint32_t levelCount) {
    for (uint32_t t = BVHB_THREAD_ID; t < levelCount; t += BVHB_THREADS) {
        const uint32_t node = levelBase + t;
        if (node == 0) s.pre[0] = 0;
        if (s.kind[node] == KIND_LEAF) continue;
        s.pre[s.left[node]] = s.pre[node] + 1;
        s.pre[s.left[node] + 1] = s.pre[node] + 1 + s.size[s.left[node]];
    }
}
BVHB_GLOBAL void k_emit(State s, uint32_t nNodes, PbrtGpuBVHNode *out) {
    for (uint32_t node = BVHB_THREAD_ID; node < nNodes; node += BVHB_THREADS) {
        PbrtGpuBVHNode o;
        for (int k = 0; k < 3; ++k) { o.bmin[k] = KeyFloat(s.bbKey[6 * (size_t)node + k]); o.bmax[k] = KeyFloat(s.bbKey[6 * (size_t)node + 3 + k]); }
        if (s.kind[node] == KIND_LEAF) {
            o.offset = s.start[node];                                   // primitivesOffset
            o.nPrimitives = (uint8_t)(s.end[node] - s.start[node]);
            o.axis = 0;
        } else {
            o.offset = s.pre[s.left[node] + 1];                         // secondChildOffset
            o.nPrimitives = 0;
            o.axis = s.dim[node];
        }
        o.pad[0] = o.pad[1] = 0;
        out[s.pre[node]] = o;
    }
}
BVHB_GLOBAL void k_iota(State s) {
    for (uint32_t i = BVHB_THREAD_ID; i < s.n; i += BVHB_THREADS) { s.idx[i] = i; s.nodeOf[i] = 0; }
}

// ---- driver ----------------------------------------------------------------------------------------------
// BVHB_LAUNCH(kernel, work, stream, args...), BVHB_ALLOC / BVHB_FREE, BVHB_H2D / BVHB_D2H (synchronous) come from
// the including file: bvhbuild.cu (CUDA) or tests/native/bvhbuild_on_host.cpp (loops and memcpy).
struct Buffers {
    std::vector<void *> all;
    template <typename T> bool Get(T **p, size_t count) {
        void *q = nullptr;
        if (!BVHB_ALLOC(&q, (count ? count : 1) * sizeof(T))) return false;
        all.push_back(q);
        *p = static_cast<T *>(q);
        return true;
    }
    ~Buffers() { for (void *q : all) BVHB_FREE(q); }
};

// returns 0, or 1 for a failed allocation / copy / launch (the caller reports the CUDA error)
inline int Build(const float *hostBounds, uint32_t n, uint32_t maxPrimsInNode, PbrtGpuBVHNode *hostNodes, uint32_t *nNodesOut,
                 uint32_t *hostOrder, uint32_t *nLevelsOut, void *stream) {
    *nNodesOut = 0;
    if (nLevelsOut) *nLevelsOut = 0;
    if (n == 0) return 0;
    Buffers buf;
    State s;
    memset(&s, 0, sizeof(s));
    s.n = n; s.maxPrimsInNode = maxPrimsInNode;
    const size_t maxNodes = 2 * (size_t)n;
    const uint32_t nChunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK, nSuper = (nChunks + SCAN_CHUNK - 1) / SCAN_CHUNK;
    const size_t maxBucketNodes = (size_t)n / 5 + 1;   // nodes of one level with more than 4 primitives
    float *dBounds = nullptr;
    PbrtGpuBVHNode *dOut = nullptr;
    bool ok = buf.Get(&dBounds, (size_t)n * 6) && buf.Get(&s.idx, n) && buf.Get(&s.nodeOf, n) && buf.Get(&s.flag, (size_t)nChunks * SCAN_CHUNK) &&
              buf.Get(&s.prefix, (size_t)n + 1) && buf.Get(&s.sum1, nChunks) && buf.Get(&s.sum2, nSuper) && buf.Get(&s.pairR, n) &&
              buf.Get(&s.start, maxNodes) && buf.Get(&s.end, maxNodes) && buf.Get(&s.mid, maxNodes) &&
              buf.Get(&s.left, maxNodes) && buf.Get(&s.bbKey, maxNodes * 6) && buf.Get(&s.cbKey, maxNodes * 6) &&
              buf.Get(&s.kind, maxNodes) && buf.Get(&s.dim, maxNodes) && buf.Get(&s.minCostSplit, maxNodes) &&
              buf.Get(&s.size, maxNodes) && buf.Get(&s.pre, maxNodes) && buf.Get(&s.bCount, maxBucketNodes * N_BUCKETS) &&
              buf.Get(&s.bKey, maxBucketNodes * N_BUCKETS * 6) && buf.Get(&s.sahSlot, maxNodes) && buf.Get(&s.sahCount, 1) &&
              buf.Get(&s.nodeCount, 1) && buf.Get(&dOut, maxNodes);
    if (!ok) return 1;
    s.bounds = dBounds;
    if (!BVHB_H2D(dBounds, hostBounds, (size_t)n * 6 * sizeof(float), stream)) return 1;
    const uint32_t rootRange[2] = {0u, n}, one = 1;
    if (!BVHB_H2D(s.start, &rootRange[0], 4, stream) || !BVHB_H2D(s.end, &rootRange[1], 4, stream) ||
        !BVHB_H2D(s.nodeCount, &one, 4, stream))
        return 1;
    BVHB_LAUNCH(k_iota, n, stream, s);

    std::vector<uint32_t> levelBase(1, 0u), levelCount(1, 1u);
    for (;;) {
        const uint32_t base = levelBase.back(), count = levelCount.back();
        BVHB_LAUNCH(k_init_level, count, stream, s, base, count);
        BVHB_LAUNCH(k_reduce, n, stream, s);
        BVHB_LAUNCH(k_decide, count, stream, s, base, count);
        BVHB_LAUNCH(k_bucket, n, stream, s);
        BVHB_LAUNCH(k_sah, count, stream, s, base, count);
        BVHB_LAUNCH(k_flags, nChunks * SCAN_CHUNK, stream, s, nChunks * SCAN_CHUNK);
        BVHB_LAUNCH(k_scan_sum_flags, nChunks, stream, s, nChunks);
        BVHB_LAUNCH(k_scan_sum_sums, nSuper, stream, s, nChunks, nSuper);
        BVHB_LAUNCH(k_scan_top, 1, stream, s, nSuper);
        BVHB_LAUNCH(k_scan_down_sums, nSuper, stream, s, nChunks, nSuper);
        BVHB_LAUNCH(k_scan_down_flags, nChunks, stream, s, nChunks);
        BVHB_LAUNCH(k_mid, count, stream, s, base, count);
        BVHB_LAUNCH(k_pairs, n, stream, s);
        BVHB_LAUNCH(k_swap, n, stream, s);
        BVHB_LAUNCH(k_children, count, stream, s, base, count);
        BVHB_LAUNCH(k_assign, n, stream, s);
        uint32_t total = 0;
        if (!BVHB_D2H(&total, s.nodeCount, 4, stream)) return 1;
        const uint32_t nextBase = base + count;
        if (total == nextBase) break;
        if (levelBase.size() > (size_t)n + 1) return 1;   // deeper than a chain of n nodes: cannot be a tree
        levelBase.push_back(nextBase);
        levelCount.push_back(total - nextBase);
    }
    const uint32_t nNodes = levelBase.back() + levelCount.back();
    for (size_t l = levelBase.size(); l-- > 0;) BVHB_LAUNCH(k_sizes, levelCount[l], stream, s, levelBase[l], levelCount[l]);
    for (size_t l = 0; l < levelBase.size(); ++l) BVHB_LAUNCH(k_preorder, levelCount[l], stream, s, levelBase[l], levelCount[l]);
    BVHB_LAUNCH(k_emit, nNodes, stream, s, nNodes, dOut);
    if (!BVHB_D2H(hostNodes, dOut, (size_t)nNodes * sizeof(PbrtGpuBVHNode), stream)) return 1;
    if (!BVHB_D2H(hostOrder, s.idx, (size_t)n * sizeof(uint32_t), stream)) return 1;
    *nNodesOut = nNodes;
    if (nLevelsOut) *nLevelsOut = (uint32_t)levelBase.size();
    return 0;
}

}  // namespace bvhb
