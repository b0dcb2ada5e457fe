// dloopsubdiv.cuh -- LoopSubdiv::Refine on the device.
//
// Replaces the body of LoopSubdiv::Refine (shapes/loopsubdiv.cpp:221-422): the subdivision levels, the limit-surface
// positions and normals, and the triangle list handed to TriangleMesh.  Its input is what the LoopSubdiv constructor
// leaves behind (loopsubdiv.cpp:156-216) -- the SDVertex / SDFace arrays with startFace, boundary / regular flags and
// the faces' neighbour links -- as index arrays; that topology pass stays in the host front end (it is a std::map walk
// over the CONTROL mesh, which is 4^nLevels times smaller than what comes out here).
//
// The result is the reference's mesh, vertex for vertex and triangle for triangle, because everything the reference
// orders by iteration is a function of indices:
//   * the children of face j are faces 4j .. 4j+3, the child of vertex j is vertex j (loopsubdiv.cpp:245-262);
//   * an edge's new vertex is created by the first (face, edge) pair that meets the edge in the reference's loop over
//     faces and edges (loopsubdiv.cpp:275-305) -- the pair itself if the edge is a boundary edge, else the
//     lexicographically smaller of the two pairs that share it -- and new vertices are numbered in that order: a prefix
//     sum over the 3 * nFaces creator flags.  The creating face also fixes the order of the two "opposite vertex" terms;
//   * one-ring sums are accumulated in SDVertex::oneRing's order (loopsubdiv.cpp:458-479), walked on the fly;
//   * float expressions are restated operation for operation (no FMA), sinf / cosf are the reference's libm (refmath.h).
//
// Like dbvhbuild.cuh the kernels need nothing but global loads / stores, so the CPU test stage runs this file with the
// launches replaced by loops (tests/native/loopsubdiv_on_host.cpp) against the host front end's LoopSubdiv::Refine.
#pragma once
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>
#include "dmath.cuh"
#include "../../../include/pbrtgpu.h"

#ifndef LSD_HOST_EMULATION
#define LSD_GLOBAL __global__
#else
#define LSD_GLOBAL static
#endif

namespace lsd {
using dev::V3;

enum { SCAN_CHUNK = 256 };
enum { VF_BOUNDARY = 1, VF_REGULAR = 2 };

struct Mesh {   // one subdivision level: SDVertex / SDFace (loopsubdiv.h:69-140) as arrays
    uint32_t nv, nf;
    float *P;          // [nv][3]
    int32_t *start;    // [nv] SDVertex::startFace, -1 for a vertex no face uses
    uint8_t *vflags;   // [nv] VF_*
    int32_t *fv;       // [nf][3] SDFace::v
    int32_t *ff;       // [nf][3] SDFace::f, -1 = no neighbour
};

#define LSD_NEXT(i) (((i) + 1) % 3)
#define LSD_PREV(i) (((i) + 2) % 3)
#define LSD_THREAD_ID (blockIdx.x * blockDim.x + threadIdx.x)
#define LSD_THREADS (gridDim.x * blockDim.x)

DEV V3 GetP(const Mesh &m, int v) { return V3(m.P[3 * (size_t)v], m.P[3 * (size_t)v + 1], m.P[3 * (size_t)v + 2]); }
DEV void SetP(float *P, size_t v, const V3 &p) { P[3 * v] = p.x; P[3 * v + 1] = p.y; P[3 * v + 2] = p.z; }
DEV int Vnum(const Mesh &m, int f, int vert) {   // SDFace::vnum, loopsubdiv.h:104-110
    for (int i = 0; i < 3; ++i) if (m.fv[3 * (size_t)f + i] == vert) return i;
    return 0;
}
DEV int NextFace(const Mesh &m, int f, int vert) { return m.ff[3 * (size_t)f + Vnum(m, f, vert)]; }
DEV int PrevFace(const Mesh &m, int f, int vert) { return m.ff[3 * (size_t)f + LSD_PREV(Vnum(m, f, vert))]; }
DEV int NextVert(const Mesh &m, int f, int vert) { return m.fv[3 * (size_t)f + LSD_NEXT(Vnum(m, f, vert))]; }
DEV int PrevVert(const Mesh &m, int f, int vert) { return m.fv[3 * (size_t)f + LSD_PREV(Vnum(m, f, vert))]; }
DEV int OtherVert(const Mesh &m, int f, int v0, int v1) {   // loopsubdiv.h:121-127
    for (int i = 0; i < 3; ++i) {
        const int v = m.fv[3 * (size_t)f + i];
        if (v != v0 && v != v1) return v;
    }
    return -1;
}
// (every walk around a vertex is capped at the face count: flags that do not match the links -- which the entry point
// cannot rule out without doing the walks itself -- then give a wrong mesh, never a kernel that does not return)
DEV int Valence(const Mesh &m, int vert) {   // SDVertex::valence, loopsubdiv.cpp:117-136
    const int start = m.start[vert];
    const int cap = (int)m.nf;
    int f = start;
    if (!(m.vflags[vert] & VF_BOUNDARY)) {
        int nf = 1;
        while ((f = NextFace(m, f, vert)) != start && f != -1 && nf <= cap) ++nf;
        return nf;
    }
    int nf = 1;
    while ((f = NextFace(m, f, vert)) != -1 && nf <= cap) ++nf;
    f = start;
    while ((f = PrevFace(m, f, vert)) != -1 && nf <= 2 * cap) ++nf;
    return nf + 1;
}
// SDVertex::oneRing (loopsubdiv.cpp:458-479) without the array: fn(k, P of the k-th ring vertex), in its order
template <typename F> DEV void ForRing(const Mesh &m, int vert, F fn) {
    int k = 0;
    const int cap = (int)m.nf;
    if (!(m.vflags[vert] & VF_BOUNDARY)) {
        const int start = m.start[vert];
        int face = start;
        do {
            fn(k++, GetP(m, NextVert(m, face, vert)));
            face = NextFace(m, face, vert);
        } while (face != start && face != -1 && k <= cap);
    } else {
        int face = m.start[vert], f2, guard = 0;
        while ((f2 = NextFace(m, face, vert)) != -1 && guard++ <= cap) face = f2;
        fn(k++, GetP(m, NextVert(m, face, vert)));
        do {
            fn(k++, GetP(m, PrevVert(m, face, vert)));
            face = PrevFace(m, face, vert);
        } while (face != -1 && k <= cap + 1);
    }
}
DEV V3 WeightOneRing(const Mesh &m, int vert, float beta) {   // loopsubdiv.cpp:447-456
    const int val = Valence(m, vert);
    V3 P = (1 - val * beta) * GetP(m, vert);
    ForRing(m, vert, [&](int, const V3 &r) { P = P + beta * r; });
    return P;
}
DEV V3 WeightBoundary(const Mesh &m, int vert, float beta) {   // loopsubdiv.cpp:481-490
    const int val = Valence(m, vert);
    V3 first(0.f, 0.f, 0.f), last(0.f, 0.f, 0.f);
    ForRing(m, vert, [&](int k, const V3 &r) { if (k == 0) first = r; if (k == val - 1) last = r; });
    V3 P = (1 - 2 * beta) * GetP(m, vert);
    P = P + beta * first;
    P = P + beta * last;
    return P;
}
DEV float LoopBeta(int valence) { return valence == 3 ? 3.f / 16.f : 3.f / (8.f * valence); }     // loopsubdiv.h:58-61
DEV float LoopGamma(int valence) { return 1.f / (valence + 3.f / (8.f * LoopBeta(valence))); }     // loopsubdiv.h:64-66

// the (face, edge) pair on the other side of edge k of face j, or face -1
DEV void Partner(const Mesh &m, int j, int k, int *f2, int *k2) {
    *f2 = m.ff[3 * (size_t)j + k];
    *k2 = 0;
    if (*f2 < 0) return;
    const int a = m.fv[3 * (size_t)j + k], b = m.fv[3 * (size_t)j + LSD_NEXT(k)];
    for (int e = 0; e < 3; ++e) {
        const int c = m.fv[3 * (size_t)*f2 + e], d = m.fv[3 * (size_t)*f2 + LSD_NEXT(e)];
        if ((c == a && d == b) || (c == b && d == a)) { *k2 = e; return; }
    }
}
DEV bool Creates(const Mesh &m, int j, int k) {   // does (j, k) meet its edge first in the loop of loopsubdiv.cpp:275-277?
    int f2, k2;
    Partner(m, j, k, &f2, &k2);
    return f2 < 0 || j < f2 || (j == f2 && k < k2);
}

// ---- one level: cur -> nxt ----------------------------------------------------------------------------------
// even vertices (loopsubdiv.cpp:249-272) and their face pointers (:308-312)
LSD_GLOBAL void k_even(Mesh cur, Mesh nxt) {
    for (uint32_t j = LSD_THREAD_ID; j < cur.nv; j += LSD_THREADS) {
        nxt.vflags[j] = cur.vflags[j];
        const int start = cur.start[j];
        if (start < 0) {   // a vertex no face refers to: carried along untouched
            SetP(nxt.P, j, GetP(cur, j));
            nxt.start[j] = -1;
            continue;
        }
        V3 P;
        if (!(cur.vflags[j] & VF_BOUNDARY)) {
            if (cur.vflags[j] & VF_REGULAR) P = WeightOneRing(cur, j, 1.f / 16.f);
            else P = WeightOneRing(cur, j, LoopBeta(Valence(cur, j)));
        } else {
            P = WeightBoundary(cur, j, 1.f / 8.f);
        }
        SetP(nxt.P, j, P);
        nxt.start[j] = 4 * start + Vnum(cur, start, j);
    }
}
LSD_GLOBAL void k_edge_flags(Mesh cur, uint8_t *flag, uint32_t nPadded) {
    for (uint32_t e = LSD_THREAD_ID; e < nPadded; e += LSD_THREADS)
        flag[e] = e < 3 * cur.nf ? (uint8_t)Creates(cur, (int)(e / 3), (int)(e % 3)) : 0;
}
// ---- exclusive prefix sum of 0/1 byte flags, in chunks (as in dbvhbuild.cuh)
LSD_GLOBAL void k_scan_sum_flags(const uint8_t *flag, uint32_t *sum1, uint32_t nChunks) {
    for (uint32_t c = LSD_THREAD_ID; c < nChunks; c += LSD_THREADS) {
        const uint4 *w = reinterpret_cast<const uint4 *>(flag + (size_t)c * SCAN_CHUNK);
        uint32_t sum = 0;
        for (int k = 0; k < SCAN_CHUNK / 16; ++k) {
            const uint4 v = w[k];
            sum += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
        }
        sum1[c] = sum;
    }
}
LSD_GLOBAL void k_scan_sum_sums(const uint32_t *sum1, uint32_t *sum2, uint32_t nChunks, uint32_t nSuper) {
    for (uint32_t c = LSD_THREAD_ID; c < nSuper; c += LSD_THREADS) {
        const uint32_t a = c * SCAN_CHUNK, b = (a + SCAN_CHUNK < nChunks) ? a + SCAN_CHUNK : nChunks;
        uint32_t sum = 0;
        for (uint32_t i = a; i < b; ++i) sum += sum1[i];
        sum2[c] = sum;
    }
}
LSD_GLOBAL void k_scan_top(uint32_t *sum2, uint32_t nSuper, uint32_t *total) {
    if (LSD_THREAD_ID != 0) return;
    uint32_t run = 0;
    for (uint32_t i = 0; i < nSuper; ++i) { const uint32_t v = sum2[i]; sum2[i] = run; run += v; }
    *total = run;
}
LSD_GLOBAL void k_scan_down_sums(uint32_t *sum1, const uint32_t *sum2, uint32_t nChunks, uint32_t nSuper) {
    for (uint32_t c = LSD_THREAD_ID; c < nSuper; c += LSD_THREADS) {
        const uint32_t a = c * SCAN_CHUNK, b = (a + SCAN_CHUNK < nChunks) ? a + SCAN_CHUNK : nChunks;
        uint32_t run = sum2[c];
        for (uint32_t i = a; i < b; ++i) { const uint32_t v = sum1[i]; sum1[i] = run; run += v; }
    }
}
LSD_GLOBAL void k_scan_down_flags(const uint8_t *flag, const uint32_t *sum1, uint32_t *rank, uint32_t nChunks) {
    for (uint32_t c = LSD_THREAD_ID; c < nChunks; c += LSD_THREADS) {
        uint32_t run = sum1[c];
        for (uint32_t i = c * SCAN_CHUNK; i < (c + 1) * SCAN_CHUNK; ++i) { rank[i] = run; run += flag[i]; }
    }
}
// odd vertices (loopsubdiv.cpp:275-305): created by the (face, edge) pairs flagged above, numbered by their rank
LSD_GLOBAL void k_odd(Mesh cur, Mesh nxt, const uint8_t *flag, const uint32_t *rank) {
    for (uint32_t e = LSD_THREAD_ID; e < 3 * cur.nf; e += LSD_THREADS) {
        if (!flag[e]) continue;
        const int j = (int)(e / 3), k = (int)(e % 3);
        const int a = cur.fv[3 * (size_t)j + k], b = cur.fv[3 * (size_t)j + LSD_NEXT(k)];
        const int v0 = a < b ? a : b, v1 = a < b ? b : a;   // SDEdge orders its two vertices (loopsubdiv.cpp:100-103)
        const int f2 = cur.ff[3 * (size_t)j + k];
        const uint32_t vert = cur.nv + rank[e];
        const V3 P0 = GetP(cur, v0), P1 = GetP(cur, v1);
        V3 P;
        if (f2 == -1) {
            P = 0.5f * P0;
            P = P + 0.5f * P1;
        } else {
            P = 3.f / 8.f * P0;
            P = P + 3.f / 8.f * P1;
            P = P + 1.f / 8.f * GetP(cur, OtherVert(cur, j, v0, v1));
            P = P + 1.f / 8.f * GetP(cur, OtherVert(cur, f2, v0, v1));
        }
        SetP(nxt.P, vert, P);
        nxt.vflags[vert] = (uint8_t)(VF_REGULAR | (f2 == -1 ? VF_BOUNDARY : 0));
        nxt.start[vert] = 4 * j + 3;
    }
}
DEV uint32_t EdgeVertex(const Mesh &cur, const uint8_t *flag, const uint32_t *rank, int j, int k) {
    if (flag[3 * (size_t)j + k]) return cur.nv + rank[3 * (size_t)j + k];
    int f2, k2;
    Partner(cur, j, k, &f2, &k2);
    return cur.nv + rank[3 * (size_t)f2 + k2];
}
// the four children of every face: neighbour links (loopsubdiv.cpp:314-331) and vertices (:333-344)
LSD_GLOBAL void k_faces(Mesh cur, Mesh nxt, const uint8_t *flag, const uint32_t *rank) {
    for (uint32_t j = LSD_THREAD_ID; j < cur.nf; j += LSD_THREADS) {
        const int c[4] = {(int)(4 * j), (int)(4 * j + 1), (int)(4 * j + 2), (int)(4 * j + 3)};
        for (int k = 0; k < 3; ++k) {
            const int vk = cur.fv[3 * (size_t)j + k];
            nxt.ff[3 * (size_t)c[3] + k] = c[LSD_NEXT(k)];
            nxt.ff[3 * (size_t)c[k] + LSD_NEXT(k)] = c[3];
            int f2 = cur.ff[3 * (size_t)j + k];
            nxt.ff[3 * (size_t)c[k] + k] = f2 != -1 ? 4 * f2 + Vnum(cur, f2, vk) : -1;
            f2 = cur.ff[3 * (size_t)j + LSD_PREV(k)];
            nxt.ff[3 * (size_t)c[k] + LSD_PREV(k)] = f2 != -1 ? 4 * f2 + Vnum(cur, f2, vk) : -1;
        }
        for (int k = 0; k < 3; ++k) {
            nxt.fv[3 * (size_t)c[k] + k] = cur.fv[3 * (size_t)j + k];   // the even vertex's child has its parent's index
            const int vert = (int)EdgeVertex(cur, flag, rank, (int)j, k);
            nxt.fv[3 * (size_t)c[k] + LSD_NEXT(k)] = vert;
            nxt.fv[3 * (size_t)c[LSD_NEXT(k)] + k] = vert;
            nxt.fv[3 * (size_t)c[3] + k] = vert;
        }
    }
}

// ---- after the last level ---------------------------------------------------------------------------------------
LSD_GLOBAL void k_limit(Mesh m, float *Plimit) {   // loopsubdiv.cpp:347-356
    for (uint32_t i = LSD_THREAD_ID; i < m.nv; i += LSD_THREADS) {
        V3 P;
        if (m.start[i] < 0) P = GetP(m, i);
        else if (m.vflags[i] & VF_BOUNDARY) P = WeightBoundary(m, i, 1.f / 5.f);
        else P = WeightOneRing(m, i, LoopGamma(Valence(m, i)));
        SetP(Plimit, i, P);
    }
}
// loopsubdiv.cpp:358-402; m.P already holds the limit positions
LSD_GLOBAL void k_normals(Mesh m, float *Ns) {
    for (uint32_t i = LSD_THREAD_ID; i < m.nv; i += LSD_THREADS) {
        V3 S(0.f, 0.f, 0.f), T(0.f, 0.f, 0.f);
        if (m.start[i] < 0) { SetP(Ns, i, V3(0.f, 0.f, 0.f)); continue; }
        const int valence = Valence(m, i);
        if (!(m.vflags[i] & VF_BOUNDARY)) {
            ForRing(m, i, [&](int k, const V3 &r) {
                S = S + dev::ref_cosf(2.f * dev::PI_F * k / valence) * r;
                T = T + dev::ref_sinf(2.f * dev::PI_F * k / valence) * r;
            });
        } else {
            V3 R[4] = {V3(0.f, 0.f, 0.f), V3(0.f, 0.f, 0.f), V3(0.f, 0.f, 0.f), V3(0.f, 0.f, 0.f)}, last(0.f, 0.f, 0.f);
            ForRing(m, i, [&](int k, const V3 &r) { if (k < 4) R[k] = r; if (k == valence - 1) last = r; });
            const V3 Pi = GetP(m, i);
            S = last - R[0];
            if (valence == 2) T = R[0] + R[1] - 2 * Pi;
            else if (valence == 3) T = R[1] - Pi;
            else if (valence == 4) T = -1 * R[0] + 2 * R[1] + 2 * R[2] + -1 * R[3] + -2 * Pi;
            else {
                const float theta = dev::PI_F / float(valence - 1);
                T = dev::ref_sinf(theta) * (R[0] + last);
                ForRing(m, i, [&](int k, const V3 &r) {
                    if (k >= 1 && k < valence - 1) {
                        const float wt = (2 * dev::ref_cosf(theta) - 2) * dev::ref_sinf((k)*theta);
                        T = T + wt * r;
                    }
                });
                T = -T;
            }
        }
        SetP(Ns, i, dev::Cross(S, T));
    }
}

// ---- driver -----------------------------------------------------------------------------------------------------
// LSD_LAUNCH(kernel, work, stream, args...), LSD_ALLOC / LSD_FREE, LSD_H2D / LSD_D2H (synchronous) come from the
// including file: loopsubdiv.cu (CUDA) or tests/native/loopsubdiv_on_host.cpp (loops and memcpy).
struct Buffers {
    std::vector<void *> all;
    template <typename T> bool Get(T **p, size_t count) {
        void *q = nullptr;
        if (!LSD_ALLOC(&q, (count ? count : 1) * sizeof(T))) return false;
        all.push_back(q);
        *p = static_cast<T *>(q);
        return true;
    }
    ~Buffers() { for (void *q : all) LSD_FREE(q); }
};

// sizes after nLevels levels (every level: F' = 4F, boundary edges B' = 2B, V' = V + E with E = (3F + B) / 2)
inline bool Counts(uint32_t nv, uint32_t nf, uint32_t nBoundaryEdges, int nLevels, uint64_t *nvOut, uint64_t *nfOut) {
    uint64_t V = nv, F = nf, B = nBoundaryEdges;
    for (int l = 0; l < nLevels; ++l) {
        if ((3 * F + B) & 1) return false;   // the neighbour links do not describe a manifold
        V += (3 * F + B) / 2;
        F *= 4;
        B *= 2;
    }
    *nvOut = V; *nfOut = F;
    return true;
}

// returns 0; 1 = allocation / copy / launch failure (the caller reports the CUDA error); 2 = inconsistent topology
inline int Refine(const PbrtGpuSubdivMesh &in, int nLevels, float *hostP, float *hostN, int32_t *hostIndices, void *stream) {
    uint32_t nB = 0;
    for (size_t i = 0; i < 3 * (size_t)in.nFaces; ++i) nB += in.faceF[i] < 0;
    uint64_t nvFinal = 0, nfFinal = 0;
    if (!Counts(in.nVertices, in.nFaces, nB, nLevels, &nvFinal, &nfFinal)) return 2;
    Buffers buf;
    Mesh m[2];
    for (int s = 0; s < 2; ++s) {
        // level l lives in m[l & 1]: the last level needs the final sizes, the one before it a quarter of the faces
        uint64_t V = in.nVertices, F = in.nFaces, B = nB, capV = V, capF = F;
        for (int l = 0; l < nLevels; ++l) {
            V += (3 * F + B) / 2; F *= 4; B *= 2;
            if (((l + 1) & 1) == s) { capV = V; capF = F; }
        }
        if (s == 1 && nLevels == 0) { capV = capF = 1; }
        memset(&m[s], 0, sizeof(Mesh));
        if (!buf.Get(&m[s].P, 3 * capV) || !buf.Get(&m[s].start, capV) || !buf.Get(&m[s].vflags, capV) ||
            !buf.Get(&m[s].fv, 3 * capF) || !buf.Get(&m[s].ff, 3 * capF))
            return 1;
    }
    const uint64_t maxEdgeSlots = nLevels > 0 ? 3 * (nfFinal / 4) : 1;
    const uint32_t maxChunks = (uint32_t)((maxEdgeSlots + SCAN_CHUNK - 1) / SCAN_CHUNK);
    const uint32_t maxSuper = (maxChunks + SCAN_CHUNK - 1) / SCAN_CHUNK;
    uint8_t *flag = nullptr;
    uint32_t *rank = nullptr, *sum1 = nullptr, *sum2 = nullptr, *total = nullptr;
    float *Plimit = nullptr, *Ns = nullptr;
    if (!buf.Get(&flag, (size_t)maxChunks * SCAN_CHUNK) || !buf.Get(&rank, (size_t)maxChunks * SCAN_CHUNK) || !buf.Get(&sum1, maxChunks) ||
        !buf.Get(&sum2, maxSuper) || !buf.Get(&total, 1) || !buf.Get(&Plimit, 3 * nvFinal) || !buf.Get(&Ns, 3 * nvFinal))
        return 1;
    m[0].nv = in.nVertices; m[0].nf = in.nFaces;
    if (!LSD_H2D(m[0].P, in.P, 3 * (size_t)in.nVertices * sizeof(float), stream) ||
        !LSD_H2D(m[0].start, in.startFace, (size_t)in.nVertices * sizeof(int32_t), stream) ||
        !LSD_H2D(m[0].vflags, in.vflags, (size_t)in.nVertices, stream) ||
        !LSD_H2D(m[0].fv, in.faceV, 3 * (size_t)in.nFaces * sizeof(int32_t), stream) ||
        !LSD_H2D(m[0].ff, in.faceF, 3 * (size_t)in.nFaces * sizeof(int32_t), stream))
        return 1;
    for (int l = 0; l < nLevels; ++l) {
        Mesh &cur = m[l & 1], &nxt = m[(l + 1) & 1];
        const uint32_t nSlots = 3 * cur.nf, nChunks = (nSlots + SCAN_CHUNK - 1) / SCAN_CHUNK, nSuper = (nChunks + SCAN_CHUNK - 1) / SCAN_CHUNK;
        LSD_LAUNCH(k_edge_flags, nChunks * SCAN_CHUNK, stream, cur, flag, nChunks * SCAN_CHUNK);
        LSD_LAUNCH(k_scan_sum_flags, nChunks, stream, flag, sum1, nChunks);
        LSD_LAUNCH(k_scan_sum_sums, nSuper, stream, sum1, sum2, nChunks, nSuper);
        LSD_LAUNCH(k_scan_top, 1, stream, sum2, nSuper, total);
        LSD_LAUNCH(k_scan_down_sums, nSuper, stream, sum1, sum2, nChunks, nSuper);
        LSD_LAUNCH(k_scan_down_flags, nChunks, stream, flag, sum1, rank, nChunks);
        uint32_t nEdges = 0;
        if (!LSD_D2H(&nEdges, total, 4, stream)) return 1;
        uint32_t nBl = 0;   // boundary edges of this level = nB * 2^l
        nBl = nB << l;
        if ((uint64_t)nEdges * 2 != 3 * (uint64_t)cur.nf + nBl) return 2;
        nxt.nv = cur.nv + nEdges;
        nxt.nf = 4 * cur.nf;
        LSD_LAUNCH(k_even, cur.nv, stream, cur, nxt);
        LSD_LAUNCH(k_odd, nSlots, stream, cur, nxt, flag, rank);
        LSD_LAUNCH(k_faces, cur.nf, stream, cur, nxt, flag, rank);
    }
    Mesh &fin = m[nLevels & 1];
    if (fin.nv != nvFinal || fin.nf != nfFinal) return 2;
    LSD_LAUNCH(k_limit, fin.nv, stream, fin, Plimit);
    Mesh lim = fin;
    lim.P = Plimit;
    LSD_LAUNCH(k_normals, fin.nv, stream, lim, Ns);
    if (!LSD_D2H(hostP, Plimit, 3 * (size_t)fin.nv * sizeof(float), stream) || !LSD_D2H(hostN, Ns, 3 * (size_t)fin.nv * sizeof(float), stream) ||
        !LSD_D2H(hostIndices, fin.fv, 3 * (size_t)fin.nf * sizeof(int32_t), stream))
        return 1;
    return 0;
}

}  // namespace lsd
