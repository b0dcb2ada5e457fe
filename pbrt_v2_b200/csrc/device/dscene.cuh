// dscene.cuh -- device-resident scene layout and the ray/primitive kernels'
// building blocks: BVHAccel::Intersect / IntersectP traversal and the
// Triangle / Sphere / Disk intersection tests.
//
// HBM layout (all read-only, loaded through the non-coherent path):
//   nodes   2 x float4 per node  = the reference's 32-byte LinearBVHNode
//           (accelerators/bvh.cpp:113-123): {bmin.xyz, bmax.x} {bmax.yz, offset, nPrims|axis<<8}
//           -> two 128-bit loads per node visit.
//   prims   3 x float4 per primitive in BVH leaf order = 48 bytes, the same
//           "12 B indices + 36 B positions" the reference touches per triangle
//           test (trianglemesh.cpp:133-135) but pre-gathered, so a leaf test is
//           three coalescable 128-bit loads instead of index -> vertex chasing:
//           {p1.xyz, type | (owner+1)<<8} {p2.xyz, index} {p3.xyz, material | (areaLight+1)<<16}
//           (owner = the TransformedPrimitive a nested primitive belongs to, -1 for the scene BVH's own)
//   everything else (normals, uvs, quadrics, materials, kd-tree, lights) is
//   touched only when a hit is shaded and stays in the PbrtGpu* POD form.
#pragma once
#include "dmath.cuh"
#include "danim.cuh"
#include "../../../include/pbrtgpu.h"
#include <cstring>
#include <vector>

// Round-2 experiment, OFF in the shipped kernels (the default build is textually unchanged): sibling-pair node layout.
// The reference's walk loads one 32-byte node per visit and learns only then whether the ray misses it; with the two
// children of every interior node side by side in one 64-byte record, one load serves the near child's test (which the
// reference performs next anyway) and the far child's.  The far child's test is the same arithmetic done earlier: its
// outcome can only change through the final `tmin < ray.maxt` (maxt shrinks), so a far child that fails now is never
// pushed, and one that passes is pushed WITH its tmin and re-compared against the then-current maxt when popped --
// exactly the decision BVHAccel::Intersect takes at that moment (bvh.cpp:126-148, 403-454).  Same nodes entered, same
// primitives tested in the same order, same result; fewer dependent loads.  Verified on the CPU
// (tests/test_device_on_host.py builds the harness with -DPBRT_PAIR_NODES=1 too); the counting (STATS) instantiations
// always take the reference-shaped walk, so the algorithmic byte counts do not change.
#ifndef PBRT_PAIR_NODES
#define PBRT_PAIR_NODES 0
#endif
// The measured-BRDF kd-tree as ONE 32-byte record per node ({splitPos, bits, p.xy} {p.z, v.rgb}: two 128-bit loads)
// instead of a node array and a sample array of 4-byte-aligned structs (eight 32-bit loads per visit).  Same values,
// same visiting order.  On since round 2 (BENCH_r01 extras: +2.3 % on C2, film unchanged on the pixels no other task
// splats into); -DPBRT_KD_INTERLEAVED=0 reads the two arrays.
#ifndef PBRT_KD_INTERLEAVED
#define PBRT_KD_INTERLEAVED 1
#endif

// Top of the scene BVH in shared memory (north_star: "the top of the BVH staged in shared memory via TMA").  Every ray
// starts at the root, so the first levels are the hottest lines of the node array; a persistent kernel copies them once
// per CTA (one cp.async.bulk = SASS UBLKCP, completion on an mbarrier) and the walk reads them with LDS instead of
// LDG while its heap index stays inside the table.  0 = off (every node from global memory / L1).
#ifndef PBRT_TOP_LEVELS
#define PBRT_TOP_LEVELS 0
#endif
// Traversal stack in shared memory (north_star: "short per-thread traversal stacks in registers/shared"): the first
// PBRT_SMEM_STACK entries of the walk's todo[] live in the CTA's shared memory, entry-major (entry i of thread t at
// [i * 128 + t]: conflict-free whatever the lanes' depths are); deeper entries spill to the local array.  0 = off.
#ifndef PBRT_SMEM_STACK
#define PBRT_SMEM_STACK 0
#endif
#define PBRT_TOP_SLOTS ((1u << PBRT_TOP_LEVELS) - 1u)

namespace dev {

#ifndef PBRT_INFINITE
#define PBRT_INFINITE 0   // InfiniteAreaLight device code, see dshade.cuh
#endif

struct DScene {
    const float4 *nodes;
    const float4 *prims;
    const uint4 *tris;
    const PbrtGpuMesh *meshes;
    const float *N;
    const float *uv;
    const PbrtGpuQuadric *quadrics;
    const PbrtGpuXform *xforms;
    const PbrtGpuMaterial *materials;
    const PbrtGpuMeasured *measured;
    const PbrtGpuKdNode *kdNodes;
    const PbrtGpuBrdfSample *kdData;
    const PbrtGpuLight *lights;
    const PbrtGpuLightShape *lightShapes;
    const PbrtGpuAnimated *animated;   // TransformedPrimitives (null / 0 for static scenes)
    uint32_t nNodes, nPrims, nLights, nAnimated;
    // the top PBRT_TOP_LEVELS levels of the scene BVH once more, in HEAP order (slot h-1 holds the node with heap index h:
    // root 1, first child 2h, second child 2h+1; slots without a node stay zero): one contiguous block that the persistent
    // kernels copy into shared memory with a TMA bulk copy (cp.async.bulk) -- see TopNodes below
    const float4 *topNodes;
#if PBRT_INFINITE
    const PbrtGpuEnvMap *envMaps;   // InfiniteAreaLights' radiance maps and sampling distributions (PbrtGpuLight::nShapes indexes)
    const float *envData;
#endif
#if PBRT_KD_INTERLEAVED
    const float4 *kdRec;          // 2 x float4 per kd node: {splitPos, bits, p.x, p.y} {p.z, v.r, v.g, v.b}
#endif
#if PBRT_PAIR_NODES
    const float4 *pairs;          // 4 x float4 per interior node: both children's node records ({box, ref, meta} each);
                                  // ref = the child's own pair index (interior) or its primitive offset (leaf)
    const uint32_t *pairOfNode;   // [nNodes] pair index of an interior node (used for the roots only)
#endif
};

// TransformedPrimitives whose interpolated transforms a path slot of the vertex wavefront carries (kwave.cuh)
constexpr uint32_t ANIM_CACHE_MAX = 2;
struct Ray {
    V3 o, d;
    float mint, maxt;
    float time;   // constant along a camera path (path.cpp:109, integrator.cpp:127,148); read by TransformedPrimitives only
    // AnimatedTransform::Interpolate(time) of the scene's first ANIM_CACHE_MAX TransformedPrimitives, computed once per camera
    // sample instead of once per ray and instance (8 float4 each -- m, mInv row by row -- `xfStride` float4 apart), or null
    const float4 *xfCache = nullptr;
    uint32_t xfStride = 0;
};
DEV V3 RayAt(const Ray &r, float t) { return r.o + r.d * t; }   // geometry.h:333

struct DiffGeom {   // core/diffgeom.h (texture differentials omitted)
    V3 p, nn, dpdu, dpdv;
    float u, v;
};
// DifferentialGeometry ctor, core/diffgeom.cpp:40-55
DEV void InitDG(DiffGeom *dg, const V3 &P, const V3 &DPDU, const V3 &DPDV, float uu, float vv, uint32_t shapeFlags) {
    dg->p = P; dg->dpdu = DPDU; dg->dpdv = DPDV;
    dg->nn = Normalize(Cross(DPDU, DPDV));
    dg->u = uu; dg->v = vv;
    bool ro = (shapeFlags & PBRTGPU_SHAPE_REVERSE_ORIENT) != 0, sh = (shapeFlags & PBRTGPU_SHAPE_SWAPS_HANDED) != 0;
    if (ro != sh) dg->nn = dg->nn * -1.f;
}

struct Hit {
    int prim;        // index into prims / -1
    float t;         // == ray.maxt after traversal
    float b1, b2;    // triangle barycentrics
    float maxtIn;    // ray.maxt at the moment the winning primitive was tested (quadric dg replay)
};

// ---- Triangle::Intersect core, shapes/trianglemesh.cpp:127-165 / :211-243 ------
DEV bool TriTest(const V3 &p1, const V3 &p2, const V3 &p3, const Ray &ray, float *tOut, float *b1Out, float *b2Out) {
    V3 e1 = p2 - p1;
    V3 e2 = p3 - p1;
    V3 s1 = Cross(ray.d, e2);
    float divisor = Dot(s1, e1);
    if (divisor == 0.f) return false;
    float invDivisor = 1.f / divisor;
    V3 s = ray.o - p1;
    float b1 = Dot(s, s1) * invDivisor;
    if (b1 < 0.f || b1 > 1.f) return false;
    V3 s2 = Cross(s, e1);
    float b2 = Dot(ray.d, s2) * invDivisor;
    if (b2 < 0.f || b1 + b2 > 1.f) return false;
    float t = Dot(e2, s2) * invDivisor;
    if (t < ray.mint || t > ray.maxt) return false;
    *tOut = t; *b1Out = b1; *b2Out = b2;
    return true;
}

DEV void GetUVs(const DScene &sc, const uint4 &tri, bool hasUV, float uv[3][2]) {   // trianglemesh.h:86-101
    if (hasUV) {
        uv[0][0] = __ldg(sc.uv + 2 * tri.x); uv[0][1] = __ldg(sc.uv + 2 * tri.x + 1);
        uv[1][0] = __ldg(sc.uv + 2 * tri.y); uv[1][1] = __ldg(sc.uv + 2 * tri.y + 1);
        uv[2][0] = __ldg(sc.uv + 2 * tri.z); uv[2][1] = __ldg(sc.uv + 2 * tri.z + 1);
    } else {
        uv[0][0] = 0.f; uv[0][1] = 0.f; uv[1][0] = 1.f; uv[1][1] = 0.f; uv[2][0] = 1.f; uv[2][1] = 1.f;
    }
}

// the rest of Triangle::Intersect (dpdu/dpdv, (u,v), DifferentialGeometry), trianglemesh.cpp:167-205
DEV void TriDG(const DScene &sc, const uint4 &tri, const V3 &p1, const V3 &p2, const V3 &p3, const Ray &ray, float t,
               float b1, float b2, DiffGeom *dg) {
    const uint32_t mflags = sc.meshes[tri.w].flags;
    V3 e1 = p2 - p1, e2 = p3 - p1;
    V3 dpdu, dpdv;
    float uvs[3][2];
    GetUVs(sc, tri, (mflags & PBRTGPU_MESH_HAS_UV) != 0, uvs);
    float du1 = uvs[0][0] - uvs[2][0];
    float du2 = uvs[1][0] - uvs[2][0];
    float dv1 = uvs[0][1] - uvs[2][1];
    float dv2 = uvs[1][1] - uvs[2][1];
    V3 dp1 = p1 - p3, dp2 = p2 - p3;
    float determinant = du1 * dv2 - dv1 * du2;
    if (determinant == 0.f) {
        CoordinateSystem(Normalize(Cross(e2, e1)), &dpdu, &dpdv);
    } else {
        float invdet = 1.f / determinant;
        dpdu = (dv2 * dp1 - dv1 * dp2) * invdet;
        dpdv = (-du2 * dp1 + du1 * dp2) * invdet;
    }
    float b0 = 1 - b1 - b2;
    float tu = b0 * uvs[0][0] + b1 * uvs[1][0] + b2 * uvs[2][0];
    float tv = b0 * uvs[0][1] + b1 * uvs[1][1] + b2 * uvs[2][1];
    InitDG(dg, RayAt(ray, t), dpdu, dpdv, tu, tv, mflags);
}

// ---- Sphere::Intersect / IntersectP, shapes/sphere.cpp:58-209 -------------------
DEV bool SphereIntersect(const DScene &sc, const PbrtGpuQuadric &q, const Ray &r, float *tHit, DiffGeom *dg) {
    const PbrtGpuXform &x = sc.xforms[q.xform];
    Ray ray;
    ray.o = XfPoint(x.mInv, r.o); ray.d = XfVector(x.mInv, r.d);
    ray.mint = r.mint; ray.maxt = r.maxt;
    const float radius = q.radius, zmin = q.zmin, zmax = q.zmax, phiMax = q.phiMax;
    float A = ray.d.x * ray.d.x + ray.d.y * ray.d.y + ray.d.z * ray.d.z;
    float B = 2 * (ray.d.x * ray.o.x + ray.d.y * ray.o.y + ray.d.z * ray.o.z);
    float C = ray.o.x * ray.o.x + ray.o.y * ray.o.y + ray.o.z * ray.o.z - radius * radius;
    float t0, t1;
    if (!Quadratic(A, B, C, &t0, &t1)) return false;
    if (t0 > ray.maxt || t1 < ray.mint) return false;
    float thit = t0;
    if (t0 < ray.mint) {
        thit = t1;
        if (thit > ray.maxt) return false;
    }
    V3 phit = RayAt(ray, thit);
    if (phit.x == 0.f && phit.y == 0.f) phit.x = 1e-5f * radius;
    float phi = ref_atan2f(phit.y, phit.x);
    if (phi < 0.f) phi += 2.f * PI_F;
    if ((zmin > -radius && phit.z < zmin) || (zmax < radius && phit.z > zmax) || phi > phiMax) {
        if (thit == t1) return false;
        if (t1 > ray.maxt) return false;
        thit = t1;
        phit = RayAt(ray, thit);
        if (phit.x == 0.f && phit.y == 0.f) phit.x = 1e-5f * radius;
        phi = ref_atan2f(phit.y, phit.x);
        if (phi < 0.f) phi += 2.f * PI_F;
        if ((zmin > -radius && phit.z < zmin) || (zmax < radius && phit.z > zmax) || phi > phiMax) return false;
    }
    if (dg) {
        float u = phi / phiMax;
        float theta = ref_acosf(Clampf(phit.z / radius, -1.f, 1.f));
        float v = (theta - q.thetaMin) / (q.thetaMax - q.thetaMin);
        float zradius = sqrtf(phit.x * phit.x + phit.y * phit.y);
        float invzradius = 1.f / zradius;
        float cosphi = phit.x * invzradius;
        float sinphi = phit.y * invzradius;
        V3 dpdu(-phiMax * phit.y, phiMax * phit.x, 0);
        V3 dpdv = (q.thetaMax - q.thetaMin) * V3(phit.z * cosphi, phit.z * sinphi, -radius * ref_sinf(theta));
        InitDG(dg, XfPoint(x.m, phit), XfVector(x.m, dpdu), XfVector(x.m, dpdv), u, v, q.flags);
    }
    *tHit = thit;
    return true;
}

// ---- Disk::Intersect / IntersectP, shapes/disk.cpp:56-130 -----------------------
DEV bool DiskIntersect(const DScene &sc, const PbrtGpuQuadric &q, const Ray &r, float *tHit, DiffGeom *dg) {
    const PbrtGpuXform &x = sc.xforms[q.xform];
    Ray ray;
    ray.o = XfPoint(x.mInv, r.o); ray.d = XfVector(x.mInv, r.d);
    ray.mint = r.mint; ray.maxt = r.maxt;
    if ((double)fabsf(ray.d.z) < 1e-7) return false;   // double compare, disk.cpp:62
    float thit = (q.height - ray.o.z) / ray.d.z;
    if (thit < ray.mint || thit > ray.maxt) return false;
    V3 phit = RayAt(ray, thit);
    float dist2 = phit.x * phit.x + phit.y * phit.y;
    if (dist2 > q.radius * q.radius || dist2 < q.innerRadius * q.innerRadius) return false;
    float phi = ref_atan2f(phit.y, phit.x);
    if (phi < 0) phi = (float)((double)phi + 2. * (double)PI_F);   // double arithmetic, disk.cpp:77
    if (phi > q.phiMax) return false;
    if (dg) {
        float u = phi / q.phiMax;
        float R = sqrtf(dist2);
        float oneMinusV = ((R - q.innerRadius) / (q.radius - q.innerRadius));
        float v = 1.f - oneMinusV;
        V3 dpdu(-q.phiMax * phit.y, q.phiMax * phit.x, 0.f);
        V3 dpdv(phit.x, phit.y, 0.f);
        dpdv = dpdv * ((q.radius - q.innerRadius) / R);
        InitDG(dg, XfPoint(x.m, phit), XfVector(x.m, dpdu), XfVector(x.m, dpdv), u, v, q.flags);
    }
    *tHit = thit;
    return true;
}

DEV_NI2 bool QuadricIntersect(const DScene &sc, const PbrtGpuQuadric &q, const Ray &r, float *tHit, DiffGeom *dg) {
    return q.type == PBRTGPU_PRIM_SPHERE ? SphereIntersect(sc, q, r, tHit, dg) : DiskIntersect(sc, q, r, tHit, dg);
}

// ---- BVH traversal ---------------------------------------------------------------
// One loop serves BVHAccel::Intersect (bvh.cpp:403-454) and IntersectP
// (:457-503): same node order (near child first by dirIsNeg[axis]), same strict
// slab comparisons (:126-148), same in-place shrinking of ray.maxt; `anyHit`
// only adds the early return.  Statistics are optional (null pointers = off).
struct TraceStats {
    unsigned long long nodes, prims;
};

DEV bool BoxTest(const float4 &n0, const float4 &n1, const Ray &ray, const V3 &invDir, const bool neg[3]) {
    // n0 = {bmin.x, bmin.y, bmin.z, bmax.x}, n1 = {bmax.y, bmax.z, offset, meta}
    float tmin = ((neg[0] ? n0.w : n0.x) - ray.o.x) * invDir.x;
    float tmax = ((neg[0] ? n0.x : n0.w) - ray.o.x) * invDir.x;
    float tymin = ((neg[1] ? n1.x : n0.y) - ray.o.y) * invDir.y;
    float tymax = ((neg[1] ? n0.y : n1.x) - ray.o.y) * invDir.y;
    if ((tmin > tymax) || (tymin > tmax)) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = ((neg[2] ? n1.y : n0.z) - ray.o.z) * invDir.z;
    float tzmax = ((neg[2] ? n0.z : n1.y) - ray.o.z) * invDir.z;
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    return (tmin < ray.maxt) && (tmax > ray.mint);
}

template <bool STATS> __device__ __noinline__ bool AnimatedIntersect(const DScene &sc, uint32_t index, Ray &ray, bool anyHit,
                                                                      Hit *hit, TraceStats *st);

// host side: the heap-ordered copy of the first `levels` levels of the scene BVH (2 x float4 per slot, like `nodes`)
inline void BuildTopNodes(const PbrtGpuSceneDesc *desc, int levels, std::vector<float4> *top) {
    const uint32_t slots = levels > 0 ? (1u << levels) - 1u : 0u;
    top->assign(2 * (size_t)slots + 2, make_float4(0.f, 0.f, 0.f, 0.f));
    if (!slots || !desc->nNodes) return;
    std::vector<uint32_t> nodeOf(slots + 1, 0xffffffffu);
    nodeOf[1] = 0;
    for (uint32_t h = 1; h <= slots; ++h) {
        const uint32_t ni = nodeOf[h];
        if (ni == 0xffffffffu) continue;
        const PbrtGpuBVHNode &n = desc->nodes[ni];
        const uint32_t meta = (uint32_t)n.nPrimitives | ((uint32_t)n.axis << 8);
        float off, mf;
        memcpy(&off, &n.offset, 4);
        memcpy(&mf, &meta, 4);
        (*top)[2 * (size_t)(h - 1)] = make_float4(n.bmin[0], n.bmin[1], n.bmin[2], n.bmax[0]);
        (*top)[2 * (size_t)(h - 1) + 1] = make_float4(n.bmax[1], n.bmax[2], off, mf);
        if (n.nPrimitives == 0 && 2 * h + 1 <= slots) { nodeOf[2 * h] = ni + 1; nodeOf[2 * h + 1] = n.offset; }
    }
}

#if PBRT_KD_INTERLEAVED
// host side: the interleaved kd records of a flattened scene
inline void BuildKdRecords(const PbrtGpuSceneDesc *desc, std::vector<float4> *rec) {
    rec->resize(2 * (size_t)desc->nKdNodes);
    for (uint32_t i = 0; i < desc->nKdNodes; ++i) {
        const PbrtGpuKdNode &n = desc->kdNodes[i];
        const PbrtGpuBrdfSample &b = desc->kdData[i];
        float bits;
        memcpy(&bits, &n.bits, 4);
        (*rec)[2 * (size_t)i] = make_float4(n.splitPos, bits, b.p[0], b.p[1]);
        (*rec)[2 * (size_t)i + 1] = make_float4(b.p[2], b.v[0], b.v[1], b.v[2]);
    }
}
#endif

#if PBRT_PAIR_NODES
// host side: the pair records of a flattened scene (scene BVH first, nested BVHs by their PbrtGpuAnimated ranges)
inline void BuildNodePairs(const PbrtGpuSceneDesc *desc, std::vector<float4> *pairs, std::vector<uint32_t> *pairOfNode) {
    const uint32_t n = desc->nNodes;
    std::vector<uint32_t> base(n, 0u);
    for (uint32_t k = 0; k < desc->nAnimated; ++k)
        for (uint32_t i = 0; i < desc->animated[k].nNodes; ++i) base[desc->animated[k].firstNode + i] = desc->animated[k].firstNode;
    pairOfNode->assign(n, 0xffffffffu);
    uint32_t nPairs = 0;
    for (uint32_t i = 0; i < n; ++i)
        if (desc->nodes[i].nPrimitives == 0) (*pairOfNode)[i] = nPairs++;
    pairs->assign(4 * (size_t)nPairs, make_float4(0.f, 0.f, 0.f, 0.f));
    auto record = [&](uint32_t child, float4 *out) {
        const PbrtGpuBVHNode &c = desc->nodes[child];
        const uint32_t ref = c.nPrimitives ? c.offset : (*pairOfNode)[child];
        const uint32_t meta = (uint32_t)c.nPrimitives | ((uint32_t)c.axis << 8);
        float rf, mf;
        memcpy(&rf, &ref, 4);
        memcpy(&mf, &meta, 4);
        out[0] = make_float4(c.bmin[0], c.bmin[1], c.bmin[2], c.bmax[0]);
        out[1] = make_float4(c.bmax[1], c.bmax[2], rf, mf);
    };
    for (uint32_t i = 0; i < n; ++i) {
        if (desc->nodes[i].nPrimitives) continue;
        float4 *rec = pairs->data() + 4 * (size_t)(*pairOfNode)[i];
        record(i + 1, rec);                                   // first child: the next node (bvh.cpp:392)
        record(base[i] + desc->nodes[i].offset, rec + 2);     // second child: secondChildOffset (bvh.cpp:394)
    }
}

// BoxTest without its last line: false if the slabs miss whatever maxt is, else *tminOut for the `tmin < maxt` that decides
DEV bool BoxTestDeferred(const float4 &n0, const float4 &n1, const Ray &ray, const V3 &invDir, const bool neg[3], float *tminOut) {
    float tmin = ((neg[0] ? n0.w : n0.x) - ray.o.x) * invDir.x;
    float tmax = ((neg[0] ? n0.x : n0.w) - ray.o.x) * invDir.x;
    float tymin = ((neg[1] ? n1.x : n0.y) - ray.o.y) * invDir.y;
    float tymax = ((neg[1] ? n0.y : n1.x) - ray.o.y) * invDir.y;
    if ((tmin > tymax) || (tymin > tmax)) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = ((neg[2] ? n1.y : n0.z) - ray.o.z) * invDir.z;
    float tzmax = ((neg[2] ? n0.z : n1.y) - ray.o.z) * invDir.z;
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    *tminOut = tmin;
    return tmax > ray.mint;
}

template <bool TOP, bool ANIM>
DEV bool TraverseBVHPairs(const DScene &sc, uint32_t nodeBase, uint32_t primBase, Ray &ray, bool anyHit, Hit *hit) {
    V3 invDir(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);
    bool neg[3] = {invDir.x < 0, invDir.y < 0, invDir.z < 0};
    uint32_t todoRef[64], todoMeta[64];
    float todoT[64];
    uint32_t todoOffset = 0;
    bool found = false;
    uint32_t ref, meta;
    {   // the root is nobody's child: its own record
        const float4 n0 = __ldg(sc.nodes + 2 * nodeBase);
        const float4 n1 = __ldg(sc.nodes + 2 * nodeBase + 1);
        if (!BoxTest(n0, n1, ray, invDir, neg)) return false;
        meta = __float_as_uint(n1.w);
        ref = (meta & 0xffu) ? __float_as_uint(n1.z) : __ldg(sc.pairOfNode + nodeBase);
    }
    while (true) {
        const uint32_t nPrims = meta & 0xffu;
        bool descend = false;
        if (nPrims > 0) {
            for (uint32_t i = 0; i < nPrims; ++i) {
                const uint32_t pi = primBase + ref + i;
                const float4 r0 = __ldg(sc.prims + 3 * pi);
                const float4 r1 = __ldg(sc.prims + 3 * pi + 1);
                const float4 r2 = __ldg(sc.prims + 3 * pi + 2);
                const uint32_t type = __float_as_uint(r0.w) & 0xffu;
                float t, b1 = 0.f, b2 = 0.f;
                bool h;
                if (type == PBRTGPU_PRIM_TRIANGLE)
                    h = TriTest(V3(r0.x, r0.y, r0.z), V3(r1.x, r1.y, r1.z), V3(r2.x, r2.y, r2.z), ray, &t, &b1, &b2);
                else if (TOP && ANIM && type == PBRTGPU_PRIM_ANIMATED) {
                    if (AnimatedIntersect<false>(sc, __float_as_uint(r1.w), ray, anyHit, hit, nullptr)) {
                        if (anyHit) return true;
                        found = true;
                    }
                    continue;
                } else
                    h = QuadricIntersect(sc, sc.quadrics[__float_as_uint(r1.w)], ray, &t, nullptr);
                if (h) {
                    if (anyHit) { hit->prim = (int)pi; return true; }
                    hit->prim = (int)pi; hit->b1 = b1; hit->b2 = b2; hit->maxtIn = ray.maxt;
                    ray.maxt = t;
                    found = true;
                }
            }
        } else {
            const float4 *p = sc.pairs + 4 * (size_t)ref;
            const float4 a0 = __ldg(p), a1 = __ldg(p + 1), b0 = __ldg(p + 2), b1 = __ldg(p + 3);
            const bool secondFirst = neg[(meta >> 8) & 0xffu];   // bvh.cpp:436-444: the far child waits on the stack
            const float4 &N0 = secondFirst ? b0 : a0, &N1 = secondFirst ? b1 : a1;
            const float4 &F0 = secondFirst ? a0 : b0, &F1 = secondFirst ? a1 : b1;
            const bool hitNear = BoxTest(N0, N1, ray, invDir, neg);
            float tFar = 0.f;
            const bool hitFar = BoxTestDeferred(F0, F1, ray, invDir, neg, &tFar) && tFar < ray.maxt;
            if (hitNear) {
                if (hitFar) {
                    todoRef[todoOffset] = __float_as_uint(F1.z); todoMeta[todoOffset] = __float_as_uint(F1.w); todoT[todoOffset] = tFar;
                    ++todoOffset;
                }
                ref = __float_as_uint(N1.z); meta = __float_as_uint(N1.w);
                descend = true;
            } else if (hitFar) {
                ref = __float_as_uint(F1.z); meta = __float_as_uint(F1.w);
                descend = true;
            }
        }
        if (descend) continue;
        bool got = false;
        while (todoOffset > 0) {
            --todoOffset;
            if (todoT[todoOffset] < ray.maxt) {   // the far child's own box test, with the maxt of this moment
                ref = todoRef[todoOffset]; meta = todoMeta[todoOffset];
                got = true;
                break;
            }
        }
        if (!got) break;
    }
    return found;
}
#endif   // PBRT_PAIR_NODES

// TOP: the scene's own BVH (nodes/primitives from 0); otherwise the BVH nested in a TransformedPrimitive,
// whose node and primitive offsets are relative to nodeBase / primBase.
template <bool STATS, bool TOP, bool ANIM>
DEV bool TraverseBVH(const DScene &sc, uint32_t nodeBase, uint32_t primBase, Ray &ray, bool anyHit, Hit *hit, TraceStats *st,
                     const float4 *smemTop = nullptr, uint32_t *smemStack = nullptr) {
#if PBRT_PAIR_NODES
    if (!STATS) return TraverseBVHPairs<TOP, ANIM>(sc, nodeBase, primBase, ray, anyHit, hit);
#endif
#if PBRT_TOP_LEVELS
    // heap index of the node to test next while it lies inside the shared-memory table, else 0; todoHeap shadows todo[]
    uint32_t heap = (TOP && smemTop) ? 1u : 0u;
    uint16_t todoHeap[64];
#endif
#if PBRT_SMEM_STACK
    uint32_t *const todoS = smemStack;   // this thread's column: entry i at todoS[i * 128]
#define TODO_PUSH(v) do { if (todoS && todoOffset < PBRT_SMEM_STACK) todoS[todoOffset * 128] = (v); else todo[todoOffset] = (v); ++todoOffset; } while (0)
#define TODO_POP() ((--todoOffset, (todoS && todoOffset < PBRT_SMEM_STACK)) ? todoS[todoOffset * 128] : todo[todoOffset])
#else
#define TODO_PUSH(v) (todo[todoOffset++] = (v))
#define TODO_POP() (todo[--todoOffset])
#endif
    V3 invDir(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);
    bool neg[3] = {invDir.x < 0, invDir.y < 0, invDir.z < 0};
    uint32_t todo[64];
    uint32_t todoOffset = 0, nodeNum = 0;
    bool found = false;
    // "while-while" form of the reference's loop (Aila & Laine 2009): the inner loop walks interior nodes until THIS ray
    // reaches a leaf whose box it hits (or runs out of nodes); the lanes of a warp leave it one by one and reconverge
    // behind it, so the primitive tests below run with every lane that has a leaf instead of being serialised against the
    // other lanes' node tests.  Per ray nothing moves: same nodes, same order, same maxt at every test.
    while (true) {
        uint32_t leafOffset = 0, leafCount = 0;
        bool done = false;
        while (true) {
            float4 n0, n1;
#if PBRT_TOP_LEVELS
            if (heap) { n0 = smemTop[2 * (heap - 1)]; n1 = smemTop[2 * (heap - 1) + 1]; }
            else
#endif
            {
                n0 = __ldg(sc.nodes + 2 * (nodeBase + nodeNum));
                n1 = __ldg(sc.nodes + 2 * (nodeBase + nodeNum) + 1);
            }
            if (STATS) st->nodes++;
            if (BoxTest(n0, n1, ray, invDir, neg)) {
                const uint32_t offset = __float_as_uint(n1.z), meta = __float_as_uint(n1.w);
                const uint32_t nPrims = meta & 0xffu;
                if (nPrims > 0) { leafOffset = offset; leafCount = nPrims; break; }
                const uint32_t axis = (meta >> 8) & 0xffu;
#if PBRT_TOP_LEVELS
                const uint32_t kids = (heap && 2 * heap + 1 <= PBRT_TOP_SLOTS) ? 2 * heap : 0u;   // heap index of the first child
#endif
                if (neg[axis]) {
#if PBRT_TOP_LEVELS
                    todoHeap[todoOffset] = (uint16_t)kids;
                    heap = kids ? kids + 1 : 0u;
#endif
                    TODO_PUSH(nodeNum + 1);
                    nodeNum = offset;
                } else {
#if PBRT_TOP_LEVELS
                    todoHeap[todoOffset] = (uint16_t)(kids ? kids + 1 : 0u);
                    heap = kids;
#endif
                    TODO_PUSH(offset);
                    nodeNum = nodeNum + 1;
                }
            } else {
                if (todoOffset == 0) { done = true; break; }
                nodeNum = TODO_POP();
#if PBRT_TOP_LEVELS
                heap = todoHeap[todoOffset];
#endif
            }
        }
        if (done) break;
        for (uint32_t i = 0; i < leafCount; ++i) {
            const uint32_t pi = primBase + leafOffset + i;
            const float4 r0 = __ldg(sc.prims + 3 * pi);
            const float4 r1 = __ldg(sc.prims + 3 * pi + 1);
            const float4 r2 = __ldg(sc.prims + 3 * pi + 2);
            if (STATS) st->prims++;
            const uint32_t type = __float_as_uint(r0.w) & 0xffu;
            float t, b1 = 0.f, b2 = 0.f;
            bool h;
            if (type == PBRTGPU_PRIM_TRIANGLE)
                h = TriTest(V3(r0.x, r0.y, r0.z), V3(r1.x, r1.y, r1.z), V3(r2.x, r2.y, r2.z), ray, &t, &b1, &b2);
            else if (TOP && ANIM && type == PBRTGPU_PRIM_ANIMATED) {
                // TransformedPrimitive::Intersect fills the hit and shrinks ray.maxt itself
                if (AnimatedIntersect<STATS>(sc, __float_as_uint(r1.w), ray, anyHit, hit, st)) {
                    if (anyHit) return true;
                    found = true;
                }
                continue;
            } else
                h = QuadricIntersect(sc, sc.quadrics[__float_as_uint(r1.w)], ray, &t, nullptr);
            if (h) {
                if (anyHit) { hit->prim = (int)pi; return true; }
                hit->prim = (int)pi; hit->b1 = b1; hit->b2 = b2; hit->maxtIn = ray.maxt;
                ray.maxt = t;   // GeometricPrimitive::Intersect, primitive.cpp:175
                found = true;
            }
        }
        if (todoOffset == 0) break;
        nodeNum = TODO_POP();
#if PBRT_TOP_LEVELS
        heap = todoHeap[todoOffset];
#endif
    }
    return found;
#undef TODO_PUSH
#undef TODO_POP
}

// one primitive tested directly (a TransformedPrimitive around a single refined primitive: no BVH, api.cpp:1040-1041)
DEV bool SinglePrimTest(const DScene &sc, uint32_t pi, Ray &ray, bool anyHit, Hit *hit) {
    const float4 r0 = __ldg(sc.prims + 3 * pi);
    const float4 r1 = __ldg(sc.prims + 3 * pi + 1);
    const float4 r2 = __ldg(sc.prims + 3 * pi + 2);
    const uint32_t type = __float_as_uint(r0.w) & 0xffu;
    float t, b1 = 0.f, b2 = 0.f;
    bool h;
    if (type == PBRTGPU_PRIM_TRIANGLE)
        h = TriTest(V3(r0.x, r0.y, r0.z), V3(r1.x, r1.y, r1.z), V3(r2.x, r2.y, r2.z), ray, &t, &b1, &b2);
    else
        h = QuadricIntersect(sc, sc.quadrics[__float_as_uint(r1.w)], ray, &t, nullptr);
    if (!h) return false;
    hit->prim = (int)pi;
    if (anyHit) return true;
    hit->b1 = b1; hit->b2 = b2; hit->maxtIn = ray.maxt;
    ray.maxt = t;
    return true;
}

// w2p(r): Transform::operator()(const Ray &), transform.h:253-262
DEV void ToPrimitiveSpace(const DXf &w2p, const Ray &r, Ray *out) {
    out->o = XfPoint(w2p.m, r.o);
    out->d = XfVector(w2p.m, r.d);
    out->mint = r.mint; out->maxt = r.maxt; out->time = r.time;
    out->xfCache = r.xfCache; out->xfStride = r.xfStride;
}

// WorldToPrimitive.Interpolate(ray.time) of TransformedPrimitive `index` (primitive.cpp:97,112): from the ray's cache if it
// has one -- the same function's result, stored when the camera sample (and with it ray.time) was chosen -- else computed
DEV void RayW2P(const DScene &sc, uint32_t index, const Ray &r, DXf *out) {
    if (r.xfCache && index < ANIM_CACHE_MAX) {
        const float4 *c = r.xfCache + (size_t)index * 8 * r.xfStride;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float4 a = c[(size_t)e * r.xfStride], b = c[(size_t)(4 + e) * r.xfStride];
            out->m[4 * e] = a.x; out->m[4 * e + 1] = a.y; out->m[4 * e + 2] = a.z; out->m[4 * e + 3] = a.w;
            out->mInv[4 * e] = b.x; out->mInv[4 * e + 1] = b.y; out->mInv[4 * e + 2] = b.z; out->mInv[4 * e + 3] = b.w;
        }
    } else
        AnimInterpolate(sc.animated + index, r.time, out);
}
// ... and filling such a cache (`dst` = the slot's first cache field)
DEV void FillXfCache(const DScene &sc, float time, float4 *dst, uint32_t stride) {
    for (uint32_t i = 0; i < sc.nAnimated && i < ANIM_CACHE_MAX; ++i) {
        DXf x;
        AnimInterpolate(sc.animated + i, time, &x);
        float4 *c = dst + (size_t)i * 8 * stride;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            c[(size_t)e * stride] = make_float4(x.m[4 * e], x.m[4 * e + 1], x.m[4 * e + 2], x.m[4 * e + 3]);
            c[(size_t)(4 + e) * stride] = make_float4(x.mInv[4 * e], x.mInv[4 * e + 1], x.mInv[4 * e + 2], x.mInv[4 * e + 3]);
        }
    }
}

// TransformedPrimitive::Intersect / IntersectP, core/primitive.cpp:95-124 (the DifferentialGeometry is taken back
// to world space in HitGeometry, when -- and if -- the hit is shaded)
template <bool STATS>
__device__ __noinline__ bool AnimatedIntersect(const DScene &sc, uint32_t index, Ray &r, bool anyHit, Hit *hit, TraceStats *st) {
    const PbrtGpuAnimated *a = sc.animated + index;
    DXf w2p;
    RayW2P(sc, index, r, &w2p);
    Ray ray;
    ToPrimitiveSpace(w2p, r, &ray);
    bool found;
    if (a->nNodes > 0) found = TraverseBVH<STATS, false, false>(sc, a->firstNode, a->firstPrimitive, ray, anyHit, hit, st);
    else found = SinglePrimTest(sc, a->firstPrimitive, ray, anyHit, hit);
    if (found) r.maxt = ray.maxt;
    return found;
}

// ANIM: the scene may contain TransformedPrimitives (instantiations without it carry no trace of them)
template <bool STATS, bool ANIM>
DEV bool Traverse(const DScene &sc, Ray &ray, bool anyHit, Hit *hit, TraceStats *st, const float4 *smemTop = nullptr,
                  uint32_t *smemStack = nullptr) {
    hit->prim = -1;
    const bool found = TraverseBVH<STATS, true, ANIM>(sc, 0u, 0u, ray, anyHit, hit, st, smemTop, smemStack);
    hit->t = ray.maxt;
    return found;
}

// ---- step-wise traversal (wavefront trace kernel) ---------------------------------------
// Same walk as Traverse<>, cut into single-node steps so that a persistent lane can pick up a new
// ray from the queue the moment its current one finishes (no lane waits for its warp's longest ray).
struct Trav {
    Ray ray;            // in the space of the BVH being walked (primitive space inside a TransformedPrimitive)
    V3 invDir;
    uint32_t nodeNum, todoOffset;
    bool neg[3];
    bool anyHit, found;
    Hit hit;
    // ---- TransformedPrimitive in progress (ANIM instantiation only)
    int inst;                    // animated[] index or -1
    uint32_t nodeBase, primBase; // of the BVH being walked
    uint32_t todoBase;           // stack level at which the nested walk started
    uint32_t leafNext, leafEnd;  // rest of the scene-BVH leaf that contained the TransformedPrimitive
    V3 wo, wd;                   // the world-space ray
    uint32_t leafFirst, leafLast; // the leaf TravDescend stopped in front of (primitive range)
#if PBRT_SMEM_STACK
    uint32_t *sstk;              // this thread's column of the CTA's shared-memory stack (entry i at sstk[i * 128]) or null
#endif
#if PBRT_TOP_LEVELS
    uint32_t heap;               // heap index of nodeNum while it lies inside the shared-memory top table, else 0
    uint16_t todoHeap[64];
#endif
    uint32_t todo[64];
#if PBRT_PAIR_NODES
    // pair walk: the node to enter next (it has passed its box test) or `needRoot`; todo[] holds the refs
    uint32_t curRef, curMeta;
    bool needRoot;
    uint32_t todoMeta[64];
    float todoT[64];
#endif
};
DEV void TravSetRay(Trav &t) {
    t.invDir = V3(1.f / t.ray.d.x, 1.f / t.ray.d.y, 1.f / t.ray.d.z);
    t.neg[0] = t.invDir.x < 0; t.neg[1] = t.invDir.y < 0; t.neg[2] = t.invDir.z < 0;
}
DEV void TravInit(Trav &t, const Ray &ray, bool anyHit) {
    t.ray = ray;
    TravSetRay(t);
    t.nodeNum = 0; t.todoOffset = 0;
    t.anyHit = anyHit; t.found = false;
    t.hit.prim = -1; t.hit.b1 = t.hit.b2 = 0.f; t.hit.maxtIn = 0.f; t.hit.t = 0.f;
    t.inst = -1; t.nodeBase = t.primBase = 0; t.todoBase = 0; t.leafNext = t.leafEnd = 0;
    t.leafFirst = t.leafLast = 0;
#if PBRT_TOP_LEVELS
    t.heap = 1;
#endif
#if PBRT_PAIR_NODES
    t.curRef = t.curMeta = 0; t.needRoot = true;
#endif
}
#if PBRT_PAIR_NODES
// TravStep over the pair layout: one step = the root's test, one interior node (both children tested), or one leaf
template <bool ANIM>
DEV bool TravStepPairs(const DScene &sc, Trav &t) {
    uint32_t first = 0, end = 0;
    bool leaf = false;
    if (ANIM && t.inst < 0 && t.leafNext < t.leafEnd) {   // back from a TransformedPrimitive: the rest of that leaf
        first = t.leafNext; end = t.leafEnd;
        t.leafNext = t.leafEnd = 0;
        leaf = true;
    } else if (t.needRoot) {
        t.needRoot = false;
        const uint32_t ni = ANIM ? t.nodeBase : 0u;
        const float4 n0 = __ldg(sc.nodes + 2 * ni);
        const float4 n1 = __ldg(sc.nodes + 2 * ni + 1);
        if (BoxTest(n0, n1, t.ray, t.invDir, t.neg)) {
            t.curMeta = __float_as_uint(n1.w);
            t.curRef = (t.curMeta & 0xffu) ? __float_as_uint(n1.z) : __ldg(sc.pairOfNode + ni);
            return false;
        }
    } else {
        const uint32_t nPrims = t.curMeta & 0xffu;
        if (nPrims > 0) {
            first = (ANIM ? t.primBase : 0u) + t.curRef;
            end = first + nPrims;
            leaf = true;
        } else {
            const float4 *p = sc.pairs + 4 * (size_t)t.curRef;
            const float4 a0 = __ldg(p), a1 = __ldg(p + 1), b0 = __ldg(p + 2), b1 = __ldg(p + 3);
            const bool secondFirst = t.neg[(t.curMeta >> 8) & 0xffu];
            const float4 &N0 = secondFirst ? b0 : a0, &N1 = secondFirst ? b1 : a1;
            const float4 &F0 = secondFirst ? a0 : b0, &F1 = secondFirst ? a1 : b1;
            const bool hitNear = BoxTest(N0, N1, t.ray, t.invDir, t.neg);
            float tFar = 0.f;
            const bool hitFar = BoxTestDeferred(F0, F1, t.ray, t.invDir, t.neg, &tFar) && tFar < t.ray.maxt;
            if (hitNear) {
                if (hitFar) {
                    t.todo[t.todoOffset] = __float_as_uint(F1.z); t.todoMeta[t.todoOffset] = __float_as_uint(F1.w);
                    t.todoT[t.todoOffset] = tFar;
                    ++t.todoOffset;
                }
                t.curRef = __float_as_uint(N1.z); t.curMeta = __float_as_uint(N1.w);
                return false;
            }
            if (hitFar) {
                t.curRef = __float_as_uint(F1.z); t.curMeta = __float_as_uint(F1.w);
                return false;
            }
        }
    }
    if (leaf) {
        for (uint32_t pi = first; pi < end; ++pi) {
            const float4 r0 = __ldg(sc.prims + 3 * pi);
            const float4 r1 = __ldg(sc.prims + 3 * pi + 1);
            const float4 r2 = __ldg(sc.prims + 3 * pi + 2);
            const uint32_t type = __float_as_uint(r0.w) & 0xffu;
            float tt, b1 = 0.f, b2 = 0.f;
            bool h;
            if (type == PBRTGPU_PRIM_TRIANGLE)
                h = TriTest(V3(r0.x, r0.y, r0.z), V3(r1.x, r1.y, r1.z), V3(r2.x, r2.y, r2.z), t.ray, &tt, &b1, &b2);
            else if (ANIM && type == PBRTGPU_PRIM_ANIMATED) {
                const PbrtGpuAnimated *a = sc.animated + __float_as_uint(r1.w);
                DXf w2p;
                RayW2P(sc, __float_as_uint(r1.w), t.ray, &w2p);
                Ray rl;
                ToPrimitiveSpace(w2p, t.ray, &rl);
                if (a->nNodes == 0) {
                    if (SinglePrimTest(sc, a->firstPrimitive, rl, t.anyHit, &t.hit)) {
                        t.found = true;
                        t.ray.maxt = rl.maxt;
                        if (t.anyHit) { t.hit.t = t.ray.maxt; return true; }
                    }
                    continue;
                }
                t.wo = t.ray.o; t.wd = t.ray.d;
                t.ray = rl;
                TravSetRay(t);
                t.inst = (int)__float_as_uint(r1.w);
                t.nodeBase = a->firstNode; t.primBase = a->firstPrimitive;
                t.todoBase = t.todoOffset;
                t.leafNext = pi + 1; t.leafEnd = end;
                if (t.leafNext == t.leafEnd) t.leafNext = t.leafEnd = 0;
                t.needRoot = true;
                return false;
            } else
                h = QuadricIntersect(sc, sc.quadrics[__float_as_uint(r1.w)], t.ray, &tt, nullptr);
            if (h) {
                t.hit.prim = (int)pi;
                t.found = true;
                if (t.anyHit) { t.hit.t = t.ray.maxt; return true; }
                t.hit.b1 = b1; t.hit.b2 = b2; t.hit.maxtIn = t.ray.maxt;
                t.ray.maxt = tt;
            }
        }
    }
    // ---- next node from the stack: entries whose tmin no longer beats maxt fail the box test they were waiting for
    const uint32_t floor = (ANIM && t.inst >= 0) ? t.todoBase : 0u;
    while (t.todoOffset > floor) {
        --t.todoOffset;
        if (t.todoT[t.todoOffset] < t.ray.maxt) {
            t.curRef = t.todo[t.todoOffset]; t.curMeta = t.todoMeta[t.todoOffset];
            return false;
        }
    }
    if (ANIM && t.inst >= 0) {
        // the nested walk is over: back to world space (r.maxt = ray.maxt, primitive.cpp:102)
        t.ray.o = t.wo; t.ray.d = t.wd;
        TravSetRay(t);
        t.inst = -1; t.nodeBase = t.primBase = 0;
        if (t.leafNext < t.leafEnd) return false;   // the next step tests the rest of the leaf
        while (t.todoOffset > 0) {
            --t.todoOffset;
            if (t.todoT[t.todoOffset] < t.ray.maxt) {
                t.curRef = t.todo[t.todoOffset]; t.curMeta = t.todoMeta[t.todoOffset];
                return false;
            }
        }
    }
    t.hit.t = t.ray.maxt;
    return true;
}
#endif   // PBRT_PAIR_NODES

// ---- the step-wise walk in two halves (the wavefront trace kernel's "while-while" loop) ------------------------
// TravDescend walks interior nodes until the ray stands before a leaf whose box it hits (1), is finished (2), or has
// used up `maxSteps` node tests (0); TravLeaf tests that leaf's primitives and pops the next node (true: finished).
// k_trace runs the two as separate warp-wide phases: all lanes descend, reconverge, then every lane that has a leaf
// tests it -- the primitive tests (double-precision cross products) no longer serialise against the other lanes' node
// tests.  Per ray the sequence of node tests, primitive tests and maxt updates is BVHAccel::Intersect's, as before.
#if PBRT_SMEM_STACK
#define TRAV_PUSH(t, v) do { if ((t).sstk && (t).todoOffset < PBRT_SMEM_STACK) (t).sstk[(t).todoOffset * 128] = (v); else (t).todo[(t).todoOffset] = (v); ++(t).todoOffset; } while (0)
#define TRAV_POP(t) ((--(t).todoOffset, ((t).sstk && (t).todoOffset < PBRT_SMEM_STACK)) ? (t).sstk[(t).todoOffset * 128] : (t).todo[(t).todoOffset])
#else
#define TRAV_PUSH(t, v) ((t).todo[(t).todoOffset++] = (v))
#define TRAV_POP(t) ((t).todo[--(t).todoOffset])
#endif

// the next node from the stack; true when there is none left (the ray is finished: result in t.found / t.hit / t.ray.maxt)
template <bool ANIM>
DEV bool TravPop(Trav &t) {
    if (ANIM && t.inst >= 0) {
        if (t.todoOffset > t.todoBase) {
            t.nodeNum = TRAV_POP(t);
#if PBRT_TOP_LEVELS
            t.heap = 0;
#endif
            return false;
        }
        // the nested walk is over: back to world space (r.maxt = ray.maxt, primitive.cpp:102)
        t.ray.o = t.wo; t.ray.d = t.wd;
        TravSetRay(t);
        t.inst = -1; t.nodeBase = t.primBase = 0;
        if (t.leafNext < t.leafEnd) return false;   // the rest of the scene-BVH leaf comes first
    }
    if (t.todoOffset == 0) { t.hit.t = t.ray.maxt; return true; }
    t.nodeNum = TRAV_POP(t);
#if PBRT_TOP_LEVELS
    t.heap = t.todoHeap[t.todoOffset];
#endif
    return false;
}

template <bool STATS, bool ANIM>
DEV int TravDescend(const DScene &sc, Trav &t, TraceStats *st, int maxSteps, const float4 *smemTop = nullptr) {
    for (int k = 0; k < maxSteps; ++k) {
        if (ANIM && t.inst < 0 && t.leafNext < t.leafEnd) {   // back from a TransformedPrimitive: the rest of that leaf
            t.leafFirst = t.leafNext; t.leafLast = t.leafEnd;
            t.leafNext = t.leafEnd = 0;
            return 1;
        }
        const uint32_t ni = (ANIM ? t.nodeBase : 0u) + t.nodeNum;
        float4 n0, n1;
#if PBRT_TOP_LEVELS
        if (smemTop && t.heap) { n0 = smemTop[2 * (t.heap - 1)]; n1 = smemTop[2 * (t.heap - 1) + 1]; }
        else
#endif
        {
            n0 = __ldg(sc.nodes + 2 * ni);
            n1 = __ldg(sc.nodes + 2 * ni + 1);
        }
        if (STATS) st->nodes++;
        if (BoxTest(n0, n1, t.ray, t.invDir, t.neg)) {
            const uint32_t offset = __float_as_uint(n1.z), meta = __float_as_uint(n1.w);
            const uint32_t nPrims = meta & 0xffu;
            if (nPrims > 0) {
                t.leafFirst = (ANIM ? t.primBase : 0u) + offset;
                t.leafLast = t.leafFirst + nPrims;
                return 1;
            }
            const uint32_t axis = (meta >> 8) & 0xffu;
#if PBRT_TOP_LEVELS
            const uint32_t kids = (t.heap && 2 * t.heap + 1 <= PBRT_TOP_SLOTS) ? 2 * t.heap : 0u;
#endif
            if (t.neg[axis]) {
#if PBRT_TOP_LEVELS
                t.todoHeap[t.todoOffset] = (uint16_t)kids;
                t.heap = kids ? kids + 1 : 0u;
#endif
                TRAV_PUSH(t, t.nodeNum + 1);
                t.nodeNum = offset;
            } else {
#if PBRT_TOP_LEVELS
                t.todoHeap[t.todoOffset] = (uint16_t)(kids ? kids + 1 : 0u);
                t.heap = kids;
#endif
                TRAV_PUSH(t, offset);
                t.nodeNum = t.nodeNum + 1;
            }
        } else if (TravPop<ANIM>(t))
            return 2;
    }
    return 0;
}

template <bool STATS, bool ANIM>
DEV bool TravLeaf(const DScene &sc, Trav &t, TraceStats *st) {
    const uint32_t end = t.leafLast;
    for (uint32_t pi = t.leafFirst; pi < end; ++pi) {
        const float4 r0 = __ldg(sc.prims + 3 * pi);
        const float4 r1 = __ldg(sc.prims + 3 * pi + 1);
        const float4 r2 = __ldg(sc.prims + 3 * pi + 2);
        if (STATS) st->prims++;
        const uint32_t type = __float_as_uint(r0.w) & 0xffu;
        float tt, b1 = 0.f, b2 = 0.f;
        bool h;
        if (type == PBRTGPU_PRIM_TRIANGLE)
            h = TriTest(V3(r0.x, r0.y, r0.z), V3(r1.x, r1.y, r1.z), V3(r2.x, r2.y, r2.z), t.ray, &tt, &b1, &b2);
        else if (ANIM && type == PBRTGPU_PRIM_ANIMATED) {
            // TransformedPrimitive::Intersect(P): interpolate, go to primitive space, walk the nested BVH
            const PbrtGpuAnimated *a = sc.animated + __float_as_uint(r1.w);
            DXf w2p;
            RayW2P(sc, __float_as_uint(r1.w), t.ray, &w2p);
            Ray rl;
            ToPrimitiveSpace(w2p, t.ray, &rl);
            if (a->nNodes == 0) {
                if (SinglePrimTest(sc, a->firstPrimitive, rl, t.anyHit, &t.hit)) {
                    t.found = true;
                    t.ray.maxt = rl.maxt;
                    if (t.anyHit) { t.hit.t = t.ray.maxt; return true; }
                }
                continue;
            }
            t.wo = t.ray.o; t.wd = t.ray.d;
            t.ray = rl;
            TravSetRay(t);
            t.inst = (int)__float_as_uint(r1.w);
            t.nodeBase = a->firstNode; t.primBase = a->firstPrimitive;
            t.todoBase = t.todoOffset;
            t.leafNext = pi + 1; t.leafEnd = end;
            if (t.leafNext == t.leafEnd) t.leafNext = t.leafEnd = 0;
            t.nodeNum = 0;
#if PBRT_TOP_LEVELS
            t.heap = 0;
#endif
            return false;   // (descend into the nested BVH next)
        } else
            h = QuadricIntersect(sc, sc.quadrics[__float_as_uint(r1.w)], t.ray, &tt, nullptr);
        if (h) {
            t.hit.prim = (int)pi;
            t.found = true;
            if (t.anyHit) { t.hit.t = t.ray.maxt; return true; }
            t.hit.b1 = b1; t.hit.b2 = b2; t.hit.maxtIn = t.ray.maxt;
            t.ray.maxt = tt;
        }
    }
    return TravPop<ANIM>(t);
}

// one node test (+ the leaf it leads to): the walk in single steps.  true when the ray is finished.
template <bool STATS, bool ANIM>
DEV bool TravStep(const DScene &sc, Trav &t, TraceStats *st, const float4 *smemTop = nullptr) {
#if PBRT_PAIR_NODES
    if (!STATS) return TravStepPairs<ANIM>(sc, t);
#endif
    const int r = TravDescend<STATS, ANIM>(sc, t, st, 1, smemTop);
    if (r == 2) return true;
    if (r == 1) return TravLeaf<STATS, ANIM>(sc, t, st);
    return false;
}

// Re-derive the winning primitive's DifferentialGeometry (what Shape::Intersect
// filled in at the moment the hit was accepted) and its rayEpsilon.
// Under a TransformedPrimitive the geometry is computed in primitive space with the transformed ray and then
// taken to world space (primitive.cpp:104-117); *w2pOut receives the interpolated transform (what
// isect->WorldToObject is multiplied by, primitive.cpp:106) and the return value says whether it applies.
template <bool ANIM>
DEV_NI3 bool HitGeometry(const DScene &sc, const Ray &rayIn, const Hit &hit, DiffGeom *dg, float *rayEpsilon,
                     int *material, int *areaLight, uint4 *triOut, uint32_t *primType, DXf *w2p) {
    const float4 r0 = __ldg(sc.prims + 3 * hit.prim);
    const float4 r1 = __ldg(sc.prims + 3 * hit.prim + 1);
    const float4 r2 = __ldg(sc.prims + 3 * hit.prim + 2);
    const uint32_t tw = __float_as_uint(r0.w);
    const uint32_t type = tw & 0xffu, index = __float_as_uint(r1.w), ml = __float_as_uint(r2.w);
    const int owner = ANIM ? (int)(tw >> 8) - 1 : -1;
    *material = (int)(ml & 0xffffu);
    *areaLight = (int)(ml >> 16) - 1;
    *primType = type;
    Ray ray = rayIn;
    if (ANIM && owner >= 0) {
        RayW2P(sc, (uint32_t)owner, rayIn, w2p);
        ToPrimitiveSpace(*w2p, rayIn, &ray);
    }
    if (type == PBRTGPU_PRIM_TRIANGLE) {
        const uint4 tri = __ldg(sc.tris + index);
        *triOut = tri;
        TriDG(sc, tri, V3(r0.x, r0.y, r0.z), V3(r1.x, r1.y, r1.z), V3(r2.x, r2.y, r2.z), ray, hit.t, hit.b1, hit.b2, dg);
        *rayEpsilon = 1e-3f * hit.t;   // trianglemesh.cpp:206
    } else {
        ray.maxt = hit.maxtIn;
        float t;
        QuadricIntersect(sc, sc.quadrics[index], ray, &t, dg);
        *rayEpsilon = 5e-4f * t;       // sphere.cpp:143, disk.cpp:100
        *triOut = make_uint4(0, 0, 0, 0);
    }
    if (ANIM && owner >= 0 && !MatIsIdentity(w2p->m)) {
        // PrimitiveToWorld = Inverse(w2p): points and vectors by w2p.mInv, normals by the transpose of w2p.m
        dg->p = XfPoint(w2p->mInv, dg->p);
        dg->nn = Normalize(XfNormal(w2p->m, dg->nn));
        dg->dpdu = XfVector(w2p->mInv, dg->dpdu);
        dg->dpdv = XfVector(w2p->mInv, dg->dpdv);
        return true;
    }
    return false;
}

}  // namespace dev
