// device_on_host.cpp -- TEST INFRASTRUCTURE.  The device headers' traversal and hit-geometry code (dscene.cuh, danim.cuh,
// dmath.cuh, refmath.h) compiled for the HOST by g++ (no FMA contraction, like the device build's -fmad=false), so that
// the CPU test stage can compare the product's own arithmetic with the oracle, bit for bit, without a GPU: a guard for
// edits made while no GPU is at hand.  It is never linked into the product; only tests/ load it.
//
// The shims below give the handful of CUDA intrinsics the headers use their plain C++ meaning.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>   // float4 / uint4 / make_float4; __device__ etc. expand to nothing for a host compiler

#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
using std::isnan;
using std::isinf;
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
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
static inline void __syncwarp(unsigned = 0xffffffffu) {}

#include "../../pbrt_v2_b200/csrc/device/dscene.cuh"
#include "../../pbrt_v2_b200/csrc/device/dshade.cuh"

using namespace dev;

namespace {
struct HostScene {
    std::vector<float4> nodes, prims;
    std::vector<uint4> tris;
    DScene d;
};
// the same packing as pbrtgpu_scene_create (device/pbrtgpu.cu): nodes -> 2 x float4, primitives -> 3 x float4
void Pack(const PbrtGpuSceneDesc *desc, HostScene *h) {
    h->nodes.resize(2 * (size_t)desc->nNodes);
    for (uint32_t i = 0; i < desc->nNodes; ++i) {
        const PbrtGpuBVHNode &n = desc->nodes[i];
        uint32_t meta = (uint32_t)n.nPrimitives | ((uint32_t)n.axis << 8);
        h->nodes[2 * i] = make_float4(n.bmin[0], n.bmin[1], n.bmin[2], n.bmax[0]);
        h->nodes[2 * i + 1] = make_float4(n.bmax[1], n.bmax[2], __uint_as_float(n.offset), __uint_as_float(meta));
    }
    std::vector<int> owner(desc->nPrimitives, -1);
    for (uint32_t k = 0; k < desc->nAnimated; ++k)
        for (uint32_t i = 0; i < desc->animated[k].nPrimitives; ++i) owner[desc->animated[k].firstPrimitive + i] = (int)k;
    h->prims.resize(3 * (size_t)desc->nPrimitives);
    for (uint32_t i = 0; i < desc->nPrimitives; ++i) {
        const PbrtGpuPrimitive &p = desc->primitives[i];
        float v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        uint32_t w0 = p.type, w1 = p.index, w2 = 0;
        if (p.type != PBRTGPU_PRIM_ANIMATED) {
            if (p.type == PBRTGPU_PRIM_TRIANGLE) {
                const PbrtGpuTriangle &t = desc->triangles[p.index];
                for (int k = 0; k < 3; ++k) memcpy(v + 3 * k, desc->P + 3 * (size_t)t.v[k], 12);
            }
            w0 = p.type | ((uint32_t)(owner[i] + 1) << 8);
            w2 = (uint32_t)p.material | ((uint32_t)(p.areaLight + 1) << 16);
        }
        h->prims[3 * i] = make_float4(v[0], v[1], v[2], __uint_as_float(w0));
        h->prims[3 * i + 1] = make_float4(v[3], v[4], v[5], __uint_as_float(w1));
        h->prims[3 * i + 2] = make_float4(v[6], v[7], v[8], __uint_as_float(w2));
    }
    h->tris.resize(desc->nTriangles);
    for (uint32_t i = 0; i < desc->nTriangles; ++i)
        h->tris[i] = make_uint4(desc->triangles[i].v[0], desc->triangles[i].v[1], desc->triangles[i].v[2], desc->triangles[i].mesh);
    DScene &d = h->d;
    memset(&d, 0, sizeof(d));
    d.nodes = h->nodes.data(); d.prims = h->prims.data(); d.tris = h->tris.data(); d.meshes = desc->meshes;
    d.N = desc->N; d.uv = desc->uv; d.quadrics = desc->quadrics; d.xforms = desc->xforms; d.materials = desc->materials;
    d.measured = desc->measured; d.kdNodes = desc->kdNodes; d.kdData = desc->kdData; d.lights = desc->lights;
    d.lightShapes = desc->lightShapes; d.animated = desc->animated;
    d.nNodes = desc->nNodes; d.nPrims = desc->nPrimitives; d.nLights = desc->nLights; d.nAnimated = desc->nAnimated;
}
}  // namespace

extern "C" {

// k_intersect<false>'s body (device/pbrtgpu.cu) on the host: BVHAccel::Intersect + the hit's differential geometry
int doh_intersect(const PbrtGpuSceneDesc *desc, const PbrtGpuRay *rays, int n, PbrtGpuHit *hits, int stepwise) {
    HostScene hs;
    Pack(desc, &hs);
    const DScene &sc = hs.d;
    for (int i = 0; i < n; ++i) {
        Ray ray;
        ray.o = V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]);
        ray.d = V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        ray.mint = rays[i].mint; ray.maxt = rays[i].maxt; ray.time = rays[i].time;
        Hit hit;
        TraceStats st;
        st.nodes = st.prims = 0;
        bool found;
        if (stepwise) {   // the wavefront's step-wise walk (k_trace)
            Trav t;
            TravInit(t, ray, false);
            while (!TravStep<false, true>(sc, t, &st)) {}
            found = t.found; hit = t.hit; ray.maxt = t.ray.maxt;
        } else
            found = Traverse<false, true>(sc, ray, false, &hit, &st);
        PbrtGpuHit h;
        memset(&h, 0, sizeof(h));
        h.t = ray.maxt;
        if (found) {
            DiffGeom dg;
            float eps;
            int mat, al;
            uint4 tri;
            uint32_t ptype;
            DXf w2p;
            HitGeometry<true>(sc, ray, hit, &dg, &eps, &mat, &al, &tri, &ptype, &w2p);
            h.hit = 1; h.primitive = (uint32_t)hit.prim; h.b1 = hit.b1; h.b2 = hit.b2;
            h.p[0] = dg.p.x; h.p[1] = dg.p.y; h.p[2] = dg.p.z;
            h.n[0] = dg.nn.x; h.n[1] = dg.nn.y; h.n[2] = dg.nn.z;
        }
        hits[i] = h;
    }
    return 0;
}

int doh_intersect_p(const PbrtGpuSceneDesc *desc, const PbrtGpuRay *rays, int n, uint8_t *occ, int stepwise) {
    HostScene hs;
    Pack(desc, &hs);
    const DScene &sc = hs.d;
    for (int i = 0; i < n; ++i) {
        Ray ray;
        ray.o = V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]);
        ray.d = V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        ray.mint = rays[i].mint; ray.maxt = rays[i].maxt; ray.time = rays[i].time;
        Hit hit;
        TraceStats st;
        st.nodes = st.prims = 0;
        if (stepwise) {
            Trav t;
            TravInit(t, ray, true);
            while (!TravStep<false, true>(sc, t, &st)) {}
            occ[i] = t.found ? 1 : 0;
        } else
            occ[i] = Traverse<false, true>(sc, ray, true, &hit, &st) ? 1 : 0;
    }
    return 0;
}

// the shading steps of path_body.inc at the first hit of each ray, laid out like oracle_shade_test's output
int doh_shade_test(const PbrtGpuSceneDesc *desc, const PbrtGpuRay *rays, int n, const float *u, float *out) {
    HostScene hs;
    Pack(desc, &hs);
    const DScene &sc = hs.d;
    const int nLights = (int)sc.nLights;
    for (int i = 0; i < n; ++i) {
        float *o = out + 36 * (size_t)i;
        for (int k = 0; k < 36; ++k) o[k] = 0.f;
        Ray ray;
        ray.o = V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]);
        ray.d = V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        ray.mint = rays[i].mint; ray.maxt = rays[i].maxt; ray.time = rays[i].time;
        Hit hit;
        TraceStats st;
        st.nodes = st.prims = 0;
        if (!Traverse<false, true>(sc, ray, false, &hit, &st)) continue;
        ShadeCtx ctx;
        BuildShadeCtx<true>(sc, ray, hit, &ctx, nullptr);
        const V3 wo = -ray.d;
        auto put = [&](int at, const V3 &v) { o[at] = v.x; o[at + 1] = v.y; o[at + 2] = v.z; };
        put(0, ctx.p); put(3, ctx.nn); put(6, ctx.sn); put(9, ctx.tn); put(12, ctx.ng);
        o[15] = ctx.rayEpsilon;
        if (nLights == 0) continue;
        const float *uu = u + 6 * (size_t)i;
        const int lightNum = mini(Floor2Int(uu[5] * nLights), nLights - 1);
        const PbrtGpuLight &light = sc.lights[lightNum];
        V3 wiW;
        float sPdf;
        Ray vis;
        Spec Li = LightSample_L(sc, light, ctx.p, ctx.rayEpsilon, uu[0], uu[1], &wiW, &sPdf, &vis);
        put(16, wiW); o[19] = sPdf; o[20] = Li.r; o[21] = Li.g; o[22] = Li.b; o[23] = vis.maxt;
        BxSet bx;
        MaterialBxDFs(sc, ctx.material, &bx);
        {   // BSDF::f (path_body.inc, the light-sampling half) and BSDF::Pdf
            V3 evalWi = WorldToLocal(ctx, wiW), evalWo = WorldToLocal(ctx, wo);
            int flags = BSDF_ALL & ~BSDF_SPECULAR;
            if (Dot(wiW, ctx.ng) * Dot(wo, ctx.ng) > 0) flags &= ~BSDF_TRANSMISSION;
            else flags &= ~BSDF_REFLECTION;
            Spec f(0.f);
            for (int k = 0; k < bx.n; ++k)
                if (Matches(bx.type[k], flags)) f = f + BxF(sc, bx, k, evalWo, evalWi);
            o[24] = f.r; o[25] = f.g; o[26] = f.b;
            o[27] = BSDF_PdfLocal(bx, evalWo, evalWi, BSDF_ALL & ~BSDF_SPECULAR);
        }
        {   // BSDF::Sample_f
            V3 evalWo, evalWi, wi2;
            float pdf2;
            int sampledType, evalFlags;
            if (BSDF_SampleDir(ctx, bx, wo, uu[2], uu[3], uu[4], BSDF_ALL, &evalWo, &evalWi, &wi2, &pdf2, &sampledType, &evalFlags)) {
                Spec f(0.f);
                for (int k = 0; k < bx.n; ++k)
                    if (Matches(bx.type[k], evalFlags)) f = f + BxF(sc, bx, k, evalWo, evalWi);
                o[28] = f.r; o[29] = f.g; o[30] = f.b;
                put(31, wi2); o[34] = pdf2;
                o[35] = LightPdf(sc, light, ctx.p, wi2);
            }
        }
    }
    return 0;
}

}  // extern "C"
