// dvertex.cuh -- one path VERTEX of PathIntegrator::Li (integrators/path.cpp:52-123) as straight-line code in two halves
// around the BSDF-evaluation site, for the vertex-synchronous kernels (k_vertex in pbrtgpu.cu) and the host-compiled
// test harness (tests/native/device_on_host.cpp).
//
// What happens at a vertex, in the reference's order:
//     L += T * Le                                   (bounce 0, or after a specular bounce; path.cpp:62-63)
//     L += T * UniformSampleOneLight(...)           (integrator.cpp:82-114 -> EstimateDirect :117-174)
//         light half:  Sample_L, f(wo, wi_l), shadow ray           -> Ld += f * Li * (|wi.n| * w / pdf)
//         BSDF half:   Sample_f -> wi_m, f(wo, wi_m), closest hit  -> Ld += f * Le * |wi.n| * w / pdf
//     BSDF::Sample_f -> wi_p, f(wo, wi_p); T *= f * |wi.n| / pdf; Russian roulette; next ray   (path.cpp:86-121)
// The three directions wi_l, wi_m, wi_p depend on the shading context and the vertex's random numbers only -- not on each
// other's BSDF values or ray results -- so VertexA chooses all three, the caller evaluates the (up to) three BSDF sums at
// ONE site, VertexB forms the three rays, and the caller traces them back to back.  Every float expression is the one
// path_body.inc has (the wavefront / DirectLightingIntegrator state machine); only the schedule differs: same operands,
// same operation order per quantity, same RNG draw accounting (7 + 3 + [1] draws per vertex from bounce 3 on).
#pragma once
#include "dshade.cuh"
#include "dpath.cuh"

namespace dev {

struct VertexRandoms {
    float lnU;            // UniformSampleOneLight: light choice
    float lsU0, lsU1;     // LightSample uPos (uComponent selects a shape of the ShapeSet: one shape in scope)
    float bsU0, bsU1, bsUc;   // its BSDFSample
    float pU0, pU1, pUc;  // the path's BSDFSample (path.cpp:86-90)
};

enum { VX_LIGHT = 0, VX_MIS = 1, VX_PATH = 2 };

struct Vertex {
    // ---- VertexA
    V3 p;                 // bsdf->dgShading.p: origin of all three rays
    float eps;            // isect.rayEpsilon
    int material;
    int lightNum;
    V3 evalWo;            // wo in the shading frame
    V3 evalWi[3];         // the three wi in the shading frame
    V3 wiW[3];            // ... in world space
    int evalFlags[3];
    bool need[3];         // a BSDF value is wanted for this direction
    Spec Li;              // light half: Sample_L's radiance
    float scalarL;        //   |wi.n| * weight / pdf  (|wi.n| / pdf for a delta light)
    float visMaxt;        //   the shadow segment's maxt
    float absM, weightM, pdfM;   // BSDF half: |wi.n|, power-heuristic weight, bsdfPdf
    Spec Lemit;           //   the light's L
    float absP, pdfP;     // path sample: |wi.n|, pdf
#if PBRT_SPECULAR
    bool specSampled;     // the path sample chose a specular component: its f is specF (not a sum of f())
    Spec specF;
#endif
    // ---- VertexB
    bool ray[3];          // trace this ray
    Spec pendS, pendM;    // what the shadow / MIS ray adds to Ld when it comes back free / on the light
    Spec Tnew;            // throughput after the vertex (valid if ray[VX_PATH])
};

// Part A: everything up to the BSDF evaluations.  `L` receives the emitted light; the context of the hit is in *ctx.
// Returns with v->need[] / evalWo / evalWi[] / evalFlags[] filled for the evaluation site.
template <bool ANIM>
DEV void VertexA(const DScene &sc, const Ray &ray, const Hit &hit, int bounces, bool afterSpecular, const VertexRandoms &u,
                 const Spec &T, Spec &L, ShadeCtx &ctx, Vertex *v, unsigned long long *shadedN) {
    BuildShadeCtx<ANIM>(sc, ray, hit, &ctx, shadedN);
    const V3 wo = -ray.d;
    if (bounces == 0 || afterSpecular) L = L + T * HitLe(sc, ctx.areaLight, ctx.ng, wo);   // path.cpp:62-63
    v->p = ctx.p; v->eps = ctx.rayEpsilon; v->material = ctx.material;
    v->need[0] = v->need[1] = v->need[2] = false;
    v->evalFlags[0] = v->evalFlags[1] = v->evalFlags[2] = 0;
    v->evalWo = WorldToLocal(ctx, wo);
    v->lightNum = 0;
    v->scalarL = 0.f; v->visMaxt = 0.f; v->Li = Spec(0.f);
    v->absM = v->weightM = v->pdfM = 0.f; v->Lemit = Spec(0.f);
    v->absP = v->pdfP = 0.f;
#if PBRT_SPECULAR
    v->specSampled = false; v->specF = Spec(0.f);
#endif
    BxSet bx;
    MaterialBxDFs(sc, ctx.material, &bx);
    const int nLights = (int)sc.nLights;
    if (nLights > 0) {
        // ---- UniformSampleOneLight (integrator.cpp:82-114)
        const int lightNum = mini(Floor2Int(u.lnU * nLights), nLights - 1);
        v->lightNum = lightNum;
        const PbrtGpuLight &light = sc.lights[lightNum];
        // ---- EstimateDirect, light-sampling half (integrator.cpp:123-142)
        Ray vis;
        float sPdf;
        V3 wiW;
        const Spec Li = LightSample_L(sc, light, ctx.p, ctx.rayEpsilon, u.lsU0, u.lsU1, &wiW, &sPdf, &vis);
        v->visMaxt = vis.maxt;   // vis.o == p, vis.mint == rayEpsilon, vis.d == wi bit for bit
        v->wiW[VX_LIGHT] = wiW;
        if (sPdf > 0.f && !Li.IsBlack()) {
            v->Li = Li;
            const V3 evalWi = WorldToLocal(ctx, wiW);
            int flags = BSDF_ALL & ~BSDF_SPECULAR;
            if (Dot(wiW, ctx.ng) * Dot(wo, ctx.ng) > 0) flags &= ~BSDF_TRANSMISSION;
            else flags &= ~BSDF_REFLECTION;
            v->evalWi[VX_LIGHT] = evalWi; v->evalFlags[VX_LIGHT] = flags; v->need[VX_LIGHT] = true;
            if (light.type == PBRTGPU_LIGHT_POINT)
                v->scalarL = AbsDot(wiW, ctx.nn) / sPdf;
            else {
                const float bsdfPdf = BSDF_PdfLocal(bx, v->evalWo, evalWi, BSDF_ALL & ~BSDF_SPECULAR);
                const float weight = PowerHeuristic(1, sPdf, 1, bsdfPdf);
                v->scalarL = AbsDot(wiW, ctx.nn) * weight / sPdf;
            }
        }
        // ---- EstimateDirect, BSDF-sampling half (integrator.cpp:145-172): the direction and its weight
        if (light.type != PBRTGPU_LIGHT_POINT) {
            V3 eWo, eWi, wiM;
            float pdfM;
            int sampledType, flagsM;
#if PBRT_SPECULAR
            Spec specUnused;
#endif
            if (BSDF_SampleDir(ctx, bx, wo, u.bsU0, u.bsU1, u.bsUc, BSDF_ALL & ~BSDF_SPECULAR, &eWo, &eWi, &wiM, &pdfM,
                               &sampledType, &flagsM
#if PBRT_SPECULAR
                               , &specUnused
#endif
                               )) {
                // (sampledType is never specular here: the flags exclude it)
                if (pdfM > 0.f) {
                    const float lightPdf = LightPdf(sc, light, ctx.p, wiM);
                    if (lightPdf != 0.f) {   // else `return Ld`
                        v->evalWi[VX_MIS] = eWi; v->evalFlags[VX_MIS] = flagsM; v->need[VX_MIS] = true;
                        v->wiW[VX_MIS] = wiM;
                        v->absM = AbsDot(wiM, ctx.nn);
                        v->weightM = PowerHeuristic(1, pdfM, 1, lightPdf);
                        v->pdfM = pdfM;
                        v->Lemit = Spec(light.L[0], light.L[1], light.L[2]);   // Li is Lemit or black (diffuse.h:51-53)
#if PBRT_INFINITE
                        // an infinite light contributes when the MIS ray ESCAPES: Li = light->Le(ray) (integrator.cpp:163-164)
                        if (light.type == PBRTGPU_LIGHT_INFINITE) v->Lemit = LightLe(sc, light, wiM);
#endif
                    }
                }
            }
        }
    }
    // ---- path.cpp:86-94: the continuation direction
    {
        V3 eWo, eWi, wiP;
        float pdfP;
        int sampledType, flagsP;
#if PBRT_SPECULAR
        Spec specF(0.f);
        if (BSDF_SampleDir(ctx, bx, wo, u.pU0, u.pU1, u.pUc, BSDF_ALL, &eWo, &eWi, &wiP, &pdfP, &sampledType, &flagsP, &specF)) {
            v->specSampled = (sampledType & BSDF_SPECULAR) != 0;
            v->specF = specF;
#else
        if (BSDF_SampleDir(ctx, bx, wo, u.pU0, u.pU1, u.pUc, BSDF_ALL, &eWo, &eWi, &wiP, &pdfP, &sampledType, &flagsP)) {
#endif
            v->evalWi[VX_PATH] = eWi; v->evalFlags[VX_PATH] = flagsP; v->need[VX_PATH] = true;
            v->wiW[VX_PATH] = wiP;
            v->absP = AbsDot(wiP, ctx.nn);
            v->pdfP = pdfP;
        }
    }
}

// The evaluation site: the sum of the matching BxDF::f (reflection.cpp:572-576, 620-623) for one direction.
// (vectors by value: a reference parameter of a non-inlined function forces the caller's whole Vertex into local memory)
DEV_NOINLINE Spec VertexEval(const DScene &sc, int material, V3 wo, V3 wi, int flags) {
    BxSet bx;
    MaterialBxDFs(sc, material, &bx);
    Spec f(0.f);
    for (int i = 0; i < bx.n; ++i)
        if (Matches(bx.type(i), flags)) f = f + BxF(sc, bx, i, wo, wi);
    return f;
}
// ... for a material without a measured BRDF (matte, plastic, mirror, glass): inlined, no kd-tree stacks in the caller's frame
DEV Spec VertexEvalAnalytic(const DScene &sc, int material, const V3 &wo, const V3 &wi, int flags) {
    BxSet bx;
    MaterialBxDFs(sc, material, &bx);
    Spec f(0.f);
    for (int i = 0; i < bx.n; ++i)
        if (Matches(bx.type(i), flags)) f = f + BxFAnalytic(bx, i, wo, wi);
    return f;
}

// Does the path sample with BSDF value `fp` lead to a continuation ray (maxDepth aside)?  VertexB's test, for the kernels
// that queue the ray before VertexB runs (kwave.cuh): the value is not black and, when `roulette` (bounces > 3), the path
// survives path.cpp:103-108 -- the same expressions on the same operands as in VertexB below.
DEV bool PathSampleAlive(const Spec &fp, const Spec &T, float absP, float pdfP, bool roulette, float rrU) {
    if (fp.IsBlack()) return false;
    if (roulette) {
        const Spec Tn = T * (fp * absP / pdfP);
        const float continueProbability = minf(.5f, Tn.y());
        if (rrU > continueProbability) return false;
    }
    return true;
}

// Part B: with the three BSDF values, the terms the rays carry and the new throughput.  `rrU` is the Russian-roulette
// draw (looked at only when bounces > 3 and the path sample's f is not black: *rrDrawn says whether it was consumed).
// Returns true when the path continues (ray[VX_PATH]); `bounces` is NOT advanced here.
DEV bool VertexB(const Spec f[3], int bounces, int maxDepth, const Spec &T, float rrU, Vertex *v, bool *rrDrawn) {
    v->ray[0] = v->ray[1] = v->ray[2] = false;
    v->pendS = Spec(0.f); v->pendM = Spec(0.f); v->Tnew = T;
    *rrDrawn = false;
    if (v->need[VX_LIGHT] && !f[VX_LIGHT].IsBlack()) {
        v->pendS = f[VX_LIGHT] * v->Li * v->scalarL;
        v->ray[VX_LIGHT] = true;
    }
    if (v->need[VX_MIS] && !f[VX_MIS].IsBlack()) {
        v->pendM = f[VX_MIS] * v->Lemit * v->absM * v->weightM / v->pdfM;
        v->ray[VX_MIS] = true;
    }
    if (v->need[VX_PATH]) {
#if PBRT_SPECULAR
        const Spec fp = v->specSampled ? v->specF : f[VX_PATH];
#else
        const Spec fp = f[VX_PATH];
#endif
        if (!fp.IsBlack()) {
            Spec Tn = T * (fp * v->absP / v->pdfP);
            bool alive = true;
            if (bounces > 3) {   // path.cpp:103-108
                const float continueProbability = minf(.5f, Tn.y());
                *rrDrawn = true;
                if (rrU > continueProbability) alive = false;
                else Tn = Tn / continueProbability;
            }
            v->Tnew = Tn;
            if (alive && bounces != maxDepth) v->ray[VX_PATH] = true;
        }
    }
    return v->ray[VX_PATH];
}

// EstimateDirect's `lightIsect.primitive->GetAreaLight() == light` + Le(-wi) test on the MIS ray's hit (integrator.cpp:160-165)
template <bool ANIM>
DEV bool MisHitIsLight(const DScene &sc, const Ray &ray, const Hit &hit, int lightNum) {
    DiffGeom dgl;
    float eps;
    int mat, al;
    uint4 tri;
    uint32_t ptype;
    DXf w2pl;
    HitGeometry<ANIM>(sc, ray, hit, &dgl, &eps, &mat, &al, &tri, &ptype, &w2pl);
    const PbrtGpuLight &light = sc.lights[lightNum];
    const bool emits = !Spec(light.L[0], light.L[1], light.L[2]).IsBlack();
    return al == lightNum && Dot(dgl.nn, -ray.d) > 0.f && emits;
}

// Does the MIS ray's outcome add its term (integrator.cpp:158-169)?  Area light: the ray hit that light, facing it; infinite
// light (PBRT_INFINITE builds): the ray escaped.
template <bool ANIM>
DEV bool MisRayContributes(const DScene &sc, const Ray &ray, const Hit &hit, bool found, int lightNum) {
#if PBRT_INFINITE
    if (sc.lights[lightNum].type == PBRTGPU_LIGHT_INFINITE) return !found;
#endif
    return found && MisHitIsLight<ANIM>(sc, ray, hit, lightNum);
}
// the random numbers of a vertex at depth `d` < 3 from the camera sample's table row (path.cpp:71-76, 86-88)
DEV void VertexRandomsFromTables(const uint16_t *row, const uint32_t *scr, int d, VertexRandoms *u) {
    u->lnU = VanDerCorput(row[3 + 4 * d + 1], scr[5 + 4 * d + 1]);
    const uint32_t pl = row[17 + 3 * d + 0];
    u->lsU0 = VanDerCorput(pl, scr[19 + 2 * (3 * d + 0)]);
    u->lsU1 = Sobol2(pl, scr[20 + 2 * (3 * d + 0)]);
    const uint32_t pb = row[17 + 3 * d + 1];
    u->bsU0 = VanDerCorput(pb, scr[19 + 2 * (3 * d + 1)]);
    u->bsU1 = Sobol2(pb, scr[20 + 2 * (3 * d + 1)]);
    u->bsUc = VanDerCorput(row[3 + 4 * d + 2], scr[5 + 4 * d + 2]);
    const uint32_t pp = row[17 + 3 * d + 2];
    u->pU0 = VanDerCorput(pp, scr[19 + 2 * (3 * d + 2)]);
    u->pU1 = Sobol2(pp, scr[20 + 2 * (3 * d + 2)]);
    u->pUc = VanDerCorput(row[3 + 4 * d + 3], scr[5 + 4 * d + 3]);
}

// LDPixelSample's camera sample + PerspectiveCamera::GenerateRayDifferential (perspective.cpp:81-138); returns the
// film-footprint bits of the sample (FilmFootprint)
DEV int CameraRay(const RenderParams &rp, const uint16_t *row, const uint32_t *scr, int xPos, int yPos, Ray *ray) {
    const uint32_t pi = row[0];
    const float imageX = xPos + VanDerCorput(pi, scr[0]);
    const float imageY = yPos + Sobol2(pi, scr[1]);
    const int code = FilmFootprint(rp, xPos, yPos, imageX, imageY);
    V3 Pcamera = XfPoint(rp.rasterToCamera, V3(imageX, imageY, 0));
    V3 dir = Normalize(V3(Pcamera.x, Pcamera.y, Pcamera.z));
    ray->o = V3(0, 0, 0); ray->d = dir; ray->mint = 0.f; ray->maxt = INF_F;
    ray->time = Lerp(VanDerCorput(row[2], scr[4]), rp.shutterOpen, rp.shutterClose);   // montecarlo.cpp:243
    if (rp.lensRadius > 0.f) {
        const uint32_t pl = row[1];
        float lensU, lensV;
        ConcentricSampleDisk(VanDerCorput(pl, scr[2]), Sobol2(pl, scr[3]), &lensU, &lensV);
        lensU *= rp.lensRadius;
        lensV *= rp.lensRadius;
        float ft = rp.focalDistance / ray->d.z;
        V3 Pfocus = RayAt(*ray, ft);
        ray->o = V3(lensU, lensV, 0.f);
        ray->d = Normalize(Pfocus - ray->o);
    }
    ray->o = XfPoint(rp.cameraToWorld, ray->o);
    ray->d = XfVector(rp.cameraToWorld, ray->d);
    return code;
}

}  // namespace dev
