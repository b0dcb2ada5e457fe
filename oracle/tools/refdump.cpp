// refdump.cpp -- TEST INFRASTRUCTURE.  OUR program, linked against the UNMODIFIED
// reference library (oracle/_ref/lib/libpbrt.a, built by oracle/Makefile from
// /root/reference/src), that prints known-answer vectors by calling the
// reference's own public classes.  tests/golden/make_golden.py turns its output
// into the committed fixtures.
//
//   refdump ld   <x0> <x1> <y0> <y1> <spp> <xres> <yres> <nTasks> <taskNum> <shutterOpen> <shutterClose>
//        first pixel of sub-sampler `taskNum`: spp rows of 37 floats
//        (imageX imageY lensU lensV time, 14 one-D, 9 two-D pairs) + the RNG draw that follows
//   refdump ldx  <the same 11 arguments> <n1D> <size>... <n2D> <size>...
//        the same for an arbitrary Sample layout (e.g. DirectLightingIntegrator's per-light arrays): the requests are
//        made through the reference's own Sample::Add1D / Add2D from a stub SurfaceIntegrator; every array's values
//        are printed in Sample::oneD / twoD order
//   refdump bxdf <exponent> <n> <seed>
//        the reference's own Lambertian / Microfacet(Blinn, FresnelDielectric(1.5, 1)) / FresnelDielectric classes on n
//        random direction pairs (tools/bsdftest.cpp exercises the same classes): per test
//        wo wi u1 u2 | lambert f, pdf | blinn f, pdf | blinn Sample_f wi, pdf, f | lambert Sample_f wi, pdf | Fr(cos)
//   refdump rng  <seed> <n>        n RandomUInt() values of RNG(seed)
//   refdump bvh  <bounds.bin> <maxPrimsInNode> <out.bin>
//        the reference's own BVHAccel (split method "sah") over n box primitives (bounds.bin: n x 6 float32 = pMin, pMax):
//        out.bin = uint32 nNodes, the LinearBVHNode array as it lies in memory (32 bytes per node, bvh.cpp:113-123), then
//        n uint32: the input index of the primitive at each position of the accelerator's reordered primitive list
//   refdump metal
//        the RGB coefficients of the default MetalMaterial's eta and k (copper) and its roughness, as hex floats
//   refdump envmap <map.pfm | -> <Lr> <Lg> <Lb> <out.bin>
//        the reference's own InfiniteAreaLight(identity, L, 1, map) ("-": no map): out.bin = uint32 width, height of
//        radianceMap (level 0), nu, nv of the distribution; 3*width*height float32 level-0 texels (Texel(0, s, t), row t
//        first); nu*nv func; nv*(nu+1) conditional cdf; nv conditional funcInt; nv+1 marginal cdf; 1 marginal funcInt
//   refdump loopsubdiv <mesh.bin> <nlevels> <out.bin>
//        the reference's own LoopSubdiv::Refine (mesh.bin: int32 nTriangles, nVertices, 3*nt int32 indices, 3*nv float32 P;
//        identity transform): out.bin = uint32 nv, nt, then 3*nv float32 P, 3*nv float32 N, 3*nt int32 indices of the
//        TriangleMesh it creates
#include <vector>
#include <map>
#include <set>
#include <string>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include "pbrt.h"
// Distribution1D / 2D keep their arrays private: this TOOL reads them where they lie (see the note further down)
#define private public
#include "montecarlo.h"
#undef private
#include "rng.h"
#include "sampler.h"
#include "samplers/lowdiscrepancy.h"
#include "integrators/path.h"
#include "integrators/emission.h"
#include "integrator.h"
#include "intersection.h"
#include "reflection.h"
#include "montecarlo.h"

#include <vector>
#include <map>
#include <set>
#include <string>
#include <algorithm>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "primitive.h"
#include "shape.h"
#include "paramset.h"
// the accelerator's node array and the refined mesh are private / protected members of the reference's classes: this TOOL
// reads them where they lie (the library it links is the unmodified build; access specifiers do not change a layout)
#define private public
#define protected public
#include "accelerators/bvh.h"
#include "shapes/trianglemesh.h"
#include "shapes/loopsubdiv.h"
#include "materials/metal.h"
#include "lights/infinite.h"
#undef private
#undef protected
#include "texture.h"

// a primitive that is nothing but its bounds: all BVHAccel's construction looks at (bvh.cpp:153-202)
class BoxPrimitive : public Primitive {
public:
    BoxPrimitive(const BBox &b, uint32_t i) : bounds(b), index(i) {}
    BBox WorldBound() const { return bounds; }
    bool Intersect(const Ray &, Intersection *) const { return false; }
    bool IntersectP(const Ray &) const { return false; }
    const AreaLight *GetAreaLight() const { return NULL; }
    BSDF *GetBSDF(const DifferentialGeometry &, const Transform &, MemoryArena &) const { return NULL; }
    BSSRDF *GetBSSRDF(const DifferentialGeometry &, const Transform &, MemoryArena &) const { return NULL; }
    BBox bounds;
    uint32_t index;
};
struct DumpedNode {   // LinearBVHNode, accelerators/bvh.cpp:113-123 (defined in the .cpp: restated here to size the dump)
    float bounds[6];
    uint32_t offset;
    uint8_t nPrimitives, axis, pad[2];
};
static uint32_t CountNodes(const DumpedNode *nodes, uint32_t i) {   // depth-first layout: a subtree is a contiguous range
    if (nodes[i].nPrimitives) return 1;
    const uint32_t left = CountNodes(nodes, i + 1);
    return 1 + left + CountNodes(nodes, nodes[i].offset);
}

// a SurfaceIntegrator that only requests sample arrays (what DirectLightingIntegrator::RequestSamples does through
// LightSampleOffsets / BSDFSampleOffsets, light.cpp:64-68, reflection.cpp:502-506)
class LayoutIntegrator : public SurfaceIntegrator {
public:
    std::vector<int> n1, n2;
    void RequestSamples(Sampler *, Sample *sample, const Scene *) {
        for (size_t i = 0; i < n1.size(); ++i) sample->Add1D(n1[i]);
        for (size_t i = 0; i < n2.size(); ++i) sample->Add2D(n2[i]);
    }
    Spectrum Li(const Scene *, const Renderer *, const RayDifferential &, const Intersection &, const Sample *, RNG &,
                MemoryArena &) const { return Spectrum(0.f); }
};

int main(int argc, char **argv) {
    if (argc >= 2 && !strcmp(argv[1], "metal")) {
        // the default MetalMaterial (materials/metal.cpp:101-112): RGBSpectrum::FromSampled of copper's n and k
        ParamSet geom, mat;
        std::map<string, Reference<Texture<float> > > ft;
        std::map<string, Reference<Texture<Spectrum> > > st;
        TextureParams tp(geom, mat, ft, st);
        Transform identity;
        MetalMaterial *m = CreateMetalMaterial(identity, tp);
        DifferentialGeometry dg;
        float rgb[3];
        m->eta->Evaluate(dg).ToRGB(rgb);
        printf("eta %a %a %a\n", rgb[0], rgb[1], rgb[2]);
        m->k->Evaluate(dg).ToRGB(rgb);
        printf("k %a %a %a\n", rgb[0], rgb[1], rgb[2]);
        printf("roughness %a\n", m->roughness->Evaluate(dg));
        return 0;
    }
    if (argc >= 7 && !strcmp(argv[1], "envmap")) {
        float rgb[3] = {(float)atof(argv[3]), (float)atof(argv[4]), (float)atof(argv[5])};
        InfiniteAreaLight light(Transform(), Spectrum::FromRGB(rgb), 1, strcmp(argv[2], "-") ? argv[2] : "");
        const MIPMap<RGBSpectrum> *m = light.radianceMap;
        const Distribution2D *d = light.distribution;
        FILE *f = fopen(argv[6], "wb");
        if (!f) return 1;
        const uint32_t w = m->Width(), h = m->Height();
        const uint32_t nu = (uint32_t)d->pConditionalV[0]->count, nv = (uint32_t)d->pMarginal->count;
        const uint32_t head[4] = {w, h, nu, nv};
        fwrite(head, 4, 4, f);
        for (uint32_t t = 0; t < h; ++t)
            for (uint32_t s = 0; s < w; ++s) {
                float c[3];
                m->Texel(0, (int)s, (int)t).ToRGB(c);
                fwrite(c, 4, 3, f);
            }
        for (uint32_t v = 0; v < nv; ++v) fwrite(d->pConditionalV[v]->func, 4, nu, f);
        for (uint32_t v = 0; v < nv; ++v) fwrite(d->pConditionalV[v]->cdf, 4, nu + 1, f);
        for (uint32_t v = 0; v < nv; ++v) fwrite(&d->pConditionalV[v]->funcInt, 4, 1, f);
        fwrite(d->pMarginal->cdf, 4, nv + 1, f);
        fwrite(&d->pMarginal->funcInt, 4, 1, f);
        fclose(f);
        return 0;
    }
    if (argc >= 5 && !strcmp(argv[1], "bvh")) {
        FILE *f = fopen(argv[2], "rb");
        if (!f) return 2;
        std::vector<float> b;
        float v[6];
        while (fread(v, sizeof(float), 6, f) == 6) b.insert(b.end(), v, v + 6);
        fclose(f);
        const uint32_t n = (uint32_t)(b.size() / 6);
        vector<Reference<Primitive> > prims;
        for (uint32_t i = 0; i < n; ++i)
            prims.push_back(new BoxPrimitive(BBox(Point(b[6 * i], b[6 * i + 1], b[6 * i + 2]), Point(b[6 * i + 3], b[6 * i + 4], b[6 * i + 5])), i));
        BVHAccel accel(prims, (uint32_t)atoi(argv[3]), "sah");
        static_assert(sizeof(DumpedNode) == 32, "LinearBVHNode is 32 bytes");
        const DumpedNode *nodes = reinterpret_cast<const DumpedNode *>(accel.nodes);
        const uint32_t nNodes = n ? CountNodes(nodes, 0) : 0;
        FILE *o = fopen(argv[4], "wb");
        if (!o) return 2;
        fwrite(&nNodes, 4, 1, o);
        fwrite(nodes, sizeof(DumpedNode), nNodes, o);
        for (size_t i = 0; i < accel.primitives.size(); ++i) {
            const uint32_t idx = static_cast<const BoxPrimitive *>(accel.primitives[i].GetPtr())->index;
            fwrite(&idx, 4, 1, o);
        }
        fclose(o);
        return 0;
    }
    if (argc >= 5 && !strcmp(argv[1], "loopsubdiv")) {
        FILE *f = fopen(argv[2], "rb");
        if (!f) return 2;
        int nt = 0, nv = 0;
        if (fread(&nt, 4, 1, f) != 1 || fread(&nv, 4, 1, f) != 1) return 2;
        std::vector<int> vi(3 * (size_t)nt);
        std::vector<float> P(3 * (size_t)nv);
        if (fread(vi.data(), 4, vi.size(), f) != vi.size() || fread(P.data(), 4, P.size(), f) != P.size()) return 2;
        fclose(f);
        std::vector<Point> pts(nv);
        for (int i = 0; i < nv; ++i) pts[i] = Point(P[3 * i], P[3 * i + 1], P[3 * i + 2]);
        Transform identity;
        LoopSubdiv subdiv(&identity, &identity, false, nt, nv, vi.data(), pts.data(), atoi(argv[3]));
        vector<Reference<Shape> > refined;
        subdiv.Refine(refined);
        if (refined.size() != 1) return 3;
        const TriangleMesh *mesh = static_cast<const TriangleMesh *>(refined[0].GetPtr());
        FILE *o = fopen(argv[4], "wb");
        if (!o) return 2;
        const uint32_t onv = (uint32_t)mesh->nverts, ont = (uint32_t)mesh->ntris;
        fwrite(&onv, 4, 1, o);
        fwrite(&ont, 4, 1, o);
        fwrite(mesh->p, sizeof(Point), onv, o);
        fwrite(mesh->n, sizeof(Normal), onv, o);
        fwrite(mesh->vertexIndex, sizeof(int), 3 * (size_t)ont, o);
        fclose(o);
        return 0;
    }
    if (argc >= 4 && !strcmp(argv[1], "rng")) {
        RNG rng((uint32_t)strtoul(argv[2], NULL, 10));
        int n = atoi(argv[3]);
        for (int i = 0; i < n; ++i) printf("%u\n", rng.RandomUInt());
        return 0;
    }
    if (argc >= 13 && !strcmp(argv[1], "ld")) {
        int x0 = atoi(argv[2]), x1 = atoi(argv[3]), y0 = atoi(argv[4]), y1 = atoi(argv[5]);
        int spp = atoi(argv[6]);
        int nTasks = atoi(argv[9]), taskNum = atoi(argv[10]);
        float so = (float)atof(argv[11]), sc = (float)atof(argv[12]);
        LDSampler mainSampler(x0, x1, y0, y1, spp, so, sc);
        PathIntegrator surf(5);
        EmissionIntegrator vol(1.f);
        Sample *origSample = new Sample(&mainSampler, &surf, &vol, NULL);
        Sampler *sub = mainSampler.GetSubSampler(taskNum, nTasks);
        if (!sub) { printf("EMPTY\n"); return 0; }
        RNG rng(taskNum);
        int maxSamples = sub->MaximumSampleCount();
        Sample *samples = origSample->Duplicate(maxSamples);
        int n = sub->GetMoreSamples(samples, rng);
        printf("%d %d %d\n", n, (int)origSample->n1D.size(), (int)origSample->n2D.size());
        for (int i = 0; i < n; ++i) {
            printf("%.9g %.9g %.9g %.9g %.9g", samples[i].imageX, samples[i].imageY, samples[i].lensU,
                   samples[i].lensV, samples[i].time);
            for (size_t j = 0; j < samples[i].n1D.size(); ++j) printf(" %.9g", samples[i].oneD[j][0]);
            for (size_t j = 0; j < samples[i].n2D.size(); ++j)
                printf(" %.9g %.9g", samples[i].twoD[j][0], samples[i].twoD[j][1]);
            printf("\n");
        }
        printf("%u\n", rng.RandomUInt());
        return 0;
    }
    if (argc >= 5 && !strcmp(argv[1], "bxdf")) {
        float exponent = (float)atof(argv[2]);
        int n = atoi(argv[3]);
        RNG rng((uint32_t)strtoul(argv[4], NULL, 10));
        Spectrum R(0.5f);
        Lambertian lambert(R);
        FresnelDielectric fresnel(1.5f, 1.f);
        Blinn blinn(exponent);
        Microfacet micro(R, &fresnel, &blinn);
        for (int i = 0; i < n; ++i) {
            // every random number is drawn in its own statement: no dependence on argument evaluation order
            float a = rng.RandomFloat(), b = rng.RandomFloat(), c = rng.RandomFloat(), d = rng.RandomFloat();
            float u1 = rng.RandomFloat(), u2 = rng.RandomFloat();
            Vector wo = UniformSampleSphere(a, b), wi = UniformSampleSphere(c, d);
            float rgb[3];
            printf("%a %a %a %a %a %a %a %a", wo.x, wo.y, wo.z, wi.x, wi.y, wi.z, u1, u2);
            lambert.f(wo, wi).ToRGB(rgb); printf(" %a %a", rgb[0], lambert.Pdf(wo, wi));
            micro.f(wo, wi).ToRGB(rgb);   printf(" %a %a", rgb[0], micro.Pdf(wo, wi));
            Vector ws; float pdf = 0.f;
            Spectrum f = micro.Sample_f(wo, &ws, u1, u2, &pdf);
            f.ToRGB(rgb); printf(" %a %a %a %a %a", ws.x, ws.y, ws.z, pdf, rgb[0]);
            pdf = 0.f;
            f = lambert.Sample_f(wo, &ws, u1, u2, &pdf);
            printf(" %a %a %a %a", ws.x, ws.y, ws.z, pdf);
            fresnel.Evaluate(wo.z).ToRGB(rgb); printf(" %a\n", rgb[0]);
        }
        return 0;
    }
    if (argc >= 15 && !strcmp(argv[1], "ldx")) {
        int x0 = atoi(argv[2]), x1 = atoi(argv[3]), y0 = atoi(argv[4]), y1 = atoi(argv[5]);
        int spp = atoi(argv[6]);
        int nTasks = atoi(argv[9]), taskNum = atoi(argv[10]);
        float so = (float)atof(argv[11]), sc = (float)atof(argv[12]);
        LayoutIntegrator surf;
        int a = 13;
        int c1 = atoi(argv[a++]);
        for (int i = 0; i < c1; ++i) surf.n1.push_back(atoi(argv[a++]));
        int c2 = atoi(argv[a++]);
        for (int i = 0; i < c2; ++i) surf.n2.push_back(atoi(argv[a++]));
        LDSampler mainSampler(x0, x1, y0, y1, spp, so, sc);
        Sample *origSample = new Sample(&mainSampler, &surf, NULL, NULL);
        Sampler *sub = mainSampler.GetSubSampler(taskNum, nTasks);
        if (!sub) { printf("EMPTY\n"); return 0; }
        RNG rng(taskNum);
        Sample *samples = origSample->Duplicate(sub->MaximumSampleCount());
        int n = sub->GetMoreSamples(samples, rng);
        printf("%d %d %d\n", n, (int)origSample->n1D.size(), (int)origSample->n2D.size());
        for (int i = 0; i < n; ++i) {
            printf("%.9g %.9g %.9g %.9g %.9g", samples[i].imageX, samples[i].imageY, samples[i].lensU,
                   samples[i].lensV, samples[i].time);
            for (size_t j = 0; j < samples[i].n1D.size(); ++j)
                for (uint32_t k = 0; k < samples[i].n1D[j]; ++k) printf(" %.9g", samples[i].oneD[j][k]);
            for (size_t j = 0; j < samples[i].n2D.size(); ++j)
                for (uint32_t k = 0; k < 2 * samples[i].n2D[j]; ++k) printf(" %.9g", samples[i].twoD[j][k]);
            printf("\n");
        }
        printf("%u\n", rng.RandomUInt());
        return 0;
    }
    fprintf(stderr, "usage: refdump ld ... | refdump ldx ... | refdump rng seed n\n");
    return 1;
}
