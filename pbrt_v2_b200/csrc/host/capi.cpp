// capi.cpp -- C entry points of libpbrthost.so, the host-side half of the
// drop-in (scene front end + BVH build + flattening).  Used by the Python
// tests/bench (ctypes) and by the `pbrt_b200` command line.  No compute happens
// here: rendering goes through include/pbrtgpu.h.
#include "scene.h"

#include <cmath>

#include <cstdio>
#include <cstring>

using namespace pb;

extern "C" {

struct PbrtHostScene {
    LoadedScene ls;
};

// ncores, quick: PbrtOptions.nCores (--ncores), PbrtOptions.quickRender (--quick), core/pbrt.h:126-138
PbrtHostScene *pbrthost_load_options(const char *filename, int ncores, int quick, char *err, int errlen) {
    PbrtHostScene *h = new PbrtHostScene;
    Options opt;
    opt.quickRender = quick != 0;
    opt.nCores = ncores;
    if (!LoadSceneFile(filename, opt, &h->ls)) {
        if (err && errlen > 0) snprintf(err, errlen, "%s", h->ls.error.c_str());
        delete h;
        return nullptr;
    }
    return h;
}

PbrtHostScene *pbrthost_load(const char *filename, int ncores, char *err, int errlen) {
    return pbrthost_load_options(filename, ncores, 0, err, errlen);
}

// the synthetic random-triangle scene of BASELINE.json configs[4], built without going through text
PbrtHostScene *pbrthost_load_synthetic(long long ntris, int xres, int yres, int spp, int maxdepth, int ncores,
                                       char *err, int errlen) {
    PbrtHostScene *h = new PbrtHostScene;
    Options opt;
    opt.nCores = ncores;
    SyntheticParams sp;
    sp.nTriangles = ntris; sp.xres = xres; sp.yres = yres; sp.spp = spp; sp.maxDepth = maxdepth;
    if (!LoadSyntheticScene(sp, opt, &h->ls)) {
        if (err && errlen > 0) snprintf(err, errlen, "%s", h->ls.error.c_str());
        delete h;
        return nullptr;
    }
    return h;
}

// the same scene as pbrt text (for the unmodified reference)
int pbrthost_write_synthetic(const char *filename, long long ntris, int xres, int yres, int spp, int maxdepth,
                             const char *outfile) {
    SyntheticParams sp;
    sp.nTriangles = ntris; sp.xres = xres; sp.yres = yres; sp.spp = spp; sp.maxDepth = maxdepth;
    sp.outfile = outfile;
    std::string err;
    return WriteSyntheticSceneFile(filename, sp, &err) ? 0 : 1;
}

void pbrthost_free(PbrtHostScene *h) { delete h; }

const PbrtGpuSceneDesc *pbrthost_scene_desc(PbrtHostScene *h) { return &h->ls.scene->desc; }

void pbrthost_render_desc(PbrtHostScene *h, PbrtGpuRenderDesc *rd) { h->ls.renderer->FillRenderDesc(rd); }

void pbrthost_film_extent(PbrtHostScene *h, int *xStart, int *yStart, int *xCount, int *yCount) {
    const ImageFilm *f = h->ls.renderer->film.get();
    *xStart = f->xPixelStart; *yStart = f->yPixelStart; *xCount = f->xPixelCount; *yCount = f->yPixelCount;
}

// bytes of the flattened scene arrays (= what pbrtgpu_scene_create copies host->device)
unsigned long long pbrthost_scene_bytes(PbrtHostScene *h) {
    const PbrtGpuSceneDesc &d = h->ls.scene->desc;
    unsigned long long b = 0;
    b += (unsigned long long)d.nNodes * sizeof(PbrtGpuBVHNode);
    b += (unsigned long long)d.nPrimitives * sizeof(PbrtGpuPrimitive);
    b += (unsigned long long)d.nTriangles * sizeof(PbrtGpuTriangle);
    b += (unsigned long long)d.nMeshes * sizeof(PbrtGpuMesh);
    b += (unsigned long long)d.nVertices * 12ull * (1 + (d.N ? 1 : 0)) + (d.uv ? (unsigned long long)d.nVertices * 8ull : 0);
    b += (unsigned long long)d.nQuadrics * sizeof(PbrtGpuQuadric) + (unsigned long long)d.nXforms * sizeof(PbrtGpuXform);
    b += (unsigned long long)d.nMaterials * sizeof(PbrtGpuMaterial) + (unsigned long long)d.nMeasured * sizeof(PbrtGpuMeasured);
    b += (unsigned long long)d.nKdNodes * (sizeof(PbrtGpuKdNode) + sizeof(PbrtGpuBrdfSample));
    b += (unsigned long long)d.nLights * sizeof(PbrtGpuLight) + (unsigned long long)d.nLightShapes * sizeof(PbrtGpuLightShape);
    b += (unsigned long long)d.nEnvMaps * sizeof(PbrtGpuEnvMap) + (unsigned long long)d.nEnvData * sizeof(float);
    return b;
}

void pbrthost_scene_counts(PbrtHostScene *h, unsigned *nNodes, unsigned *nPrims, unsigned *nTris, unsigned *nLights) {
    const PbrtGpuSceneDesc &d = h->ls.scene->desc;
    *nNodes = d.nNodes; *nPrims = d.nPrimitives; *nTris = d.nTriangles; *nLights = d.nLights;
}

// ImageFilm::WriteImage's XYZ->RGB / weight normalisation (film/image.cpp:178-205)
int pbrthost_pixels_to_rgb(PbrtHostScene *h, const PbrtGpuPixel *pixels, float *rgb) {
    ImageFilm *f = h->ls.renderer->film.get();
    size_t n = (size_t)f->xPixelCount * f->yPixelCount;
    memcpy(f->pixels.data(), pixels, n * sizeof(PbrtGpuPixel));
    std::vector<float> out;
    f->ToRGB(&out);
    memcpy(rgb, out.data(), out.size() * sizeof(float));
    return 0;
}

int pbrthost_write_image(PbrtHostScene *h, const PbrtGpuPixel *pixels, const char *filename) {
    ImageFilm *f = h->ls.renderer->film.get();
    size_t n = (size_t)f->xPixelCount * f->yPixelCount;
    memcpy(f->pixels.data(), pixels, n * sizeof(PbrtGpuPixel));
    if (filename && filename[0]) f->filename = filename;
    return f->WriteImage() ? 0 : 1;
}

// full Renderer::Render(scene) (GPU required)
int pbrthost_render(PbrtHostScene *h, int writeImage, PbrtGpuPixel *pixelsOut, PbrtGpuStats *stats) {
    GpuSamplerRenderer *r = h->ls.renderer.get();
    r->writeImage = writeImage != 0;
    if (!r->Render(h->ls.scene.get())) return 1;
    if (pixelsOut) memcpy(pixelsOut, r->film->pixels.data(), r->film->pixels.size() * sizeof(PbrtGpuPixel));
    if (stats) *stats = r->lastStats;
    return 0;
}

// test / bench hooks: the input of the top-level accelerator's build -- the world bounds of the refined primitives in the
// order BVHAccel receives them -- kept by scenes loaded after pbrthost_keep_build_input(1); returns the primitive count
void pbrthost_keep_build_input(int on) { g_keepBuildInput = on != 0; }
// builds of at least n primitives run their subtrees on several threads (default 131072; tests force it down to 2)
void pbrthost_set_bvh_parallel_min(unsigned n) { g_bvhParallelMin = n ? n : 1u; }
unsigned pbrthost_build_input(PbrtHostScene *h, float *out) {
    const GpuBVHAccel *a = h->ls.scene->aggregate.get();
    if (out && !a->buildInput.empty()) memcpy(out, a->buildInput.data(), a->buildInput.size() * sizeof(float));
    return (unsigned)(a->buildInput.size() / 6);
}
// ... and its flattened node array
void pbrthost_scene_nodes(PbrtHostScene *h, PbrtGpuBVHNode *out) {
    const GpuBVHAccel *a = h->ls.scene->aggregate.get();
    if (!a->nodes.empty()) memcpy(out, a->nodes.data(), a->nodes.size() * sizeof(PbrtGpuBVHNode));
}

// the host front end's BVH build (BVHAccel::recursiveBuild + flattenBVHTree as the reference runs them, bvh.cpp) on raw
// primitive bounds: same arguments and results as pbrtgpu_bvh_build, which it is the comparison partner and the timing
// baseline of.  nodes: capacity 2n - 1.
int pbrthost_bvh_build(const float *bounds, uint32_t n, uint32_t maxPrimsInNode, int splitMethod, PbrtGpuBVHNode *nodes,
                       uint32_t *nNodes, uint32_t *order) {
    if (splitMethod < 0 || splitMethod > 2 || maxPrimsInNode < 1 || maxPrimsInNode > 255) return 1;
    for (size_t i = 0; i < (size_t)6 * n; ++i)
        if (!std::isfinite(bounds[i])) return 1;   // (as pbrtgpu_bvh_build: the reference's bucket index is int(NaN) there)
    std::vector<PbrtGpuBVHNode> nv;
    std::vector<uint32_t> ov;
    BuildBVHOnHost(bounds, n, maxPrimsInNode, splitMethod, &nv, &ov);
    *nNodes = (uint32_t)nv.size();
    if (!nv.empty()) memcpy(nodes, nv.data(), nv.size() * sizeof(PbrtGpuBVHNode));
    if (!ov.empty()) memcpy(order, ov.data(), ov.size() * sizeof(uint32_t));
    return 0;
}

// InfiniteAreaLight's constructor on a .pfm file (NULL / "": no map, one texel = L) -> the arrays a PbrtGpuEnvMap points at,
// in the layout of `refdump envmap` (oracle/tools/refdump.cpp), its comparison partner: uint32 width, height, nu, nv; level-0
// texels; func; conditional cdfs; conditional funcInts; marginal cdf; marginal funcInt.  Returns the float count written
// after the 4-word head (out may be NULL to size the buffer), 0 on a read error.
unsigned long long pbrthost_envmap_build(const char *pfm, const float L[3], uint32_t head[4], float *out) {
    std::vector<float> rgb;
    int w = 1, h = 1;
    const bool haveMap = pfm && *pfm;
    std::string err;
    if (haveMap && !ReadImage(pfm, &w, &h, &rgb, &err)) return 0;
    EnvMap em;
    BuildEnvMap(L, w, h, haveMap ? rgb.data() : nullptr, &em);
    head[0] = (uint32_t)em.width; head[1] = (uint32_t)em.height; head[2] = (uint32_t)em.nu; head[3] = (uint32_t)em.nv;
    const std::vector<float> *parts[5] = {&em.texels, &em.func, &em.cdf, &em.funcInt, &em.margCdf};
    unsigned long long n = 0;
    for (const std::vector<float> *v : parts) {
        if (out) memcpy(out + n, v->data(), v->size() * sizeof(float));
        n += v->size();
    }
    if (out) out[n] = em.margFuncInt;
    return n + 1;
}

// ReadImage (core/imageio.cpp:58-81) for .pfm / .exr: RGB float triples, rows in the reference's order.  out NULL: sizes only.
int pbrthost_read_image(const char *filename, int *width, int *height, float *out) {
    std::vector<float> rgb;
    std::string err;
    if (!ReadImage(filename, width, height, &rgb, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    if (out) memcpy(out, rgb.data(), rgb.size() * sizeof(float));
    return 0;
}

// WriteImageEXR (core/imageio.cpp:170-193) on a raw float RGB image (data window = display window)
int pbrthost_write_exr(const char *filename, const float *rgb, int width, int height) {
    std::vector<float> v(rgb, rgb + (size_t)3 * width * height);
    return WriteImageEXR(filename, v, width, height, width, height, 0, 0) ? 0 : 1;
}

// LoopSubdiv on raw arrays (test hooks and the comparison partner of pbrtgpu_loopsubdiv_refine).
// topology: the constructor's pass (loopsubdiv.cpp:156-216) -> the index arrays of PbrtGpuSubdivMesh
int pbrthost_loopsubdiv_topology(int nVertices, int nFaces, const int *indices, const float *P, int32_t *startFace, uint8_t *vflags,
                                 int32_t *faceV, int32_t *faceF) {
    std::vector<V3> Pv(nVertices);
    for (int i = 0; i < nVertices; ++i) Pv[i] = V3(P[3 * i], P[3 * i + 1], P[3 * i + 2]);
    Transform identity;
    LoopSubdiv ls(&identity, &identity, false, nFaces, nVertices, indices, Pv.data(), 0);
    std::vector<int32_t> sf, fv, ff;
    std::vector<uint8_t> vf;
    ls.Topology(&sf, &vf, &fv, &ff);
    memcpy(startFace, sf.data(), sf.size() * 4);
    memcpy(vflags, vf.data(), vf.size());
    memcpy(faceV, fv.data(), fv.size() * 4);
    memcpy(faceF, ff.data(), ff.size() * 4);
    return 0;
}
// refine: the host front end's LoopSubdiv::Refine; first call with null outputs for the sizes
int pbrthost_loopsubdiv_refine(int nVertices, int nFaces, const int *indices, const float *P, int nLevels, uint32_t *nvOut,
                               uint32_t *ntOut, float *Pout, float *Nout, int32_t *indicesOut) {
    std::vector<V3> Pv(nVertices);
    for (int i = 0; i < nVertices; ++i) Pv[i] = V3(P[3 * i], P[3 * i + 1], P[3 * i + 2]);
    Transform identity;
    LoopSubdiv ls(&identity, &identity, false, nFaces, nVertices, indices, Pv.data(), nLevels);
    std::unique_ptr<TriangleMesh> m(ls.Refine());
    if (!m) return 1;
    *nvOut = (uint32_t)m->nverts;
    *ntOut = (uint32_t)m->ntris;
    if (!Pout) return 0;
    // TriangleMesh keeps P in world space (identity here) and N as given
    for (int i = 0; i < m->nverts; ++i) {
        Pout[3 * i] = m->p[i].x; Pout[3 * i + 1] = m->p[i].y; Pout[3 * i + 2] = m->p[i].z;
        Nout[3 * i] = m->n[i].x; Nout[3 * i + 1] = m->n[i].y; Nout[3 * i + 2] = m->n[i].z;
    }
    for (int i = 0; i < 3 * m->ntris; ++i) indicesOut[i] = m->vertexIndex[i];
    return 0;
}

}  // extern "C"
