// renderer.cpp -- host side of the drop-in Renderer: camera/film set-up,
// scene flattening into the PbrtGpuSceneDesc POD arrays, the single
// pbrtgpu_render() call, and ImageFilm::WriteImage (PFM).
#include "scene.h"
#include <unistd.h>
#include <zlib.h>
#include <cctype>
#include <cmath>

#include <cstdio>
#include <map>
#include <thread>

namespace pb {

// ---------------------------------------------------------------- ImageFilm --
ImageFilm::ImageFilm(int xres, int yres, float fxw, float fyw, const float crop[4], const std::string &fn)
    : xResolution(xres), yResolution(yres), filterXWidth(fxw), filterYWidth(fyw), filename(fn) {
    for (int i = 0; i < 256; ++i) filterTable[i] = 1.f;   // BoxFilter::Evaluate (filters/box.cpp:39-41); others: parser.cpp
    // film/image.cpp:40-51
    memcpy(cropWindow, crop, 4 * sizeof(float));
    xPixelStart = (int)ceilf(xResolution * cropWindow[0]);
    xPixelCount = std::max(1, (int)ceilf(xResolution * cropWindow[1]) - xPixelStart);
    yPixelStart = (int)ceilf(yResolution * cropWindow[2]);
    yPixelCount = std::max(1, (int)ceilf(yResolution * cropWindow[3]) - yPixelStart);
    pixels.assign((size_t)xPixelCount * yPixelCount, PbrtGpuPixel{{0, 0, 0}, 0});
}

void ImageFilm::GetSampleExtent(int *xstart, int *xend, int *ystart, int *yend) const {
    // film/image.cpp:157-166
    *xstart = (int)floorf(xPixelStart + 0.5f - filterXWidth);
    *xend = (int)ceilf(xPixelStart - 0.5f + xPixelCount + filterXWidth);
    *ystart = (int)floorf(yPixelStart + 0.5f - filterYWidth);
    *yend = (int)ceilf(yPixelStart - 0.5f + yPixelCount + filterYWidth);
}

void ImageFilm::ToRGB(std::vector<float> *rgbOut) const {   // film/image.cpp:178-205
    size_t nPix = (size_t)xPixelCount * yPixelCount;
    rgbOut->assign(3 * nPix, 0.f);
    float *rgb = rgbOut->data();
    for (size_t o = 0; o < nPix; ++o) {
        const float *xyz = pixels[o].Lxyz;
        rgb[3 * o] = 3.240479f * xyz[0] - 1.537150f * xyz[1] - 0.498535f * xyz[2];
        rgb[3 * o + 1] = -0.969256f * xyz[0] + 1.875991f * xyz[1] + 0.041556f * xyz[2];
        rgb[3 * o + 2] = 0.055648f * xyz[0] - 0.204043f * xyz[1] + 1.057311f * xyz[2];
        float weightSum = pixels[o].weightSum;
        if (weightSum != 0.f) {
            float invWt = 1.f / weightSum;
            rgb[3 * o] = std::max(0.f, rgb[3 * o] * invWt);
            rgb[3 * o + 1] = std::max(0.f, rgb[3 * o + 1] * invWt);
            rgb[3 * o + 2] = std::max(0.f, rgb[3 * o + 2] * invWt);
        }
        // splatXYZ is always zero on this path (only MetropolisRenderer splats)
    }
}

static bool WritePFM(const std::string &name, const std::vector<float> &rgb, int w, int h) {
    // PFM, bottom-to-top rows, little-endian (negative scale): core/imageio.cpp:419-455
    FILE *fp = fopen(name.c_str(), "wb");
    if (!fp) { fprintf(stderr, "Error: unable to open output file \"%s\"\n", name.c_str()); return false; }
    fprintf(fp, "PF\n%d %d\n-1.000000\n", w, h);
    for (int y = h - 1; y >= 0; --y)
        fwrite(&rgb[(size_t)3 * y * w], sizeof(float), (size_t)3 * w, fp);
    fclose(fp);
    return true;
}

static bool WriteTGA(const std::string &name, const std::vector<float> &rgb, int w, int h) {
    // WriteImageTGA (core/imageio.cpp:200-225): 24-bit uncompressed BGR, top-to-bottom rows, gamma 2.2, written the way
    // tga_write_bgr fills the header (core/targa.c:566-620)
    FILE *fp = fopen(name.c_str(), "wb");
    if (!fp) { fprintf(stderr, "Error: unable to open output file \"%s\"\n", name.c_str()); return false; }
    unsigned char hdr[18] = {0};
    hdr[2] = 2;                                      // TGA_IMAGE_TYPE_BGR
    hdr[12] = (unsigned char)(w & 0xff); hdr[13] = (unsigned char)((w >> 8) & 0xff);
    hdr[14] = (unsigned char)(h & 0xff); hdr[15] = (unsigned char)((h >> 8) & 0xff);
    hdr[16] = 24;
    hdr[17] = 0x20;                                  // TGA_T_TO_B_BIT
    fwrite(hdr, 1, sizeof(hdr), fp);
    std::vector<unsigned char> row((size_t)3 * w);
    auto toByte = [](float v) -> unsigned char {
        float b = 255.f * powf(v, 1.f / 2.2f);
        b = b < 0.f ? 0.f : (b > 255.f ? 255.f : b);   // Clamp (NaN from a negative base stays NaN in the reference: -> 0 here)
        return (unsigned char)(b == b ? b : 0.f);
    };
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            const float *p = &rgb[3 * ((size_t)y * w + x)];
            row[3 * x] = toByte(p[2]); row[3 * x + 1] = toByte(p[1]); row[3 * x + 2] = toByte(p[0]);
        }
        fwrite(row.data(), 1, row.size(), fp);
    }
    // the TGA 2.0 footer tga_write_to_FILE appends (targa.c:25-30, 556-558): no extension area, no developer directory
    static const char footer[26] = "\0\0\0\0\0\0\0\0TRUEVISION-XFILE.";
    fwrite(footer, 1, sizeof(footer), fp);
    fclose(fp);
    return true;
}

// WriteImage's .png branch (core/imageio.cpp:107-126): 8-bit, gamma 2.2, and -- as the reference stores them -- the blue
// value in the file's first channel and the red one in its third (stbi_write_png is handed B, G, R bytes as "RGB").  The
// file itself: signature, IHDR, one IDAT holding the zlib stream of the filter-0 scan lines, IEND; CRCs by zlib.
static bool WritePNG(const std::string &name, const std::vector<float> &rgb, int w, int h) {
    auto toByte = [](float v) -> unsigned char {
        float b = 255.f * powf(v, 1.f / 2.2f);
        b = b < 0.f ? 0.f : (b > 255.f ? 255.f : b);
        return (unsigned char)(b == b ? b : 0.f);
    };
    std::vector<unsigned char> raw((size_t)h * (1 + 3 * (size_t)w));
    for (int y = 0; y < h; ++y) {
        unsigned char *row = &raw[(size_t)y * (1 + 3 * (size_t)w)];
        row[0] = 0;   // filter type: none
        for (int x = 0; x < w; ++x) {
            const float *p = &rgb[3 * ((size_t)y * w + x)];
            row[1 + 3 * x] = toByte(p[2]); row[2 + 3 * x] = toByte(p[1]); row[3 + 3 * x] = toByte(p[0]);
        }
    }
    uLongf n = compressBound((uLong)raw.size());
    std::vector<unsigned char> z(n);
    if (compress(z.data(), &n, raw.data(), (uLong)raw.size()) != Z_OK) return false;
    FILE *fp = fopen(name.c_str(), "wb");
    if (!fp) { fprintf(stderr, "Error: Error writing PNG \"%s\"\n", name.c_str()); return false; }
    auto be32 = [](unsigned char *o, uint32_t v) { o[0] = (unsigned char)(v >> 24); o[1] = (unsigned char)(v >> 16); o[2] = (unsigned char)(v >> 8); o[3] = (unsigned char)v; };
    auto chunk = [&](const char *type, const unsigned char *data, uint32_t len) {
        unsigned char head[8];
        be32(head, len);
        memcpy(head + 4, type, 4);
        fwrite(head, 1, 8, fp);
        if (len) fwrite(data, 1, len, fp);
        uLong crc = crc32(0L, Z_NULL, 0);
        crc = crc32(crc, head + 4, 4);
        if (len) crc = crc32(crc, data, len);
        unsigned char c[4];
        be32(c, (uint32_t)crc);
        fwrite(c, 1, 4, fp);
    };
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    fwrite(sig, 1, 8, fp);
    unsigned char ihdr[13];
    be32(ihdr, (uint32_t)w); be32(ihdr + 4, (uint32_t)h);
    ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;   // 8 bits, colour type 2 (RGB), deflate, filter 0, no interlace
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", z.data(), (uint32_t)n);
    chunk("IEND", nullptr, 0);
    fclose(fp);
    return true;
}

bool ImageFilm::WriteImage() const {
    // ::WriteImage picks the format by the suffix (core/imageio.cpp:85-140): .exr (exr.cpp), .tga, .pfm, .png.  Any other
    // name is an error in the reference ("Can't determine image file type"); here the float image still goes to
    // "<name>.pfm", with a warning.
    std::vector<float> rgb;
    ToRGB(&rgb);
    std::string suffix = filename.size() >= 5 ? filename.substr(filename.size() - 4) : std::string();
    for (char &c : suffix) c = (char)tolower((unsigned char)c);
    if (suffix == ".pfm") return WritePFM(filename, rgb, xPixelCount, yPixelCount);
    if (suffix == ".tga") return WriteTGA(filename, rgb, xPixelCount, yPixelCount);
    if (suffix == ".png") return WritePNG(filename, rgb, xPixelCount, yPixelCount);
    if (suffix == ".exr") return WriteImageEXR(filename, rgb, xPixelCount, yPixelCount, xResolution, yResolution, xPixelStart, yPixelStart);
    const std::string alt = filename + ".pfm";
    fprintf(stderr, "Warning: image files are written as .exr, .tga, .pfm or .png; \"%s\" is written as \"%s\" (float RGB, lossless)\n",
            filename.c_str(), alt.c_str());
    return WritePFM(alt, rgb, xPixelCount, yPixelCount);
}

// -------------------------------------------------------- PerspectiveCamera --
PerspectiveCamera::PerspectiveCamera(const Transform &cam2world, const float screenWindow[4], float sopen,
                                     float sclose, float lensr, float focald, float fov, ImageFilm *f)
    : CameraToWorld(cam2world), shutterOpen(sopen), shutterClose(sclose), lensRadius(lensr),
      focalDistance(focald), film(f) {
    // ProjectiveCamera ctor, core/camera.cpp:83-102; Perspective(fov, 1e-2, 1000), perspective.cpp:44
    Transform CameraToScreen = Perspective(fov, 1e-2f, 1000.f);
    Transform ScreenToRaster = Scale(float(film->xResolution), float(film->yResolution), 1.f) *
                               Scale(1.f / (screenWindow[1] - screenWindow[0]),
                                     1.f / (screenWindow[2] - screenWindow[3]), 1.f) *
                               Translate(V3(-screenWindow[0], -screenWindow[3], 0.f));
    Transform RasterToScreen = Inverse(ScreenToRaster);
    RasterToCamera = Inverse(CameraToScreen) * RasterToScreen;
}

// ------------------------------------------------------------------ Flatten --
bool Scene::Flatten(std::string *err) {
    PrepTimer timer("Flatten");
    const GpuBVHAccel &bvh = *aggregate;
    std::map<const Transform *, uint32_t> xformIndex;
    auto xformOf = [&](const Transform *t) -> uint32_t {
        auto it = xformIndex.find(t);
        if (it != xformIndex.end()) return it->second;
        PbrtGpuXform x;
        memcpy(x.m, t->m.m, sizeof(x.m));
        memcpy(x.mInv, t->mInv.m, sizeof(x.mInv));
        fXforms.push_back(x);
        xformIndex[t] = (uint32_t)fXforms.size() - 1;
        return (uint32_t)fXforms.size() - 1;
    };
    auto shapeFlags = [](const Shape *s) -> uint32_t {
        return (s->ReverseOrientation ? PBRTGPU_SHAPE_REVERSE_ORIENT : 0u) |
               (s->TransformSwapsHandedness ? PBRTGPU_SHAPE_SWAPS_HANDED : 0u);
    };
    std::map<const Shape *, uint32_t> meshIndex, quadricIndex;
    hasN = hasUV = false;
    // first pass: which optional vertex streams exist at all
    // the scene BVH first, then the nested BVH (or single primitive) of every TransformedPrimitive
    // One unit -- one copy of the nested aggregate's nodes and primitives -- per TransformedPrimitive, also for the
    // instances of one object (ObjectInstance), which share the aggregate on the host: a primitive record of the device
    // scene names the ONE TransformedPrimitive it lies under (HitGeometry finds the transform through it), so a shared
    // range cannot be expressed yet; pbrtgpu_scene_validate refuses overlapping ranges.  Costs memory, not correctness.
    std::vector<const GpuBVHAccel *> units(1, &bvh);
    std::map<const TransformedPrimitive *, uint32_t> animatedIndex;
    std::vector<uint32_t> unitOfAnimated;
    for (const RefinedPrim &rp : bvh.primitives)
        if (rp.type == PBRTGPU_PRIM_ANIMATED && !animatedIndex.count(rp.animated)) {
            animatedIndex[rp.animated] = (uint32_t)unitOfAnimated.size();
            unitOfAnimated.push_back((uint32_t)units.size());
            units.push_back(rp.animated->primitive.get());
        }
    for (const GpuBVHAccel *u : units)
        for (const RefinedPrim &rp : u->primitives)
            if (rp.type == PBRTGPU_PRIM_TRIANGLE) {
                const TriangleMesh *m = static_cast<const TriangleMesh *>(rp.shape);
                if (!m->n.empty()) hasN = true;
                if (!m->uvs.empty()) hasUV = true;
            }
    auto meshOf = [&](const TriangleMesh *m) -> uint32_t {
        auto it = meshIndex.find(m);
        if (it != meshIndex.end()) return it->second;
        PbrtGpuMesh gm;
        gm.flags = shapeFlags(m) | (m->n.empty() ? 0u : PBRTGPU_MESH_HAS_N) | (m->uvs.empty() ? 0u : PBRTGPU_MESH_HAS_UV);
        gm.xform = xformOf(m->ObjectToWorld);
        gm.firstVertex = (uint32_t)(fP.size() / 3);
        gm.nVertices = (uint32_t)m->nverts;
        for (int i = 0; i < m->nverts; ++i) {
            fP.push_back(m->p[i].x); fP.push_back(m->p[i].y); fP.push_back(m->p[i].z);
            if (hasN) {
                V3 nn = m->n.empty() ? V3(0, 0, 0) : m->n[i];
                fN.push_back(nn.x); fN.push_back(nn.y); fN.push_back(nn.z);
            }
            if (hasUV) {
                fUV.push_back(m->uvs.empty() ? 0.f : m->uvs[2 * i]);
                fUV.push_back(m->uvs.empty() ? 0.f : m->uvs[2 * i + 1]);
            }
        }
        fMeshes.push_back(gm);
        meshIndex[m] = (uint32_t)fMeshes.size() - 1;
        return (uint32_t)fMeshes.size() - 1;
    };
    auto quadricOf = [&](const Shape *s) -> uint32_t {
        auto it = quadricIndex.find(s);
        if (it != quadricIndex.end()) return it->second;
        PbrtGpuQuadric q;
        memset(&q, 0, sizeof(q));
        q.flags = shapeFlags(s);
        q.xform = xformOf(s->ObjectToWorld);
        if (s->kind == Shape::SPHERE) {
            const Sphere *sp = static_cast<const Sphere *>(s);
            q.type = PBRTGPU_PRIM_SPHERE;
            q.radius = sp->radius; q.zmin = sp->zmin; q.zmax = sp->zmax;
            q.thetaMin = sp->thetaMin; q.thetaMax = sp->thetaMax; q.phiMax = sp->phiMax;
        } else {
            const Disk *d = static_cast<const Disk *>(s);
            q.type = PBRTGPU_PRIM_DISK;
            q.radius = d->radius; q.height = d->height; q.innerRadius = d->innerRadius; q.phiMax = d->phiMax;
        }
        q.area = s->Area();
        fQuadrics.push_back(q);
        quadricIndex[s] = (uint32_t)fQuadrics.size() - 1;
        return (uint32_t)fQuadrics.size() - 1;
    };

    {   // the arrays' final sizes are known: no growth by doubling (at 10 M triangles that is gigabytes of copying)
        size_t nNodes = 0, nPrims = 0, nTris = 0, nVerts = 0;
        std::map<const Shape *, bool> seenMesh;
        for (const GpuBVHAccel *u : units) {
            nNodes += u->nodes.size();
            nPrims += u->primitives.size();
            for (const RefinedPrim &rp : u->primitives)
                if (rp.type == PBRTGPU_PRIM_TRIANGLE) {
                    ++nTris;
                    if (!seenMesh.count(rp.shape)) {
                        seenMesh[rp.shape] = true;
                        nVerts += (size_t)static_cast<const TriangleMesh *>(rp.shape)->nverts;
                    }
                }
        }
        fNodes.reserve(nNodes); fPrims.reserve(nPrims); fTris.reserve(nTris); fP.reserve(3 * nVerts);
        if (hasN) fN.reserve(3 * nVerts);
        if (hasUV) fUV.reserve(2 * nVerts);
    }
    fAnimated.resize(unitOfAnimated.size());
    for (size_t ui = 0; ui < units.size(); ++ui) {
    const GpuBVHAccel &unit = *units[ui];
    if (ui > 0)
        for (size_t ai = 0; ai < unitOfAnimated.size(); ++ai)
            if (unitOfAnimated[ai] == ui) {
                PbrtGpuAnimated &ga = fAnimated[ai];
                ga.firstNode = (uint32_t)fNodes.size();        ga.nNodes = (uint32_t)unit.nodes.size();
                ga.firstPrimitive = (uint32_t)fPrims.size();   ga.nPrimitives = (uint32_t)unit.primitives.size();
            }
    fNodes.insert(fNodes.end(), unit.nodes.begin(), unit.nodes.end());
    for (const RefinedPrim &rp : unit.primitives) {
        PbrtGpuPrimitive gp;
        gp.type = rp.type;
        gp.material = rp.material;
        gp.areaLight = rp.areaLight;
        if (rp.type == PBRTGPU_PRIM_ANIMATED) {
            if (ui > 0) { *err = "nested TransformedPrimitives are outside the hot-path scope"; return false; }
            const AnimatedTransform &at = rp.animated->WorldToPrimitive;
            gp.index = animatedIndex[rp.animated];
            PbrtGpuAnimated &ga = fAnimated[gp.index];
            memcpy(ga.start.m, at.startTransform.m.m, sizeof(ga.start.m));
            memcpy(ga.start.mInv, at.startTransform.mInv.m, sizeof(ga.start.mInv));
            memcpy(ga.end.m, at.endTransform.m.m, sizeof(ga.end.m));
            memcpy(ga.end.mInv, at.endTransform.mInv.m, sizeof(ga.end.mInv));
            ga.startTime = at.startTime; ga.endTime = at.endTime;
            ga.actuallyAnimated = at.actuallyAnimated ? 1u : 0u;
            for (int k = 0; k < 2; ++k) {
                ga.T[k][0] = at.T[k].x; ga.T[k][1] = at.T[k].y; ga.T[k][2] = at.T[k].z;
                ga.R[k][0] = at.R[k].v.x; ga.R[k][1] = at.R[k].v.y; ga.R[k][2] = at.R[k].v.z; ga.R[k][3] = at.R[k].w;
                memcpy(ga.S[k], at.S[k].m, sizeof(ga.S[k]));
            }
        } else if (rp.type == PBRTGPU_PRIM_TRIANGLE) {
            const TriangleMesh *m = static_cast<const TriangleMesh *>(rp.shape);
            uint32_t mi = meshOf(m);
            PbrtGpuTriangle t;
            uint32_t base = fMeshes[mi].firstVertex;
            t.v[0] = base + (uint32_t)m->vertexIndex[3 * rp.tri];
            t.v[1] = base + (uint32_t)m->vertexIndex[3 * rp.tri + 1];
            t.v[2] = base + (uint32_t)m->vertexIndex[3 * rp.tri + 2];
            t.mesh = mi;
            fTris.push_back(t);
            gp.index = (uint32_t)fTris.size() - 1;
        } else {
            gp.index = quadricOf(rp.shape);
        }
        fPrims.push_back(gp);
    }
    }
    for (const Material &m : materials) fMaterials.push_back(m.desc);
    for (const MeasuredData &md : measured) {
        PbrtGpuMeasured gm;
        gm.firstNode = (uint32_t)fKdNodes.size();
        gm.nNodes = (uint32_t)md.nodes.size();
        fKdNodes.insert(fKdNodes.end(), md.nodes.begin(), md.nodes.end());
        fKdData.insert(fKdData.end(), md.data.begin(), md.data.end());
        fMeasured.push_back(gm);
    }
    for (const Light &l : lights) {
        PbrtGpuLight gl = l.desc;
        if (gl.type == PBRTGPU_LIGHT_DIFFUSE_AREA) {
            // ShapeSet (core/light.cpp:114-139) over one intersectable shape
            PbrtGpuLightShape ls;
            ls.type = l.shape->kind == Shape::SPHERE ? PBRTGPU_PRIM_SPHERE : PBRTGPU_PRIM_DISK;
            ls.index = quadricOf(l.shape);
            ls.area = l.shape->Area();
            gl.firstShape = (uint32_t)fLightShapes.size();
            gl.nShapes = 1;
            float sumArea = 0.f;
            sumArea += ls.area;
            gl.sumArea = sumArea;
            fLightShapes.push_back(ls);
        } else if (gl.type == PBRTGPU_LIGHT_INFINITE) {
            gl.firstShape = xformOf(&l.lightToWorld);
            const EnvMap &em = *l.envMap;
            auto append = [&](const std::vector<float> &v) {
                const uint32_t at = (uint32_t)fEnvData.size();
                fEnvData.insert(fEnvData.end(), v.begin(), v.end());
                return at;
            };
            PbrtGpuEnvMap ge;
            ge.width = (uint32_t)em.width; ge.height = (uint32_t)em.height;
            ge.texels = append(em.texels);
            ge.nu = (uint32_t)em.nu; ge.nv = (uint32_t)em.nv;
            ge.func = append(em.func);
            ge.cdf = append(em.cdf);
            ge.funcInt = append(em.funcInt);
            ge.margCdf = append(em.margCdf);
            ge.margFuncInt = em.margFuncInt;
            gl.nShapes = (uint32_t)fEnvMaps.size();
            fEnvMaps.push_back(ge);
        }
        fLights.push_back(gl);
    }
    if (bvh.nodes.empty()) { *err = "scene has no primitives"; return false; }

    memset(&desc, 0, sizeof(desc));
    desc.nNodes = (uint32_t)fNodes.size();            desc.nodes = fNodes.data();
    desc.nPrimitives = (uint32_t)fPrims.size();       desc.primitives = fPrims.data();
    desc.nTriangles = (uint32_t)fTris.size();         desc.triangles = fTris.data();
    desc.nMeshes = (uint32_t)fMeshes.size();          desc.meshes = fMeshes.data();
    desc.nVertices = (uint32_t)(fP.size() / 3);       desc.P = fP.data();
    desc.N = hasN ? fN.data() : nullptr;
    desc.uv = hasUV ? fUV.data() : nullptr;
    desc.nQuadrics = (uint32_t)fQuadrics.size();      desc.quadrics = fQuadrics.data();
    desc.nXforms = (uint32_t)fXforms.size();          desc.xforms = fXforms.data();
    desc.nMaterials = (uint32_t)fMaterials.size();    desc.materials = fMaterials.data();
    desc.nMeasured = (uint32_t)fMeasured.size();      desc.measured = fMeasured.data();
    desc.nKdNodes = (uint32_t)fKdNodes.size();        desc.kdNodes = fKdNodes.data();
    desc.kdData = fKdData.data();
    desc.nLights = (uint32_t)fLights.size();          desc.lights = fLights.data();
    desc.nEnvMaps = (uint32_t)fEnvMaps.size();        desc.envMaps = fEnvMaps.empty() ? nullptr : fEnvMaps.data();
    desc.nEnvData = (uint32_t)fEnvData.size();        desc.envData = fEnvData.empty() ? nullptr : fEnvData.data();
    desc.nLightShapes = (uint32_t)fLightShapes.size(); desc.lightShapes = fLightShapes.data();
    desc.nAnimated = (uint32_t)fAnimated.size();      desc.animated = fAnimated.empty() ? nullptr : fAnimated.data();
    return true;
}

// -------------------------------------------------------------- Integrators --
void GpuPathIntegrator::RequestSamples(const Scene &, std::vector<uint16_t> *n1D, std::vector<uint16_t> *n2D) const {
    // path.cpp:41-49: for SAMPLE_DEPTH (3) depths: LightSampleOffsets(1), Add1D(1) lightNum, BSDFSampleOffsets(1),
    // BSDFSampleOffsets(1) for the path direction
    for (int i = 0; i < 3; ++i) {
        n1D->push_back(1); n2D->push_back(1);   // light: component, pos  (light.cpp:64-68)
        n1D->push_back(1);                      // lightNum
        n1D->push_back(1); n2D->push_back(1);   // bsdf: component, dir   (reflection.cpp:502-506)
        n1D->push_back(1); n2D->push_back(1);   // path: component, dir
    }
}

static uint16_t RoundSize(int size) {   // LDSampler::RoundSize = RoundUpPow2, lowdiscrepancy.h:56
    uint32_t v = (uint32_t)std::max(1, size);
    v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; v++;
    return (uint16_t)std::min<uint32_t>(v, 32768u);
}

void GpuDirectLightingIntegrator::RequestSamples(const Scene &scene, std::vector<uint16_t> *n1D,
                                                 std::vector<uint16_t> *n2D) const {   // directlighting.cpp:53-78
    if (strategy == PBRTGPU_STRATEGY_ALL) {
        for (const Light &l : scene.lights) {
            uint16_t n = RoundSize(l.desc.nSamples);
            n1D->push_back(n); n2D->push_back(n);   // LightSampleOffsets(nSamples, sample)
            n1D->push_back(n); n2D->push_back(n);   // BSDFSampleOffsets(nSamples, sample)
        }
    } else {
        n1D->push_back(1); n2D->push_back(1);       // LightSampleOffsets(1, sample)
        n1D->push_back(1);                          // lightNumOffset
        n1D->push_back(1); n2D->push_back(1);       // BSDFSampleOffsets(1, sample)
    }
}

// ------------------------------------------------------- GpuSamplerRenderer --
static int NumSystemCores(int nCores) {   // core/parallel.cpp:878-886
    if (nCores > 0) return nCores;
    // sysconf(_SC_NPROCESSORS_ONLN), as the reference: it ignores the affinity mask / cpuset, unlike
    // std::thread::hardware_concurrency(); nTasks -- and with it the image -- must not depend on taskset
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int GpuSamplerRenderer::TaskCount() const {   // samplerrenderer.cpp:203-205
    int nPixels = film->xResolution * film->yResolution;
    int nTasks = std::max(32 * NumSystemCores(nCores), nPixels / (16 * 16));
    uint32_t v = (uint32_t)nTasks;
    v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; v++;
    return (int)v;
}

void GpuSamplerRenderer::FillRenderDesc(PbrtGpuRenderDesc *rd) const {
    memset(rd, 0, sizeof(*rd));
    memcpy(rd->rasterToCamera, camera->RasterToCamera.m.m, sizeof(rd->rasterToCamera));
    memcpy(rd->cameraToWorld, camera->CameraToWorld.m.m, sizeof(rd->cameraToWorld));
    rd->lensRadius = camera->lensRadius;
    rd->focalDistance = camera->focalDistance;
    rd->shutterOpen = camera->shutterOpen;
    rd->shutterClose = camera->shutterClose;
    rd->xResolution = film->xResolution;
    rd->yResolution = film->yResolution;
    memcpy(rd->cropWindow, film->cropWindow, sizeof(rd->cropWindow));
    rd->hasFilterTable = 1;
    memcpy(rd->filterTable, film->filterTable, sizeof(rd->filterTable));
    rd->filterXWidth = film->filterXWidth;
    rd->filterYWidth = film->filterYWidth;
    rd->samplesPerPixel = sampler.samplesPerPixel;
    rd->maxDepth = surfaceIntegrator->maxDepth;
    rd->integrator = surfaceIntegrator->kind;
    rd->lightStrategy = surfaceIntegrator->kind == PBRTGPU_INTEGRATOR_DIRECTLIGHTING
                            ? static_cast<const GpuDirectLightingIntegrator *>(surfaceIntegrator.get())->strategy : 0;
    rd->n1D = (int32_t)std::min<size_t>(n1D.size(), PBRTGPU_MAX_SAMPLE_ARRAYS);
    rd->n2D = (int32_t)std::min<size_t>(n2D.size(), PBRTGPU_MAX_SAMPLE_ARRAYS);
    for (int i = 0; i < rd->n1D; ++i) rd->size1D[i] = n1D[i];
    for (int i = 0; i < rd->n2D; ++i) rd->size2D[i] = n2D[i];
    rd->nTasks = TaskCount();
    rd->taskBegin = 0;
    rd->taskEnd = rd->nTasks;
}

}  // namespace pb
