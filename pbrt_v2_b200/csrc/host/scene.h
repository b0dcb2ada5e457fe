// scene.h -- host-side mirror of the pbrt-v2 class surface that the path-tracing
// hot path crosses.  Only the classes that sit on the path exist here, with the
// reference's names and constructor parameters, so that a maintainer recognises
// the drop-in pair:
//
//   Renderer / GpuSamplerRenderer      core/renderer.h:43-54, renderers/samplerrenderer.h
//   SurfaceIntegrator / GpuPathIntegrator   core/integrator.h:51-70, integrators/path.h:44-63
//   Aggregate / GpuBVHAccel            core/primitive.h:126-134, accelerators/bvh.h:49-76
//   Shape / TriangleMesh / Sphere / Disk / LoopSubdiv   core/shape.h, shapes/*.h
//   Material (matte / plastic / measured), Light (point / diffuse area)
//   PerspectiveCamera, ImageFilm, LDSampler (parameters only -- the work they do
//   per sample happens on the GPU behind include/pbrtgpu.h)
//
// Everything per-ray / per-sample (Intersect, Li, GetBSDF, Sample_f, AddSample)
// is NOT implemented on the host: it lives in the CUDA library.  The host
// builds the scene objects, builds the BVH with the reference's algorithm (so
// node order and therefore traversal order match), flattens them into
// PbrtGpuSceneDesc and calls pbrtgpu_render().
#ifndef PBRT_B200_HOST_SCENE_H
#define PBRT_B200_HOST_SCENE_H

#include <map>
#include <memory>
#include <string>
#include <vector>
#include <cstdint>

#include "geom.h"
#include "../../../include/pbrtgpu.h"

namespace pb {

// ---------------------------------------------------------------- ParamSet --
// core/paramset.h:46-118 -- typed name -> array bags filled by the parser.
struct ParamSet {
    std::map<std::string, std::vector<int>> ints;
    std::map<std::string, std::vector<float>> floats;
    std::map<std::string, std::vector<bool>> bools;
    std::map<std::string, std::vector<V3>> points, vectors, normals;
    std::map<std::string, std::vector<V3>> spectra;   // RGBSpectrum (core/pbrt.h:156)
    std::map<std::string, std::vector<std::string>> strings;
    std::map<std::string, std::string> textures;

    float FindOneFloat(const std::string &n, float d) const;
    int FindOneInt(const std::string &n, int d) const;
    bool FindOneBool(const std::string &n, bool d) const;
    V3 FindOnePoint(const std::string &n, const V3 &d) const;
    V3 FindOneSpectrum(const std::string &n, const V3 &d) const;
    bool HasSpectrum(const std::string &n) const { return spectra.count(n) != 0; }
    std::string FindOneString(const std::string &n, const std::string &d) const;
    const std::vector<float> *FindFloat(const std::string &n) const;
    const std::vector<int> *FindInt(const std::string &n) const;
    const std::vector<V3> *FindPoint(const std::string &n) const;
    const std::vector<V3> *FindNormal(const std::string &n) const;
};

// ------------------------------------------------------------------ Shapes --
struct Shape {   // core/shape.h:47-83
    enum Kind { TRIANGLE_MESH, SPHERE, DISK, LOOP_SUBDIV };
    Kind kind;
    const Transform *ObjectToWorld, *WorldToObject;   // owned by the TransformCache
    bool ReverseOrientation, TransformSwapsHandedness;
    Shape(Kind k, const Transform *o2w, const Transform *w2o, bool ro)
        : kind(k), ObjectToWorld(o2w), WorldToObject(w2o), ReverseOrientation(ro),
          TransformSwapsHandedness(o2w->SwapsHandedness()) {}
    virtual ~Shape() {}
    virtual bool CanIntersect() const { return true; }
    virtual float Area() const { return 0.f; }
};

struct TriangleMesh : Shape {   // shapes/trianglemesh.h:45-67, trianglemesh.cpp:42-71
    int ntris, nverts;
    std::vector<int> vertexIndex;
    std::vector<V3> p;    // WORLD space (transformed at construction)
    std::vector<V3> n;    // object space, empty if absent
    std::vector<float> uvs;
    TriangleMesh(const Transform *o2w, const Transform *w2o, bool ro, int nt, int nv, const int *vi,
                 const V3 *P, const V3 *N, const float *uv);
    bool CanIntersect() const override { return false; }
};
TriangleMesh *CreateTriangleMeshShape(const Transform *o2w, const Transform *w2o, bool ro,
                                      const ParamSet &params);

struct Sphere : Shape {   // shapes/sphere.h, sphere.cpp:39-49
    float radius, phiMax, zmin, zmax, thetaMin, thetaMax;
    Sphere(const Transform *o2w, const Transform *w2o, bool ro, float rad, float z0, float z1, float pm);
    BBox ObjectBound() const;
    float Area() const override;
};
struct Disk : Shape {   // shapes/disk.h, disk.cpp:39-46
    float height, radius, innerRadius, phiMax;
    Disk(const Transform *o2w, const Transform *w2o, bool ro, float ht, float r, float ri, float tmax);
    BBox ObjectBound() const;
    float Area() const override;
};
struct LoopSubdiv : Shape {   // shapes/loopsubdiv.h -- refined to a TriangleMesh with normals
    int nLevels;
    std::vector<int> indices;
    std::vector<V3> P;
    LoopSubdiv(const Transform *o2w, const Transform *w2o, bool ro, int nfaces, int nvertices,
               const int *vi, const V3 *P, int nl);
    bool CanIntersect() const override { return false; }
    TriangleMesh *Refine() const;   // shapes/loopsubdiv.cpp:240-446; null (and refineError set) if the device refine failed
    bool refineOnDevice = false;    // "string refiner" "gpu" / PBRTGPU_SUBDIV=gpu
    mutable std::string refineError;
    // the constructor's topology pass (loopsubdiv.cpp:156-216) as PbrtGpuSubdivMesh's index arrays
    void Topology(std::vector<int32_t> *startFace, std::vector<uint8_t> *vflags, std::vector<int32_t> *faceV,
                  std::vector<int32_t> *faceF) const;
    TriangleMesh *RefineOnDevice(std::string *err) const;
};

// --------------------------------------------------------------- Materials --
struct Material {   // materials/{matte,plastic,measured}.h -- constant textures only
    PbrtGpuMaterial desc;
    std::string measuredFile;
};
struct MeasuredData {   // materials/measured.cpp:85-135 + core/kdtree.h:101-147
    std::vector<PbrtGpuKdNode> nodes;
    std::vector<PbrtGpuBrdfSample> data;
};
bool ReadSpectrumFile(const std::string &filename, std::vector<float> *values);           // ReadFloatFile, core/floatfile.cpp:38-83
void SortedSamplesToRGB(const float *lambda, const float *v, int n, float rgb[3]);        // RGBSpectrum::FromSampled, sorted samples
bool LoadIrregIsotropicBRDF(const std::string &filename, MeasuredData *out, std::string *err);

// ------------------------------------------------------------------ Lights --
// InfiniteAreaLight::radianceMap (level 0) and ::distribution as the constructor leaves them (envmap.cpp)
struct EnvMap {
    int width = 0, height = 0;   // level 0 of the MIPMap: powers of two
    std::vector<float> texels;   // 3 * width * height
    int nu = 0, nv = 0;          // the image file's resolution = the distribution's
    std::vector<float> func, cdf, funcInt, margCdf;
    float margFuncInt = 0.f;
};
bool ReadImagePFM(const std::string &filename, int *xres, int *yres, std::vector<float> *rgb, std::string *err);
bool ReadImageEXR(const std::string &filename, int *xres, int *yres, std::vector<float> *rgb, std::string *err);   // exr.cpp
bool WriteImageEXR(const std::string &name, const std::vector<float> &rgb, int xRes, int yRes, int totalXRes, int totalYRes,
                   int xOffset, int yOffset);   // exr.cpp
bool ReadImageTGA(const std::string &filename, int *xres, int *yres, std::vector<float> *rgb, std::string *err);   // tga.cpp
// ReadImage (core/imageio.cpp:58-81): by suffix, .pfm / .tga / .exr; anything else is an error
bool ReadImage(const std::string &filename, int *xres, int *yres, std::vector<float> *rgb, std::string *err);
void BuildEnvMap(const float L[3], int width, int height, const float *rgb, EnvMap *out);

struct Light {   // core/light.h:48-82; lights/point.h, lights/diffuse.h
    PbrtGpuLight desc;
    const Shape *shape;   // DiffuseAreaLight: the ShapeSet's (single) shape
    Transform lightToWorld;   // InfiniteAreaLight: Light::LightToWorld (core/light.h:57-60)
    std::shared_ptr<EnvMap> envMap;   // InfiniteAreaLight
    Light() : shape(nullptr) { memset(&desc, 0, sizeof(desc)); }
};

// -------------------------------------------------------------- Primitives --
struct TransformedPrimitive;
struct GeometricPrimitive {   // core/primitive.h:81-101 (or, with `animated` set, a TransformedPrimitive)
    std::shared_ptr<Shape> shape;
    int material;
    int areaLight;   // index into Scene::lights or -1
    std::shared_ptr<TransformedPrimitive> animated;
    GeometricPrimitive() : material(-1), areaLight(-1) {}
};

// refined, intersectable primitive handed to the accelerator
struct RefinedPrim {
    uint32_t type;        // PBRTGPU_PRIM_*
    const Shape *shape;   // TriangleMesh (for triangles) or quadric
    int tri;              // triangle number within the mesh
    int material, areaLight;
    const TransformedPrimitive *animated;   // PBRTGPU_PRIM_ANIMATED
    RefinedPrim() : type(0), shape(nullptr), tri(0), material(-1), areaLight(-1), animated(nullptr) {}
    BBox WorldBound() const;   // Triangle::WorldBound trianglemesh.cpp:118-124 / Shape::WorldBound /
                               // TransformedPrimitive::WorldBound primitive.h:116-118
};

// --------------------------------------------------------------- Aggregate --
struct Aggregate {   // core/primitive.h:126-134
    virtual ~Aggregate() {}
};

// Drop-in for BVHAccel (accelerators/bvh.h:49-76): same constructor parameters,
// same build (SAH / middle / equal) and the same flattened LinearBVHNode[]; the
// per-ray Intersect / IntersectP run on the GPU.
struct GpuBVHAccel : Aggregate {
    GpuBVHAccel(const std::vector<GeometricPrimitive> &prims, uint32_t maxPrims = 1,
                const std::string &splitMethod = "sah", const std::string &builder = "host");
    uint32_t maxPrimsInNode;
    int splitMethod;   // 0 middle, 1 equal counts, 2 sah
    std::vector<RefinedPrim> primitives;    // in leaf order (BVHAccel::primitives after the swap)
    std::vector<PbrtGpuBVHNode> nodes;      // depth-first, bvh.cpp:383-399
    std::vector<std::shared_ptr<Shape>> owned;   // meshes produced by refinement
    std::vector<float> buildInput;          // the build's input bounds (n x 6), kept only under g_keepBuildInput (test hook)
    std::string buildError;                 // set when the device build was asked for and failed (no silent fallback)
};
extern bool g_keepBuildInput;
// PBRTGPU_PREP_TIMING=1 in the environment: wall clock of the scene-preparation phases on stderr
struct PrepTimer {
    const char *what;
    double t0;
    explicit PrepTimer(const char *w);
    ~PrepTimer() { Stop(); }
    void Stop();
};
extern uint32_t g_bvhParallelMin;   // bvh.cpp: smallest build that forks its subtrees onto several threads
// recursiveBuild + flattenBVHTree on primitive bounds (n x 6 floats); order = BVHAccel::primitives as input indices
void BuildBVHOnHost(const float *bounds, uint32_t n, uint32_t maxPrimsInNode, int splitMethod, std::vector<PbrtGpuBVHNode> *nodes,
                    std::vector<uint32_t> *order);
GpuBVHAccel *CreateGpuBVHAccelerator(const std::vector<GeometricPrimitive> &prims, const ParamSet &ps);

// core/primitive.h:104-123 as pbrtShape builds it for a shape under an animated CTM (api.cpp:1011-1043):
// the shape is made with identity transforms, refined, and -- if that gave more than one primitive --
// wrapped in its own BVHAccel with the constructor's DEFAULT arguments (maxPrimsInNode 1, "sah").
struct TransformedPrimitive {
    std::shared_ptr<Shape> shape;             // keeps the shape alive: the refined primitives point into it
    std::shared_ptr<GpuBVHAccel> primitive;   // `nodes` is empty when there is a single refined primitive; shared by the
                                              // instances of one object (ObjectInstance, api.cpp:1114-1147)
    AnimatedTransform WorldToPrimitive;
    TransformedPrimitive(const GeometricPrimitive &base, const AnimatedTransform &w2p);
    TransformedPrimitive(const std::shared_ptr<GpuBVHAccel> &object, const AnimatedTransform &w2p) : primitive(object), WorldToPrimitive(w2p) {}
    BBox WorldBound() const;
};

// ------------------------------------------------------ Camera/Film/Sampler --
struct ImageFilm {   // film/image.h:46-88, image.cpp:40-51
    int xResolution, yResolution;
    float cropWindow[4];
    float filterXWidth, filterYWidth;
    float filterTable[256];   // Filter::Evaluate at the table positions (image.cpp:57-69); box: all 1
    std::string filename;
    int xPixelStart, yPixelStart, xPixelCount, yPixelCount;
    std::vector<PbrtGpuPixel> pixels;
    ImageFilm(int xres, int yres, float fxw, float fyw, const float crop[4], const std::string &fn);
    void GetSampleExtent(int *xs, int *xe, int *ys, int *ye) const;   // image.cpp:157-166
    bool WriteImage() const;                                            // image.cpp:178-213
    void ToRGB(std::vector<float> *rgb) const;
};
struct PerspectiveCamera {   // cameras/perspective.h, core/camera.h:62-82
    Transform CameraToWorld, RasterToCamera;
    float shutterOpen, shutterClose, lensRadius, focalDistance;
    ImageFilm *film;
    PerspectiveCamera(const Transform &cam2world, const float screenWindow[4], float sopen, float sclose,
                      float lensr, float focald, float fov, ImageFilm *film);
};
struct LDSampler {   // samplers/lowdiscrepancy.h
    int xPixelStart, xPixelEnd, yPixelStart, yPixelEnd, samplesPerPixel;
};

// ------------------------------------------------------------ Integrators --
struct Scene;
struct SurfaceIntegrator {   // core/integrator.h:51-70
    int kind;       // PBRTGPU_INTEGRATOR_*
    int maxDepth;
    SurfaceIntegrator(int k, int md) : kind(k), maxDepth(md) {}
    virtual ~SurfaceIntegrator() {}
    // Integrator::RequestSamples: appends this integrator's Add1D / Add2D calls (array sizes) in call order
    virtual void RequestSamples(const Scene &scene, std::vector<uint16_t> *n1D, std::vector<uint16_t> *n2D) const = 0;
};
// Drop-in for PathIntegrator (integrators/path.h:44-63): same parameter
// ("maxdepth", default 5), same RequestSamples layout (3 depths x {light 1D+2D,
// lightNum 1D, bsdf 1D+2D, path 1D+2D}); Li() runs on the GPU.
struct GpuPathIntegrator : SurfaceIntegrator {
    explicit GpuPathIntegrator(int md) : SurfaceIntegrator(PBRTGPU_INTEGRATOR_PATH, md) {}
    void RequestSamples(const Scene &scene, std::vector<uint16_t> *n1D, std::vector<uint16_t> *n2D) const override;
};
// Drop-in for DirectLightingIntegrator (integrators/directlighting.h:47-70): parameters "maxdepth" (5) and
// "strategy" ("all" | "one"), RequestSamples as directlighting.cpp:53-78 with LDSampler::RoundSize.
struct GpuDirectLightingIntegrator : SurfaceIntegrator {
    int strategy;   // PBRTGPU_STRATEGY_*
    GpuDirectLightingIntegrator(int st, int md) : SurfaceIntegrator(PBRTGPU_INTEGRATOR_DIRECTLIGHTING, md), strategy(st) {}
    void RequestSamples(const Scene &scene, std::vector<uint16_t> *n1D, std::vector<uint16_t> *n2D) const override;
};

// -------------------------------------------------------------------- Scene --
struct Scene {   // core/scene.h:45-70
    std::unique_ptr<GpuBVHAccel> aggregate;
    std::vector<Light> lights;
    std::vector<Material> materials;
    std::vector<MeasuredData> measured;
    std::vector<std::string> measuredNames;
    // flattened form (valid after Flatten())
    std::vector<PbrtGpuPrimitive> fPrims;
    std::vector<PbrtGpuTriangle> fTris;
    std::vector<PbrtGpuMesh> fMeshes;
    std::vector<float> fP, fN, fUV;
    std::vector<PbrtGpuQuadric> fQuadrics;
    std::vector<PbrtGpuXform> fXforms;
    std::vector<PbrtGpuMaterial> fMaterials;
    std::vector<PbrtGpuMeasured> fMeasured;
    std::vector<PbrtGpuKdNode> fKdNodes;
    std::vector<PbrtGpuBrdfSample> fKdData;
    std::vector<PbrtGpuLight> fLights;
    std::vector<PbrtGpuLightShape> fLightShapes;
    std::vector<PbrtGpuEnvMap> fEnvMaps;
    std::vector<float> fEnvData;
    std::vector<PbrtGpuBVHNode> fNodes;
    std::vector<PbrtGpuAnimated> fAnimated;
    PbrtGpuSceneDesc desc;
    bool hasN, hasUV;
    bool Flatten(std::string *err);
};

// ----------------------------------------------------------------- Renderer --
struct Renderer {   // core/renderer.h:43-54
    virtual ~Renderer() {}
    virtual bool Render(Scene *scene) = 0;
};
// Drop-in for SamplerRenderer (renderers/samplerrenderer.h): Render() computes
// nTasks exactly like samplerrenderer.cpp:203-205, makes one pbrtgpu_render()
// call and fills the ImageFilm; WriteImage() then works as in the reference.
struct GpuSamplerRenderer : Renderer {
    LDSampler sampler;
    std::unique_ptr<PerspectiveCamera> camera;
    std::unique_ptr<ImageFilm> film;
    std::unique_ptr<SurfaceIntegrator> surfaceIntegrator;
    std::vector<uint16_t> n1D, n2D;   // Sample::n1D / n2D after both integrators' RequestSamples (sampler.cpp:89-91)
    int nCores;           // PbrtOptions.nCores (--ncores); 0 = hardware
    bool writeImage;
    PbrtGpuStats lastStats;
    PbrtGpuComm *comm;    // several GPUs: the library's NCCL communicator (not owned) or null; rank 0 owns the image
    int commRank;
    GpuSamplerRenderer() : nCores(0), writeImage(true), comm(nullptr), commRank(0) { memset(&lastStats, 0, sizeof(lastStats)); }
    int TaskCount() const;                       // samplerrenderer.cpp:203-205
    void FillRenderDesc(PbrtGpuRenderDesc *rd) const;
    bool Render(Scene *scene) override;          // samplerrenderer.cpp:188-222
};

// ------------------------------------------------------------ scene loading --
struct Options {   // core/pbrt.h:126-138
    int nCores = 0;
    bool quiet = false;
    bool quickRender = false;   // --quick: film resolution / 4, one pixel sample, light samples / 4
    std::string imageFile;
};
struct LoadedScene {
    std::unique_ptr<Scene> scene;
    std::unique_ptr<GpuSamplerRenderer> renderer;
    std::string error;
};
// Parses a pbrt-v2 scene file (the subset of directives the config scenes use)
// up to WorldEnd and builds Scene + Renderer; does NOT render.
bool LoadSceneFile(const std::string &filename, const Options &opt, LoadedScene *out);

// BASELINE.json configs[4]: the synthetic random-triangle scene
struct SyntheticParams {
    long long nTriangles = 10000000;
    int xres = 4096, yres = 4096, spp = 512, maxDepth = 8;
    std::string outfile = "synthetic.pfm";
};
bool WriteSyntheticSceneFile(const std::string &filename, const SyntheticParams &sp, std::string *err);
bool LoadSyntheticScene(const SyntheticParams &sp, const Options &opt, LoadedScene *out);

}  // namespace pb
#endif
