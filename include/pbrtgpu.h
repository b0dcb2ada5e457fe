/*
 * pbrtgpu.h -- C ABI of the B200 path-tracing hot path (libpbrtgpu.so).
 *
 * This is the drop-in boundary for pbrt-v2's per-pixel-sample path-tracing loop
 *   SamplerRenderer::Render      /root/reference/src/renderers/samplerrenderer.cpp:188-222
 *   SamplerRendererTask::Run     renderers/samplerrenderer.cpp:60-164
 *   PathIntegrator::Li           integrators/path.cpp:52-123
 *   BVHAccel::Intersect/P        accelerators/bvh.cpp:403-503
 *   ImageFilm::AddSample         film/image.cpp:77-137
 * pbrt-v2 has no FFI of its own (no DSO plugins, README.txt:55-61); its
 * "plugin surface" is the abstract classes Renderer (core/renderer.h:43-54),
 * SurfaceIntegrator (core/integrator.h:51-70) and Aggregate
 * (core/primitive.h:126-134).  The only batch-granular virtual on the path is
 * Renderer::Render(const Scene*), so that is where the host<->device boundary
 * sits: a host-side Renderer subclass flattens the scene objects into the POD
 * arrays below and makes ONE pbrtgpu_render() call (see INTEGRATION.md for the
 * reference-side binding).
 *
 * Conventions: plain pointers and sizes, caller owns every host buffer, the
 * library owns all device memory, every entry point returns 0 on success or a
 * PBRTGPU_E* code (never aborts; pbrtgpu_last_error() gives text), calls on
 * one scene handle must be serialised by the caller.  Matrices are row-major
 * m[4*row+col] exactly like Matrix4x4::m[row][col] (core/transform.h).
 */
#ifndef PBRTGPU_H
#define PBRTGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBRTGPU_OK            0
#define PBRTGPU_EINVAL        1   /* malformed descriptor / argument          */
#define PBRTGPU_EUNSUPPORTED  2   /* scene feature outside the hot-path scope */
#define PBRTGPU_ECUDA         3   /* CUDA runtime failure (text in last_error)*/
#define PBRTGPU_ENODEVICE     4   /* no CUDA device: there is NO CPU fallback */

/* ---- transforms: Transform{m, mInv}, core/transform.h:86-178 --------------- */
typedef struct PbrtGpuXform {
    float m[16];
    float mInv[16];
} PbrtGpuXform;

/* ---- LinearBVHNode, accelerators/bvh.cpp:113-123 (same 32-byte layout) ----- */
typedef struct PbrtGpuBVHNode {
    float    bmin[3];
    float    bmax[3];
    uint32_t offset;       /* leaf: primitivesOffset; interior: secondChildOffset */
    uint8_t  nPrimitives;  /* 0 -> interior node */
    uint8_t  axis;
    uint8_t  pad[2];
} PbrtGpuBVHNode;

/* ---- primitives, in BVHAccel::primitives (leaf) order ---------------------- */
enum { PBRTGPU_PRIM_TRIANGLE = 0, PBRTGPU_PRIM_SPHERE = 1, PBRTGPU_PRIM_DISK = 2,
       PBRTGPU_PRIM_ANIMATED = 3   /* TransformedPrimitive: index into animated[], material/areaLight unused */ };

/* GeometricPrimitive{shape, material, areaLight}, core/primitive.cpp:149-190 */
typedef struct PbrtGpuPrimitive {
    uint32_t type;       /* PBRTGPU_PRIM_*                                      */
    uint32_t index;      /* TRIANGLE: triangles[]; SPHERE/DISK: quadrics[]; ANIMATED: animated[] */
    int32_t  material;   /* into materials[]                                    */
    int32_t  areaLight;  /* into lights[] or -1 (GetAreaLight()==NULL)          */
} PbrtGpuPrimitive;

/* Triangle{mesh, v}, shapes/trianglemesh.h:70-109 */
typedef struct PbrtGpuTriangle {
    uint32_t v[3];       /* indices into P/N/uv (mesh base already added)       */
    uint32_t mesh;       /* into meshes[]                                       */
} PbrtGpuTriangle;

#define PBRTGPU_MESH_HAS_N            1u
#define PBRTGPU_MESH_HAS_UV           2u
#define PBRTGPU_SHAPE_REVERSE_ORIENT  4u   /* Shape::ReverseOrientation         */
#define PBRTGPU_SHAPE_SWAPS_HANDED    8u   /* Shape::TransformSwapsHandedness   */

/* TriangleMesh, shapes/trianglemesh.cpp:42-71: P is world space, N/uv as given */
typedef struct PbrtGpuMesh {
    uint32_t flags;
    uint32_t xform;        /* ObjectToWorld, into xforms[] (normals: :322-324)  */
    uint32_t firstVertex;
    uint32_t nVertices;
} PbrtGpuMesh;

/* Sphere (shapes/sphere.cpp:39-49) / Disk (shapes/disk.cpp:39-46) */
typedef struct PbrtGpuQuadric {
    uint32_t type;         /* PBRTGPU_PRIM_SPHERE | PBRTGPU_PRIM_DISK           */
    uint32_t flags;        /* PBRTGPU_SHAPE_*                                   */
    uint32_t xform;        /* ObjectToWorld (WorldToObject = its inverse)       */
    float radius;
    float zmin, zmax, thetaMin, thetaMax;   /* sphere */
    float phiMax;
    float height, innerRadius;               /* disk   */
    float area;            /* Shape::Area() evaluated on the host               */
} PbrtGpuQuadric;

/* TransformedPrimitive{primitive, WorldToPrimitive}, core/primitive.h:104-123, built by pbrtShape for a
 * shape under an animated CTM (core/api.cpp:1011-1043).  The nested primitive is either its own
 * BVHAccel(refined prims; default maxPrimsInNode 1, "sah") -- nNodes > 0, node and primitive offsets
 * inside those nodes are relative to firstNode / firstPrimitive -- or a single primitive (nNodes == 0).
 * start/end/T/R/S are AnimatedTransform's members (core/transform.h:321-326): the two
 * world-to-primitive transforms and their Decompose() outputs (transform.cpp:328-368).                */
typedef struct PbrtGpuAnimated {
    PbrtGpuXform start, end;
    float    startTime, endTime;     /* renderOptions->transformStartTime/EndTime (api.cpp:1031-1032) */
    uint32_t actuallyAnimated;
    float    T[2][3];
    float    R[2][4];                /* quaternion (v.x, v.y, v.z, w)                      */
    float    S[2][16];
    uint32_t firstNode, nNodes;
    uint32_t firstPrimitive, nPrimitives;
} PbrtGpuAnimated;

/* ---- materials (constant textures only), materials/{matte,plastic,measured} - */
enum { PBRTGPU_MAT_MATTE = 0, PBRTGPU_MAT_PLASTIC = 1, PBRTGPU_MAT_MEASURED = 2,
       /* SURVEY 8(f) rank 4, on the device since round 2 (both integrators).
        * MIRROR (materials/mirror.cpp): Ks = Kr.  GLASS (materials/glass.cpp): Ks = Kr, Kd = Kt, roughness = index. */
       PBRTGPU_MAT_MIRROR = 3, PBRTGPU_MAT_GLASS = 4,
       /* METAL (materials/metal.cpp): Microfacet(1, FresnelConductor(eta, k), Blinn(1 / roughness)); Kd = eta, Ks = k (RGB).
        * Host front end + oracle (pinned on the reference binary) + device code behind -DPBRT_METAL=1; the shipped
        * library answers PBRTGPU_EUNSUPPORTED until that code has run on a GPU. */
       PBRTGPU_MAT_METAL = 5 };
typedef struct PbrtGpuMaterial {
    uint32_t type;
    float Kd[3];
    float sigma;        /* matte: Oren-Nayar sigma; only 0 is in scope          */
    float Ks[3];
    float roughness;
    int32_t measured;   /* into measured[] (IrregIsotropicBRDF) or -1           */
} PbrtGpuMaterial;

/* KdTree<IrregIsotropicBRDFSample>, core/kdtree.h:44-60, reflection.h:44-50 */
typedef struct PbrtGpuKdNode {
    float    splitPos;
    uint32_t bits;      /* splitAxis:2 | hasLeftChild:1 (bit 2) | rightChild:29 (bits 3..31) */
} PbrtGpuKdNode;
typedef struct PbrtGpuBrdfSample {
    float p[3];
    float v[3];
} PbrtGpuBrdfSample;
typedef struct PbrtGpuMeasured {
    uint32_t firstNode;   /* into kdNodes[] / kdData[]                          */
    uint32_t nNodes;
} PbrtGpuMeasured;

/* ---- lights, lights/{point,diffuse}.cpp, core/light.cpp:114-171 ------------ */
enum { PBRTGPU_LIGHT_POINT = 0, PBRTGPU_LIGHT_DIFFUSE_AREA = 1,
       /* InfiniteAreaLight (lights/infinite.cpp): radiance map + sampling distribution in envMaps[] (a 1x1 map without
        * "mapname").  Host front end + oracle (pinned on the reference binary); the shipped kernels do not hold the code
        * (PBRT_INFINITE=0): pbrtgpu_scene_create answers PBRTGPU_EUNSUPPORTED. */
       PBRTGPU_LIGHT_INFINITE = 2 };
typedef struct PbrtGpuLight {
    uint32_t type;
    float    L[3];        /* POINT: Intensity; AREA: Lemit; INFINITE: L * scale  */
    float    pos[3];      /* POINT: lightPos (world)                            */
    uint32_t firstShape;  /* AREA: ShapeSet::shapes, into lightShapes[]; INFINITE: LightToWorld, into xforms[] */
    uint32_t nShapes;     /* AREA: in scope exactly 1 (sphere or disk); INFINITE: index into envMaps[] */
    float    sumArea;     /* ShapeSet::sumArea                                  */
    int32_t  nSamples;    /* Light::nSamples (core/light.h:53,78; "nsamples" of an area light); 0 reads as 1 */
} PbrtGpuLight;
/* InfiniteAreaLight's members as its constructor leaves them (lights/infinite.cpp:67-104): level 0 of the radiance MIPMap
 * (texels already times L * scale, resampled to a power of two like core/mipmap.h:118-168) and the Distribution2D over
 * luminance x sin(theta) of the filtered map (core/montecarlo.h:54-80, montecarlo.cpp:358-370).  All offsets into envData[]. */
typedef struct PbrtGpuEnvMap {
    uint32_t width, height;   /* MIPMap::Width() / Height(): pyramid[0], a power of two each            */
    uint32_t texels;          /* 3 * width * height floats, RGB, row t = 0 first (image file order)     */
    uint32_t nu, nv;          /* Distribution2D: the image file's own resolution                        */
    uint32_t func;            /* nu * nv: pConditionalV[v]->func[u]                                     */
    uint32_t cdf;             /* nv * (nu + 1): pConditionalV[v]->cdf                                   */
    uint32_t funcInt;         /* nv: pConditionalV[v]->funcInt == pMarginal->func[v]                    */
    uint32_t margCdf;         /* nv + 1: pMarginal->cdf                                                 */
    float    margFuncInt;     /* pMarginal->funcInt                                                     */
} PbrtGpuEnvMap;
typedef struct PbrtGpuLightShape {
    uint32_t type;        /* PBRTGPU_PRIM_SPHERE | PBRTGPU_PRIM_DISK            */
    uint32_t index;       /* into quadrics[]                                    */
    float    area;        /* ShapeSet::areas[i]                                 */
} PbrtGpuLightShape;

/* ---- the flattened Scene --------------------------------------------------- */
typedef struct PbrtGpuSceneDesc {
    uint32_t nNodes;        const PbrtGpuBVHNode    *nodes;       /* scene BVH first (root = 0), nested BVHs after */
    uint32_t nPrimitives;   const PbrtGpuPrimitive  *primitives;  /* scene BVH's leaf order first, nested after    */
    uint32_t nTriangles;    const PbrtGpuTriangle   *triangles;
    uint32_t nMeshes;       const PbrtGpuMesh       *meshes;
    uint32_t nVertices;     const float *P;   /* 3*nVertices                    */
                            const float *N;   /* 3*nVertices or NULL            */
                            const float *uv;  /* 2*nVertices or NULL            */
    uint32_t nQuadrics;     const PbrtGpuQuadric    *quadrics;
    uint32_t nXforms;       const PbrtGpuXform      *xforms;
    uint32_t nMaterials;    const PbrtGpuMaterial   *materials;
    uint32_t nMeasured;     const PbrtGpuMeasured   *measured;
    uint32_t nKdNodes;      const PbrtGpuKdNode     *kdNodes;
                            const PbrtGpuBrdfSample *kdData;
    uint32_t nLights;       const PbrtGpuLight      *lights;
    uint32_t nLightShapes;  const PbrtGpuLightShape *lightShapes;
    uint32_t nAnimated;     const PbrtGpuAnimated   *animated;    /* may be 0 / NULL                               */
    /* (appended in round 2: earlier offsets are unchanged) one entry per InfiniteAreaLight */
    uint32_t nEnvMaps;      const PbrtGpuEnvMap     *envMaps;     /* may be 0 / NULL                               */
    uint32_t nEnvData;      const float             *envData;
} PbrtGpuSceneDesc;

/* ---- one Render() call ----------------------------------------------------- */
enum { PBRTGPU_INTEGRATOR_PATH = 0,            /* PathIntegrator, integrators/path.cpp            */
       PBRTGPU_INTEGRATOR_DIRECTLIGHTING = 1   /* DirectLightingIntegrator, integrators/directlighting.cpp */ };
enum { PBRTGPU_STRATEGY_ALL = 0, PBRTGPU_STRATEGY_ONE = 1 };   /* LightStrategy, directlighting.h:44 */
#define PBRTGPU_MAX_SAMPLE_ARRAYS 32

typedef struct PbrtGpuRenderDesc {
    /* PerspectiveCamera (cameras/perspective.cpp:41-49, core/camera.cpp:83-102) */
    float rasterToCamera[16];   /* RasterToCamera.m                             */
    float cameraToWorld[16];    /* CameraToWorld start transform .m (static)    */
    float lensRadius, focalDistance;
    float shutterOpen, shutterClose;
    /* ImageFilm (film/image.cpp:40-51) */
    int32_t xResolution, yResolution;
    float   cropWindow[4];
    float   filterXWidth, filterYWidth;   /* Filter::xWidth / yWidth; the weights: filterTable below */
    /* LDSampler (samplers/lowdiscrepancy.cpp:40-51): power of two              */
    int32_t samplesPerPixel;
    /* PathIntegrator (integrators/path.cpp:125-128) / DirectLightingIntegrator
     * (directlighting.cpp:120-133)                                              */
    int32_t maxDepth;
    int32_t integrator;         /* PBRTGPU_INTEGRATOR_*                          */
    int32_t lightStrategy;      /* DIRECTLIGHTING: PBRTGPU_STRATEGY_*            */
    /* Sample::n1D / n2D (core/sampler.h:93-116) as the integrators' RequestSamples
     * calls built them (surface integrator first, then EmissionIntegrator's two
     * 1-D requests, samplerrenderer.cpp Sample ctor): how many values every
     * pixel sample carries in each requested array.  They fix the LD tables'
     * draw order (montecarlo.cpp:200-252).  PATH: 14 x {1}, 9 x {1}.             */
    int32_t n1D, n2D;
    uint16_t size1D[PBRTGPU_MAX_SAMPLE_ARRAYS];
    uint16_t size2D[PBRTGPU_MAX_SAMPLE_ARRAYS];
    /* SamplerRenderer::Render task decomposition (samplerrenderer.cpp:203-205):
     * nTasks is a reference INPUT (it fixes which RNG stream feeds which
     * pixel); tasks [taskBegin, taskEnd) are rendered by this call, which is
     * how an image is sharded across GPUs.                                      */
    int32_t nTasks;
    int32_t taskBegin, taskEnd;
    /* Optional interleaved sharding inside [taskBegin, taskEnd): only tasks t with
     * (t / shardBlock) % shardCount == shardIndex are rendered.  shardCount <= 1
     * (or 0) renders them all.  Interleaving blocks of rows balances cheap
     * (background) and expensive (object) pixels across GPUs; tasks share nothing,
     * so the sum of the shards' films is the one-GPU film.                       */
    int32_t shardCount, shardIndex, shardBlock;
    /* The pixel filter as ImageFilm holds it (film/image.cpp:57-69): Filter::Evaluate at the 16 x 16 table positions
     * ((i + .5) * width / 16), row = y.  hasFilterTable == 0 reads as the box filter (every weight 1).  Filters wider than
     * half a pixel (gaussian, mitchell, sinc, triangle, wide boxes) make every sample splat into neighbouring tasks' pixels:
     * host front end + oracle (pinned on the reference binary at --ncores 1, where the reference's own sums are ordered);
     * the device entry points answer PBRTGPU_EUNSUPPORTED for them unless the library was built with -DPBRT_FILTER=1
     * (atomic adds: sums complete, order undefined as in the reference above one thread; host-verified, not yet run on a
     * GPU).  (Appended in round 2: earlier offsets are unchanged.) */
    int32_t hasFilterTable;
    float   filterTable[256];
} PbrtGpuRenderDesc;

/* ImageFilm::Pixel {Lxyz[3], weightSum} (film/image.h:69-78), row-major over
 * the film's pixel extent (xPixelCount x yPixelCount, image.cpp:48-51).        */
typedef struct PbrtGpuPixel {
    float Lxyz[3];
    float weightSum;
} PbrtGpuPixel;

typedef struct PbrtGpuStats {
    double   renderMs;        /* CUDA-event time of the device work             */
    double   tableMs;         /* ... of which LD table generation               */
    double   traceMs;         /* ... of which path kernel phase 1 (table-driven bounces 0..2) */
    double   tailMs;          /* ... of which path kernel phase 2 (RNG-driven tails + film)  */
    uint64_t cameraSamples;
    uint64_t closestRays, shadowRays;
    uint64_t nodesVisited;    /* BVH nodes whose bounds were tested             */
    uint64_t primTests;
    uint64_t kernelLaunches;
    /* filled only under pbrtgpu_set_stats(scene, 2): the ray-traversal kernel (k_trace) timed launch by launch
     * with CUDA events on its stream, for the roofline figure (the render then runs on ONE lane so that
     * nothing else overlaps it) */
    double   rayKernelMs;
    uint64_t rayKernelLaunches;
} PbrtGpuStats;

typedef struct PbrtGpuRay { float o[3], d[3], mint, maxt, time; } PbrtGpuRay;   /* core/geometry.h:338-362 */
typedef struct PbrtGpuHit {
    int32_t  hit;             /* 0/1                                            */
    uint32_t primitive;       /* index into primitives[] (the nested one under a TransformedPrimitive) */
    float    t;               /* ray.maxt after BVHAccel::Intersect             */
    float    b1, b2;          /* triangle barycentrics (0 for quadrics)         */
    float    p[3], n[3];      /* dg.p, dg.nn (world space)                      */
} PbrtGpuHit;

typedef struct PbrtGpuScene PbrtGpuScene;

int  pbrtgpu_device_count(void);
int  pbrtgpu_set_device(int ordinal);
const char *pbrtgpu_last_error(void);

/* The checks pbrtgpu_scene_create / pbrtgpu_scene_update run before anything is copied, on their own (no device needed):
 * PBRTGPU_OK, or the PBRTGPU_EINVAL / PBRTGPU_EUNSUPPORTED the create call would answer.  A descriptor that passes keeps
 * every index the kernels follow inside its array, every walk finite and inside its stack. */
int  pbrtgpu_scene_validate(const PbrtGpuSceneDesc *desc);
/* replaces: Scene / BVHAccel / GeometricPrimitive object graph (core/scene.h) */
int  pbrtgpu_scene_create(const PbrtGpuSceneDesc *desc, PbrtGpuScene **out);
void pbrtgpu_scene_destroy(PbrtGpuScene *scene);
/* replaces: rebuilding the Scene when the world changes between frames (pbrtWorldEnd per frame,
 * core/api.cpp:1150-1209).  Copies a new descriptor into the device buffers of an existing handle
 * (reallocating only what grew); must not run concurrently with a render call on the same handle. */
int  pbrtgpu_scene_update(PbrtGpuScene *scene, const PbrtGpuSceneDesc *desc);
/* replaces: the PBRT_PROBES_COUNTERS build switch (core/probes.h:219-395).  on == 1
 * makes the following render calls run the counting instantiation of the path
 * kernels and fill PbrtGpuStats.{closestRays,shadowRays,nodesVisited,primTests};
 * on == 2 times every launch of the traversal kernel instead (rayKernelMs);
 * timed runs leave it off (0).                                                 */
int  pbrtgpu_set_stats(PbrtGpuScene *scene, int on);

/* The checks every render call runs on its descriptor first, on their own (no device needed), and the film's pixel
 * extent -- ImageFilm::xPixelStart, yPixelStart, xPixelCount, yPixelCount (film/image.cpp:40-51) -- that a caller sizes
 * `pixels` with.                                                                */
int  pbrtgpu_render_validate(const PbrtGpuRenderDesc *rd);
/* ... and with the scene it is for: the Sample layout against the scene's lights (DirectLightingIntegrator::RequestSamples,
 * directlighting.cpp:53-78), recursion depth against specular materials -- everything a render call refuses. */
int  pbrtgpu_render_validate_scene(const PbrtGpuSceneDesc *desc, const PbrtGpuRenderDesc *rd);
int  pbrtgpu_film_extent(const PbrtGpuRenderDesc *rd, int32_t extent[4]);
/* replaces: SamplerRenderer::Render (samplerrenderer.cpp:188-222) minus the
 * final WriteImage().  `pixels` is HOST memory (xPixelCount*yPixelCount);
 * pixels of tasks outside [taskBegin,taskEnd) are left zero.                   */
int  pbrtgpu_render(PbrtGpuScene *scene, const PbrtGpuRenderDesc *rd,
                    PbrtGpuPixel *pixels, PbrtGpuStats *stats);
/* same, but `d_pixels` is DEVICE memory on the current device (e.g. a torch
 * tensor about to be NCCL-gathered) and `stream` a cudaStream_t (NULL = default);
 * asynchronous unless stats != NULL.                                            */
int  pbrtgpu_render_device(PbrtGpuScene *scene, const PbrtGpuRenderDesc *rd,
                           void *d_pixels, void *stream, PbrtGpuStats *stats);

/* ---- one Render() over several GPUs (one process per GPU) ---------------------------------------------------------
 * replaces: nothing in the reference (pbrt-v2 is single-process); it is what SURVEY 8(e) asks of the drop-in renderer.
 * A communicator is an NCCL communicator owned by the library (libnccl.so.2 is loaded at run time: the build does not
 * link it).  Rank 0 obtains a unique id, the host distributes its PBRTGPU_COMM_ID_BYTES to every rank by whatever means
 * it has (the CLI uses a rendezvous file, bench.py torch.distributed), every rank then creates the communicator on ITS
 * current device.  pbrtgpu_render_sharded renders this rank's share of [taskBegin, taskEnd) -- blocks of `shardBlock`
 * consecutive tasks dealt round-robin to the ranks (rd->shardCount / shardIndex are overwritten from the communicator;
 * shardBlock 0 = four rows of one-pixel tasks) -- and sums the per-rank films onto rank 0 with ONE ncclReduce (float
 * sum over 4 floats per pixel: pixels owned by another rank contribute zeros, edge splats that cross an ownership
 * border add up, which a gather of owned tiles would lose).  `pixels` (HOST memory) receives the whole film on rank 0;
 * other ranks may pass NULL (so may rank 0: the summed film then stays on the device, which is what a benchmark of the
 * device-resident rate wants).                                                                                        */
#define PBRTGPU_COMM_ID_BYTES 128
typedef struct PbrtGpuComm PbrtGpuComm;
int  pbrtgpu_comm_unique_id(void *id);
int  pbrtgpu_comm_create(const void *id, int rank, int worldSize, PbrtGpuComm **out);
void pbrtgpu_comm_destroy(PbrtGpuComm *comm);
int  pbrtgpu_render_sharded(PbrtGpuScene *scene, const PbrtGpuRenderDesc *rd, PbrtGpuComm *comm,
                            PbrtGpuPixel *pixels, PbrtGpuStats *stats);

/* replaces: BVHAccel::Intersect (bvh.cpp:403-454) over a ray batch             */
int  pbrtgpu_intersect(PbrtGpuScene *scene, const PbrtGpuRay *rays, int n,
                       PbrtGpuHit *hits);
/* replaces: BVHAccel::IntersectP (bvh.cpp:457-503) over a ray batch            */
int  pbrtgpu_intersect_p(PbrtGpuScene *scene, const PbrtGpuRay *rays, int n,
                         uint8_t *occluded);
/* replaces: LDSampler::GetMoreSamples / LDPixelSample (montecarlo.cpp:200-252)
 * for the FIRST pixel of task `taskNum` (test hook).  out = spp rows of
 * 5 + sum(size1D) + 2*sum(size2D) floats: imageX,imageY,lensU,lensV,time, the 1-D
 * arrays' values, the 2-D arrays' pairs (Sample::oneD / twoD order) -- 37 floats
 * for the path integrator's layout.                                            */
int  pbrtgpu_ld_samples(const PbrtGpuRenderDesc *rd, int taskNum, float *out);

/* replaces: BVHAccel::BVHAccel's build -- recursiveBuild + flattenBVHTree (accelerators/bvh.cpp:153-202, 210-380,
 * 383-399) -- for the FullyRefine'd primitives whose world bounds are `bounds` (nPrimitives x 6 floats: pMin xyz,
 * pMax xyz, in the order BVHAccel received the primitives).  Builds on the current device and returns to HOST memory
 *   nodes[0 .. *nNodes)      the LinearBVHNode array in flattenBVHTree's depth-first order (capacity 2*nPrimitives - 1);
 *   order[0 .. nPrimitives)  BVHAccel::primitives after the build: primitives[i] = input primitive order[i].
 * The tree is the one the reference builds from the same input (same split decisions, same std::partition /
 * std::nth_element permutations, same flattening); node bounds may differ in the sign of a zero coordinate only.
 * splitMethod: PBRTGPU_SPLIT_SAH ("sah", the default of CreateBVHAccelerator, bvh.cpp:506-511); "middle" and "equal"
 * answer PBRTGPU_EUNSUPPORTED.  maxPrimsInNode 1..255.  Bounds that are not finite answer PBRTGPU_EINVAL (the
 * reference's SAH bucket index is undefined for them).                                                             */
enum { PBRTGPU_SPLIT_MIDDLE = 0, PBRTGPU_SPLIT_EQUAL_COUNTS = 1, PBRTGPU_SPLIT_SAH = 2 };   /* bvh.h:56 */
typedef struct PbrtGpuBuildStats {
    double   buildMs;         /* CUDA-event time of the build, copies included  */
    uint32_t nLevels;         /* depth of the tree + 1                          */
    uint32_t reserved;
} PbrtGpuBuildStats;
int  pbrtgpu_bvh_build(const float *bounds, uint32_t nPrimitives, uint32_t maxPrimsInNode, int splitMethod,
                       PbrtGpuBVHNode *nodes, uint32_t *nNodes, uint32_t *order, PbrtGpuBuildStats *stats);

/* replaces: LoopSubdiv::Refine (shapes/loopsubdiv.cpp:221-422) -- the subdivision levels, the limit-surface positions
 * and normals, the triangle list.  Input: what the LoopSubdiv constructor builds (loopsubdiv.cpp:156-216), the SDVertex /
 * SDFace arrays as indices; that topology pass over the control mesh stays on the host.  Output, to HOST memory: the
 * arguments of the TriangleMesh Refine creates (loopsubdiv.cpp:404-422) -- P (limit positions), N (limit normals, not
 * normalised, as the reference leaves them), vertex indices -- identical to the reference's, element for element.
 * Sizes: pbrtgpu_loopsubdiv_counts.  The control mesh must be a manifold (an edge in at most two faces, neighbour links
 * reciprocal), which the reference's algorithm assumes as well: anything else answers PBRTGPU_EINVAL.               */
typedef struct PbrtGpuSubdivMesh {
    uint32_t nVertices, nFaces;
    const float   *P;          /* nVertices x 3: SDVertex::P                                   */
    const int32_t *startFace;  /* nVertices: SDVertex::startFace (-1: no face uses the vertex) */
    const uint8_t *vflags;     /* nVertices: bit 0 SDVertex::boundary, bit 1 SDVertex::regular */
    const int32_t *faceV;      /* nFaces x 3: SDFace::v                                        */
    const int32_t *faceF;      /* nFaces x 3: SDFace::f (-1: no neighbour across edge k)       */
} PbrtGpuSubdivMesh;
int  pbrtgpu_loopsubdiv_counts(const PbrtGpuSubdivMesh *mesh, int nLevels, uint32_t *nVerticesOut, uint32_t *nTrianglesOut);
int  pbrtgpu_loopsubdiv_refine(const PbrtGpuSubdivMesh *mesh, int nLevels, float *P, float *N, int32_t *vertexIndices);

#ifdef __cplusplus
}
#endif
#endif /* PBRTGPU_H */
