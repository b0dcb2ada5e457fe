// bvh.cpp -- GpuBVHAccel: host-side construction of the bounding volume
// hierarchy the CUDA traversal kernels walk.
//
// The GPU reproduces BVHAccel::Intersect's traversal ORDER (near child first by
// split axis, accelerators/bvh.cpp:403-454), so hit/miss ties and `t` ties
// resolve as in the reference only if the tree itself is the same tree.  The
// build therefore makes the same decisions as BVHAccel::recursiveBuild
// (accelerators/bvh.cpp:210-380): same refinement order, same centroid/bucket
// float arithmetic, same std::partition / std::nth_element calls from the same
// libstdc++, same depth-first flattening (bvh.cpp:383-399).  The code is an
// array-based rewrite (no BVHBuildNode tree: nodes are emitted directly in
// depth-first order and the second-child offset is patched afterwards).
#include "scene.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cmath>
#include <memory>
#include <thread>

namespace pb {

namespace {

struct PrimInfo {   // BVHPrimitiveInfo, bvh.cpp:40-49
    int primitiveNumber;
    V3 centroid;
    BBox bounds;
};

struct Builder {
    std::vector<PrimInfo> &info;
    std::vector<PbrtGpuBVHNode> &nodes;   // this builder's subtree in depth-first order; child offsets count from nodes[0]
    uint32_t maxPrimsInNode;
    int splitMethod;

    static void SetBounds(PbrtGpuBVHNode &n, const BBox &b) {
        n.bmin[0] = b.pMin.x; n.bmin[1] = b.pMin.y; n.bmin[2] = b.pMin.z;
        n.bmax[0] = b.pMax.x; n.bmax[1] = b.pMax.y; n.bmax[2] = b.pMax.z;
    }
    static BBox GetBounds(const PbrtGpuBVHNode &n) {
        BBox b;
        b.pMin = V3(n.bmin[0], n.bmin[1], n.bmin[2]);
        b.pMax = V3(n.bmax[0], n.bmax[1], n.bmax[2]);
        return b;
    }
    // A leaf appends its range to orderedPrims (bvh.cpp:232-236).  Leaves are reached left to right and their ranges tile
    // [0, n), so orderedPrims.size() at that moment IS `start`, and orderedPrims as a whole is the build array in its final
    // order: the driver copies it once at the end.  (That is what lets subtrees be built independently of each other.)
    uint32_t MakeLeaf(uint32_t nodeIdx, uint32_t start, uint32_t end, const BBox &bbox) {
        uint32_t first = start;
        PbrtGpuBVHNode &n = nodes[nodeIdx];
        SetBounds(n, bbox);
        n.offset = first;
        n.nPrimitives = (uint8_t)(end - start);
        n.axis = 0;
        return nodeIdx;
    }
    // interior node: children are built left then right so that the left child
    // is nodeIdx+1 (what flattenBVHTree produces); the node's bounds are the
    // union of the CHILDREN's bounds (BVHBuildNode::InitInterior, bvh.cpp:60-66)
    uint32_t MakeInterior(uint32_t nodeIdx, int dim, uint32_t start, uint32_t mid, uint32_t end) {
        uint32_t c0 = Build(start, mid);
        uint32_t c1 = Build(mid, end);
        PbrtGpuBVHNode &n = nodes[nodeIdx];
        SetBounds(n, Union(GetBounds(nodes[c0]), GetBounds(nodes[c1])));
        n.offset = c1;
        n.nPrimitives = 0;
        n.axis = (uint8_t)dim;
        (void)c0;
        return nodeIdx;
    }

    uint32_t Build(uint32_t start, uint32_t end) {   // recursiveBuild, bvh.cpp:210-380
        uint32_t nodeIdx = (uint32_t)nodes.size();
        nodes.push_back(PbrtGpuBVHNode());
        memset(&nodes[nodeIdx], 0, sizeof(PbrtGpuBVHNode));
        BBox bbox;
        int dim;
        uint32_t mid;
        if (!Decide(start, end, &bbox, &dim, &mid)) return MakeLeaf(nodeIdx, start, end, bbox);
        return MakeInterior(nodeIdx, dim, start, mid, end);
    }
    // What recursiveBuild decides for [start, end) before it recurses: leaf (false; *bbox set) or interior (true; the range
    // partitioned around *mid along *dim)
    bool Decide(uint32_t start, uint32_t end, BBox *bboxOut, int *dimOut, uint32_t *midOut) {
        BBox bbox;
        for (uint32_t i = start; i < end; ++i) bbox = Union(bbox, info[i].bounds);
        *bboxOut = bbox;
        uint32_t nPrimitives = end - start;
        if (nPrimitives == 1) return false;

        BBox centroidBounds;
        for (uint32_t i = start; i < end; ++i) centroidBounds = Union(centroidBounds, info[i].centroid);
        int dim = centroidBounds.MaximumExtent();
        uint32_t mid = (start + end) / 2;
        *dimOut = dim;
        if (centroidBounds.pMax[dim] == centroidBounds.pMin[dim]) {   // bvh.cpp:239-261
            if (nPrimitives <= maxPrimsInNode) return false;
            *midOut = mid;
            return true;
        }
        PrimInfo *base = info.data();
        auto equalCounts = [&]() {   // bvh.cpp:280-286
            mid = (start + end) / 2;
            std::nth_element(base + start, base + mid, base + end,
                             [dim](const PrimInfo &a, const PrimInfo &b) { return a.centroid[dim] < b.centroid[dim]; });
        };
        if (splitMethod == 0) {   // SPLIT_MIDDLE, bvh.cpp:265-279 (falls through to equal counts)
            float pmid = .5f * (centroidBounds.pMin[dim] + centroidBounds.pMax[dim]);
            PrimInfo *midPtr = std::partition(base + start, base + end,
                                              [dim, pmid](const PrimInfo &a) { return a.centroid[dim] < pmid; });
            mid = (uint32_t)(midPtr - base);
            if (mid == start || mid == end) equalCounts();
        } else if (splitMethod == 1) {
            equalCounts();
        } else if (nPrimitives <= 4) {   // SPLIT_SAH small case, bvh.cpp:289-294
            equalCounts();
        } else {   // SPLIT_SAH, bvh.cpp:295-361
            const int nBuckets = 12;
            int count[nBuckets];
            BBox bounds[nBuckets];
            for (int i = 0; i < nBuckets; ++i) count[i] = 0;
            const float cmin = centroidBounds.pMin[dim], cmax = centroidBounds.pMax[dim];
            for (uint32_t i = start; i < end; ++i) {
                int b = nBuckets * ((info[i].centroid[dim] - cmin) / (cmax - cmin));
                if (b == nBuckets) b = nBuckets - 1;
                count[b]++;
                bounds[b] = Union(bounds[b], info[i].bounds);
            }
            float cost[nBuckets - 1];
            for (int i = 0; i < nBuckets - 1; ++i) {
                BBox b0, b1;
                int count0 = 0, count1 = 0;
                for (int j = 0; j <= i; ++j) { b0 = Union(b0, bounds[j]); count0 += count[j]; }
                for (int j = i + 1; j < nBuckets; ++j) { b1 = Union(b1, bounds[j]); count1 += count[j]; }
                // NB: an empty side gives 0 * inf = NaN exactly as in the reference
                cost[i] = .125f + (count0 * b0.SurfaceArea() + count1 * b1.SurfaceArea()) / bbox.SurfaceArea();
            }
            float minCost = cost[0];
            uint32_t minCostSplit = 0;
            for (int i = 1; i < nBuckets - 1; ++i)
                if (cost[i] < minCost) { minCost = cost[i]; minCostSplit = i; }
            if (nPrimitives > maxPrimsInNode || minCost < nPrimitives) {
                PrimInfo *pmid = std::partition(base + start, base + end, [=](const PrimInfo &p) {
                    int b = nBuckets * ((p.centroid[dim] - cmin) / (cmax - cmin));
                    if (b == nBuckets) b = nBuckets - 1;
                    return b <= (int)minCostSplit;
                });
                mid = (uint32_t)(pmid - base);
            } else {
                return false;
            }
        }
        *midOut = mid;
        return true;
    }
};

// The same tree with its subtrees built side by side.  recursiveBuild's work on [start, end) touches that range of the
// build array only and -- see MakeLeaf -- nothing it emits depends on what was built before it except the position of its
// nodes in the depth-first array.  So: the top `depth` levels fork (each call decides its node, then its two halves run
// on two threads), every range below is built by an ordinary Builder into its own node array, and the pieces are
// stitched together in preorder with their child offsets rebased.
struct ParallelBuild {
    struct Tree {
        std::vector<PbrtGpuBVHNode> piece;   // a serially built subtree (or a single leaf); empty for a forked interior node
        int dim = 0;
        std::unique_ptr<Tree> child[2];
    };
    std::vector<PrimInfo> &info;
    uint32_t maxPrimsInNode;
    int splitMethod;
    uint32_t serialBelow;   // ranges of at most this many primitives are not forked

    std::unique_ptr<Tree> Build(uint32_t start, uint32_t end, int depth) {
        std::unique_ptr<Tree> t(new Tree);
        Builder b{info, t->piece, maxPrimsInNode, splitMethod};
        if (depth <= 0 || end - start <= serialBelow) {
            t->piece.reserve(2 * (size_t)(end - start));
            b.Build(start, end);
            return t;
        }
        BBox bbox;
        uint32_t mid = 0;
        if (!b.Decide(start, end, &bbox, &t->dim, &mid)) {
            t->piece.push_back(PbrtGpuBVHNode());
            memset(&t->piece[0], 0, sizeof(PbrtGpuBVHNode));
            b.MakeLeaf(0, start, end, bbox);
            return t;
        }
        std::thread left([&]() { t->child[0] = Build(start, mid, depth - 1); });
        t->child[1] = Build(mid, end, depth - 1);
        left.join();
        return t;
    }
    // preorder positions first (a piece's size is known, a forked node takes one slot), then every piece is copied to its
    // place -- second-child offsets rebased -- on its own thread, and the forked nodes are filled in bottom-up
    struct Placed { const Tree *t; uint32_t base; };
    static uint32_t Place(const Tree &t, uint32_t base, std::vector<Placed> *pieces, std::vector<Placed> *forks) {
        if (!t.child[0]) {
            pieces->push_back(Placed{&t, base});
            return base + (uint32_t)t.piece.size();
        }
        forks->push_back(Placed{&t, base});
        const uint32_t c1 = Place(*t.child[0], base + 1, pieces, forks);
        return Place(*t.child[1], c1, pieces, forks);
    }
    static uint32_t SizeOf(const Tree &t) {
        return t.child[0] ? 1 + SizeOf(*t.child[0]) + SizeOf(*t.child[1]) : (uint32_t)t.piece.size();
    }
    static void Stitch(const Tree &t, std::vector<PbrtGpuBVHNode> *out) {
        std::vector<Placed> pieces, forks;
        const uint32_t total = Place(t, 0, &pieces, &forks);
        out->resize(total);
        PbrtGpuBVHNode *o = out->data();
        auto copyPiece = [o](const Placed &p) {
            PbrtGpuBVHNode *dst = o + p.base;
            memcpy(dst, p.t->piece.data(), p.t->piece.size() * sizeof(PbrtGpuBVHNode));
            if (p.base)
                for (size_t i = 0; i < p.t->piece.size(); ++i)
                    if (dst[i].nPrimitives == 0) dst[i].offset += p.base;   // interior: second child's index
        };
        const unsigned nThreads = std::min<unsigned>((unsigned)pieces.size(), std::max(1u, std::thread::hardware_concurrency()));
        std::vector<std::thread> workers;
        for (unsigned w = 1; w < nThreads; ++w)
            workers.emplace_back([&, w]() { for (size_t k = w; k < pieces.size(); k += nThreads) copyPiece(pieces[k]); });
        for (size_t k = 0; k < pieces.size(); k += std::max(1u, nThreads)) copyPiece(pieces[k]);
        for (std::thread &w : workers) w.join();
        for (size_t k = forks.size(); k-- > 0;) {   // preorder list walked backwards: children before parents
            const Placed &f = forks[k];
            const uint32_t c1 = f.base + 1 + SizeOf(*f.t->child[0]);
            PbrtGpuBVHNode &n = o[f.base];   // BVHBuildNode::InitInterior, bvh.cpp:60-66
            memset(&n, 0, sizeof(n));
            Builder::SetBounds(n, Union(Builder::GetBounds(o[f.base + 1]), Builder::GetBounds(o[c1])));
            n.offset = c1;
            n.nPrimitives = 0;
            n.axis = (uint8_t)f.t->dim;
        }
    }
};

// Primitive::FullyRefine (core/primitive.cpp:57-69) uses a LIFO todo list, so a
// mesh's triangles reach the accelerator in REVERSE index order; that order is
// the input order of the build and must be kept.
void FullyRefine(const GeometricPrimitive &gp, std::vector<RefinedPrim> *out,
                 std::vector<std::shared_ptr<Shape>> *owned, std::string *err) {
    if (gp.animated) {   // TransformedPrimitive::CanIntersect() is true: handed over as it is
        RefinedPrim r;
        r.type = PBRTGPU_PRIM_ANIMATED;
        r.animated = gp.animated.get();
        out->push_back(r);
        return;
    }
    const Shape *s = gp.shape.get();
    auto emitMesh = [&](const TriangleMesh *m) {
        for (int i = m->ntris - 1; i >= 0; --i) {
            RefinedPrim r;
            r.type = PBRTGPU_PRIM_TRIANGLE;
            r.shape = m;
            r.tri = i;
            r.material = gp.material;
            r.areaLight = gp.areaLight;
            out->push_back(r);
        }
    };
    switch (s->kind) {
    case Shape::TRIANGLE_MESH: emitMesh(static_cast<const TriangleMesh *>(s)); break;
    case Shape::LOOP_SUBDIV: {
        std::shared_ptr<Shape> mesh(static_cast<const LoopSubdiv *>(s)->Refine());
        if (!mesh) { *err = static_cast<const LoopSubdiv *>(s)->refineError; break; }
        owned->push_back(mesh);
        emitMesh(static_cast<const TriangleMesh *>(mesh.get()));
        break;
    }
    case Shape::SPHERE:
    case Shape::DISK: {
        RefinedPrim r;
        r.type = s->kind == Shape::SPHERE ? PBRTGPU_PRIM_SPHERE : PBRTGPU_PRIM_DISK;
        r.shape = s;
        r.tri = 0;
        r.material = gp.material;
        r.areaLight = gp.areaLight;
        out->push_back(r);
        break;
    }
    }
}

}  // namespace

bool g_keepBuildInput = false;
static double NowSeconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
PrepTimer::PrepTimer(const char *w) : what(w), t0(getenv("PBRTGPU_PREP_TIMING") ? NowSeconds() : -1.) {}
void PrepTimer::Stop() {
    if (t0 >= 0.) fprintf(stderr, "scene preparation: %-30s %.3f s\n", what, NowSeconds() - t0);
    t0 = -1.;
}
uint32_t g_bvhParallelMin = 1u << 17;   // builds of at least this many primitives run their subtrees on several threads

GpuBVHAccel::GpuBVHAccel(const std::vector<GeometricPrimitive> &p, uint32_t mp, const std::string &sm, const std::string &builder) {
    // BVHAccel::BVHAccel, bvh.cpp:153-202
    maxPrimsInNode = std::min(255u, mp);
    std::vector<RefinedPrim> refined;
    {
        PrepTimer t("FullyRefine");
        for (const GeometricPrimitive &gp : p) FullyRefine(gp, &refined, &owned, &buildError);
    }
    if (!buildError.empty()) return;
    if (sm == "sah") splitMethod = 2;
    else if (sm == "middle") splitMethod = 0;
    else if (sm == "equal") splitMethod = 1;
    else {
        fprintf(stderr, "Warning: BVH split method \"%s\" unknown.  Using \"sah\".\n", sm.c_str());
        splitMethod = 2;
    }
    if (refined.empty()) return;
    std::vector<float> bounds(refined.size() * 6);
    PrepTimer tb("WorldBound of every primitive");
    for (size_t i = 0; i < refined.size(); ++i) {
        const BBox b = refined[i].WorldBound();
        float *o = &bounds[6 * i];
        o[0] = b.pMin.x; o[1] = b.pMin.y; o[2] = b.pMin.z; o[3] = b.pMax.x; o[4] = b.pMax.y; o[5] = b.pMax.z;
    }
    tb.Stop();
    // (the reference's SAH bucket index is int(NaN) for a primitive without finite bounds, bvh.cpp:300-306: an
    // out-of-bounds write there, an error here -- as in pbrtgpu_bvh_build)
    for (float v : bounds)
        if (!std::isfinite(v)) { buildError = "a primitive's bounds are not finite (a singular or overflowing transform?)"; return; }
    std::vector<uint32_t> order;
    PrepTimer tbuild("BVH build");
    if (builder == "gpu") {   // the same tree, built on the device (include/pbrtgpu.h: pbrtgpu_bvh_build)
        nodes.resize(2 * refined.size());
        order.resize(refined.size());
        uint32_t nNodes = 0;
        const int rc = pbrtgpu_bvh_build(bounds.data(), (uint32_t)refined.size(), maxPrimsInNode, splitMethod, nodes.data(), &nNodes,
                                         order.data(), nullptr);
        if (rc != PBRTGPU_OK) {
            buildError = std::string("BVH build on the device failed: ") + pbrtgpu_last_error();
            nodes.clear();
            return;
        }
        nodes.resize(nNodes);
    } else {
        BuildBVHOnHost(bounds.data(), (uint32_t)refined.size(), maxPrimsInNode, splitMethod, &nodes, &order);
    }
    tbuild.Stop();
    PrepTimer tr("reordered primitive list");
    primitives.reserve(refined.size());
    for (uint32_t i : order) primitives.push_back(refined[i]);
    if (g_keepBuildInput) buildInput.swap(bounds);
}

void BuildBVHOnHost(const float *bounds, uint32_t n, uint32_t maxPrimsInNode, int splitMethod, std::vector<PbrtGpuBVHNode> *nodes,
                    std::vector<uint32_t> *order) {
    nodes->clear();
    order->clear();
    if (n == 0) return;
    std::vector<PrimInfo> info;
    info.reserve(n);
    for (uint32_t i = 0; i < n; ++i) {
        PrimInfo pi;
        pi.primitiveNumber = (int)i;
        const float *b = bounds + 6 * (size_t)i;
        pi.bounds.pMin = V3(b[0], b[1], b[2]);
        pi.bounds.pMax = V3(b[3], b[4], b[5]);
        pi.centroid = .5f * pi.bounds.pMin + .5f * pi.bounds.pMax;   // bvh.cpp:44-46
        info.push_back(pi);
    }
    nodes->reserve(2 * (size_t)n);
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    if (n >= g_bvhParallelMin && hw > 1) {
        int depth = 1;
        while ((1u << depth) < 2 * hw && depth < 7) ++depth;   // ~2 pieces per hardware thread, at most 128
        ParallelBuild pb{info, maxPrimsInNode, splitMethod, std::max(1u, g_bvhParallelMin / 64)};
        auto T0 = std::chrono::steady_clock::now();
        std::unique_ptr<ParallelBuild::Tree> t = pb.Build(0, n, depth);
        auto T1 = std::chrono::steady_clock::now();
        ParallelBuild::Stitch(*t, nodes);
        auto T2 = std::chrono::steady_clock::now();
        if (getenv("PBRTGPU_BVH_TIMING")) fprintf(stderr, "host bvh: build %.3f s, stitch %.3f s, depth %d\n", std::chrono::duration<double>(T1 - T0).count(), std::chrono::duration<double>(T2 - T1).count(), depth);
    } else {
        Builder b{info, *nodes, maxPrimsInNode, splitMethod};
        b.Build(0, n);
    }
    order->resize(n);   // orderedPrims: the build array in its final order (see Builder::MakeLeaf)
    for (uint32_t i = 0; i < n; ++i) (*order)[i] = (uint32_t)info[i].primitiveNumber;
}

TransformedPrimitive::TransformedPrimitive(const GeometricPrimitive &base, const AnimatedTransform &w2p)
    : shape(base.shape), WorldToPrimitive(w2p) {
    // api.cpp:1033-1043: "refine animated shape and create BVH if more than one shape created"
    std::vector<GeometricPrimitive> one(1, base);
    primitive.reset(new GpuBVHAccel(one));
    if (primitive->primitives.size() == 1) primitive->nodes.clear();   // baseprim = refinedPrimitives[0]: no BVH around it
}

BBox TransformedPrimitive::WorldBound() const {   // primitive.h:116-118
    BBox b;
    if (!primitive->nodes.empty()) {   // BVHAccel::WorldBound, bvh.cpp:205-207
        const PbrtGpuBVHNode &n = primitive->nodes[0];
        b.pMin = V3(n.bmin[0], n.bmin[1], n.bmin[2]);
        b.pMax = V3(n.bmax[0], n.bmax[1], n.bmax[2]);
    } else if (!primitive->primitives.empty()) {
        b = primitive->primitives[0].WorldBound();
    }
    return WorldToPrimitive.MotionBounds(b, true);
}

GpuBVHAccel *CreateGpuBVHAccelerator(const std::vector<GeometricPrimitive> &prims, const ParamSet &ps) {
    // CreateBVHAccelerator, bvh.cpp:506-511
    std::string splitMethod = ps.FindOneString("splitmethod", "sah");
    uint32_t maxPrimsInNode = ps.FindOneInt("maxnodeprims", 4);
    // "string builder" [ "host" | "gpu" ] (also PBRTGPU_BVH_BUILDER in the environment): where the tree is built.  The
    // result is the same tree; "host" stays the default until the device build has been timed against it.
    const char *env = getenv("PBRTGPU_BVH_BUILDER");
    std::string builder = ps.FindOneString("builder", env ? env : "host");
    if (builder != "host" && builder != "gpu") {
        fprintf(stderr, "Warning: BVH builder \"%s\" unknown.  Using \"host\".\n", builder.c_str());
        builder = "host";
    }
    if (builder == "gpu" && splitMethod != "sah") {
        fprintf(stderr, "Warning: only splitmethod \"sah\" is built on the device.  Using the host build.\n");
        builder = "host";
    }
    return new GpuBVHAccel(prims, maxPrimsInNode, splitMethod, builder);
}

}  // namespace pb
