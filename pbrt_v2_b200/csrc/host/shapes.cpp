// shapes.cpp -- host-side scene preparation for the shapes on the hot path.
// Everything here runs once at scene build; the per-ray Intersect()/Sample()
// methods of these shapes live in the CUDA library.  Arithmetic order follows
// the cited pbrt-v2 routines so that world-space vertices, bounds, areas and
// subdivision limit positions/normals are bit-identical to the reference's.
#include "scene.h"

#include <cstdio>
#include <map>

namespace pb {

// ------------------------------------------------------------- TriangleMesh --
// shapes/trianglemesh.cpp:42-71: vertices are transformed to world space once.
TriangleMesh::TriangleMesh(const Transform *o2w, const Transform *w2o, bool ro, int nt, int nv,
                           const int *vi, const V3 *P, const V3 *N, const float *uv)
    : Shape(TRIANGLE_MESH, o2w, w2o, ro), ntris(nt), nverts(nv) {
    vertexIndex.assign(vi, vi + 3 * nt);
    if (uv) uvs.assign(uv, uv + 2 * nv);
    if (N) n.assign(N, N + nv);
    p.resize(nv);
    for (int i = 0; i < nv; ++i) p[i] = ObjectToWorld->Pt(P[i]);
}

// shapes/trianglemesh.cpp:371-441 ("vector S" tangents and alpha textures are refused by the caller, parser.cpp DoShape)
TriangleMesh *CreateTriangleMeshShape(const Transform *o2w, const Transform *w2o, bool ro,
                                      const ParamSet &params) {
    const std::vector<int> *vi = params.FindInt("indices");
    const std::vector<V3> *P = params.FindPoint("P");
    const std::vector<float> *uvs = params.FindFloat("uv");
    if (!uvs) uvs = params.FindFloat("st");
    if (!vi || !P) return nullptr;
    int npi = (int)P->size();
    if (uvs && (int)uvs->size() < 2 * npi) uvs = nullptr;
    const std::vector<V3> *N = params.FindNormal("N");
    if (N && (int)N->size() != npi) N = nullptr;
    if (params.FindOneBool("discarddegenerateUVs", false) && uvs && N) {   // trianglemesh.cpp:402-421
        // "if there are normals, check for bad uv's that give degenerate mappings; discard them if so"
        const std::vector<int> &v = *vi;
        for (size_t i = 0; i + 2 < v.size(); i += 3) {
            if (v[i] < 0 || v[i + 1] < 0 || v[i + 2] < 0 || v[i] >= npi || v[i + 1] >= npi || v[i + 2] >= npi) break;   // (refused below)
            const V3 &p0 = (*P)[v[i]], &p1 = (*P)[v[i + 1]], &p2 = (*P)[v[i + 2]];
            float area = .5f * Length(Cross(p0 - p1, p2 - p1));
            if (area < 1e-7) continue;
            const float *u = uvs->data();
            if ((u[2 * v[i]] == u[2 * v[i + 1]] && u[2 * v[i] + 1] == u[2 * v[i + 1] + 1]) ||
                (u[2 * v[i + 1]] == u[2 * v[i + 2]] && u[2 * v[i + 1] + 1] == u[2 * v[i + 2] + 1]) ||
                (u[2 * v[i + 2]] == u[2 * v[i]] && u[2 * v[i + 2] + 1] == u[2 * v[i] + 1])) {
                fprintf(stderr, "Warning: Degenerate uv coordinates in triangle mesh.  Discarding all uvs.\n");
                uvs = nullptr;
                break;
            }
        }
    }
    for (int idx : *vi)
        if (idx < 0 || idx >= npi) {   // (the reference checks the upper bound only, trianglemesh.cpp:386-391, and reads out of bounds below 0)
            fprintf(stderr, "trianglemesh has out of-bounds vertex index %d (%d \"P\" values were given)\n", idx, npi);
            return nullptr;
        }
    return new TriangleMesh(o2w, w2o, ro, (int)vi->size() / 3, npi, vi->data(), P->data(),
                            N ? N->data() : nullptr, uvs ? uvs->data() : nullptr);
}

// ------------------------------------------------------------------- Sphere --
Sphere::Sphere(const Transform *o2w, const Transform *w2o, bool ro, float rad, float z0, float z1, float pm)
    : Shape(SPHERE, o2w, w2o, ro) {   // shapes/sphere.cpp:39-49
    radius = rad;
    zmin = Clampf(std::min(z0, z1), -radius, radius);
    zmax = Clampf(std::max(z0, z1), -radius, radius);
    thetaMin = acosf(Clampf(zmin / radius, -1.f, 1.f));
    thetaMax = acosf(Clampf(zmax / radius, -1.f, 1.f));
    phiMax = Radians(Clampf(pm, 0.0f, 360.0f));
}
BBox Sphere::ObjectBound() const {   // sphere.cpp:52-55
    return BBox(V3(-radius, -radius, zmin), V3(radius, radius, zmax));
}
float Sphere::Area() const { return phiMax * radius * (zmax - zmin); }   // sphere.cpp:211-213

// --------------------------------------------------------------------- Disk --
Disk::Disk(const Transform *o2w, const Transform *w2o, bool ro, float ht, float r, float ri, float tmax)
    : Shape(DISK, o2w, w2o, ro) {   // shapes/disk.cpp:39-46
    height = ht;
    radius = r;
    innerRadius = ri;
    phiMax = Radians(Clampf(tmax, 0.0f, 360.0f));
}
BBox Disk::ObjectBound() const {   // disk.cpp:49-52
    return BBox(V3(-radius, -radius, height), V3(radius, radius, height));
}
float Disk::Area() const {   // disk.cpp:133-136
    return phiMax * 0.5f * (radius * radius - innerRadius * innerRadius);
}

// -------------------------------------------------------------- RefinedPrim --
BBox RefinedPrim::WorldBound() const {
    if (type == PBRTGPU_PRIM_ANIMATED) return animated->WorldBound();
    if (type == PBRTGPU_PRIM_TRIANGLE) {   // Triangle::WorldBound, trianglemesh.cpp:118-124
        const TriangleMesh *m = static_cast<const TriangleMesh *>(shape);
        const V3 &p1 = m->p[m->vertexIndex[3 * tri]];
        const V3 &p2 = m->p[m->vertexIndex[3 * tri + 1]];
        const V3 &p3 = m->p[m->vertexIndex[3 * tri + 2]];
        return Union(BBox(p1, p2), p3);
    }
    // Shape::WorldBound, core/shape.cpp:52-54
    if (type == PBRTGPU_PRIM_SPHERE)
        return shape->ObjectToWorld->Box(static_cast<const Sphere *>(shape)->ObjectBound());
    return shape->ObjectToWorld->Box(static_cast<const Disk *>(shape)->ObjectBound());
}

// --------------------------------------------------------------- LoopSubdiv --
// Index-based restatement of shapes/loopsubdiv.cpp.  The reference works on
// SDVertex*/SDFace* pointers; wherever it orders by pointer value (SDEdge keeps
// v[0] = min(ptr), loopsubdiv.cpp:100-103) we order by vertex index, which is the
// same thing for vertices that live in one array (all level-0 vertices,
// loopsubdiv.cpp:163-167) and the allocation order for arena-allocated ones.
namespace {
struct SDV {
    V3 P;
    int startFace = -1, child = -1;
    bool regular = false, boundary = false;
};
struct SDF {
    int v[3] = {-1, -1, -1};
    int f[3] = {-1, -1, -1};
    int children[4] = {-1, -1, -1, -1};
};
inline int NEXT(int i) { return (i + 1) % 3; }
inline int PREV(int i) { return (i + 2) % 3; }

struct Mesh {
    std::vector<SDV> V;
    std::vector<SDF> F;
    int vnum(int f, int vert) const {
        for (int i = 0; i < 3; ++i) if (F[f].v[i] == vert) return i;
        fprintf(stderr, "pbrt_b200: logic error in LoopSubdiv vnum\n");
        return 0;
    }
    int nextFace(int f, int vert) const { return F[f].f[vnum(f, vert)]; }
    int prevFace(int f, int vert) const { return F[f].f[PREV(vnum(f, vert))]; }
    int nextVert(int f, int vert) const { return F[f].v[NEXT(vnum(f, vert))]; }
    int prevVert(int f, int vert) const { return F[f].v[PREV(vnum(f, vert))]; }
    int otherVert(int f, int v0, int v1) const {
        for (int i = 0; i < 3; ++i) if (F[f].v[i] != v0 && F[f].v[i] != v1) return F[f].v[i];
        return -1;
    }
    int valence(int vert) const {   // loopsubdiv.cpp:117-136
        int f = V[vert].startFace;
        if (!V[vert].boundary) {
            int nf = 1;
            while ((f = nextFace(f, vert)) != V[vert].startFace) ++nf;
            return nf;
        }
        int nf = 1;
        while ((f = nextFace(f, vert)) != -1) ++nf;
        f = V[vert].startFace;
        while ((f = prevFace(f, vert)) != -1) ++nf;
        return nf + 1;
    }
    void oneRing(int vert, V3 *P) const {   // loopsubdiv.cpp:458-479
        if (!V[vert].boundary) {
            int face = V[vert].startFace;
            do {
                *P++ = V[nextVert(face, vert)].P;
                face = nextFace(face, vert);
            } while (face != V[vert].startFace);
        } else {
            int face = V[vert].startFace, f2;
            while ((f2 = nextFace(face, vert)) != -1) face = f2;
            *P++ = V[nextVert(face, vert)].P;
            do {
                *P++ = V[prevVert(face, vert)].P;
                face = prevFace(face, vert);
            } while (face != -1);
        }
    }
    V3 weightOneRing(int vert, float beta) const {   // loopsubdiv.cpp:447-456
        int val = valence(vert);
        std::vector<V3> ring(val);
        oneRing(vert, ring.data());
        V3 P = (1 - val * beta) * V[vert].P;
        for (int i = 0; i < val; ++i) P = P + beta * ring[i];
        return P;
    }
    V3 weightBoundary(int vert, float beta) const {   // loopsubdiv.cpp:481-490
        int val = valence(vert);
        std::vector<V3> ring(val);
        oneRing(vert, ring.data());
        V3 P = (1 - 2 * beta) * V[vert].P;
        P = P + beta * ring[0];
        P = P + beta * ring[val - 1];
        return P;
    }
};
inline float LoopBeta(int valence) {   // loopsubdiv.h:58-61
    if (valence == 3) return 3.f / 16.f;
    return 3.f / (8.f * valence);
}
inline float LoopGamma(int valence) {   // loopsubdiv.h:64-66
    return 1.f / (valence + 3.f / (8.f * LoopBeta(valence)));
}
typedef std::pair<int, int> EdgeKey;
inline EdgeKey MakeEdge(int a, int b) { return EdgeKey(std::min(a, b), std::max(a, b)); }
}  // namespace

LoopSubdiv::LoopSubdiv(const Transform *o2w, const Transform *w2o, bool ro, int nfaces, int nvertices,
                       const int *vi, const V3 *Pin, int nl)
    : Shape(LOOP_SUBDIV, o2w, w2o, ro), nLevels(nl) {
    indices.assign(vi, vi + 3 * nfaces);
    P.assign(Pin, Pin + nvertices);
}

namespace {
// LoopSubdiv::LoopSubdiv's topology pass, loopsubdiv.cpp:156-216
void ControlTopology(const std::vector<int> &indices, const std::vector<V3> &P, Mesh *out) {
    Mesh &cur = *out;
    int nverts = (int)P.size(), nfaces = (int)indices.size() / 3;
    cur.V.resize(nverts);
    for (int i = 0; i < nverts; ++i) cur.V[i].P = P[i];
    cur.F.resize(nfaces);
    for (int i = 0; i < nfaces; ++i)
        for (int j = 0; j < 3; ++j) {
            int v = indices[3 * i + j];
            cur.F[i].v[j] = v;
            cur.V[v].startFace = i;
        }
    {
        struct EdgeRec { int f0, f0edgeNum; };
        std::map<EdgeKey, EdgeRec> edges;
        for (int i = 0; i < nfaces; ++i)
            for (int e = 0; e < 3; ++e) {
                EdgeKey k = MakeEdge(cur.F[i].v[e], cur.F[i].v[NEXT(e)]);
                auto it = edges.find(k);
                if (it == edges.end()) {
                    edges[k] = EdgeRec{i, e};
                } else {
                    cur.F[it->second.f0].f[it->second.f0edgeNum] = i;
                    cur.F[i].f[e] = it->second.f0;
                    edges.erase(it);
                }
            }
    }
    for (int i = 0; i < nverts; ++i) {
        int f = cur.V[i].startFace;
        if (f < 0) { cur.V[i].boundary = true; continue; }   // unreferenced vertex
        do { f = cur.nextFace(f, i); } while (f != -1 && f != cur.V[i].startFace);
        cur.V[i].boundary = (f == -1);
        int val = cur.valence(i);
        if (!cur.V[i].boundary && val == 6) cur.V[i].regular = true;
        else if (cur.V[i].boundary && val == 4) cur.V[i].regular = true;
        else cur.V[i].regular = false;
    }
}
}  // namespace

// the control mesh as the index arrays of include/pbrtgpu.h's PbrtGpuSubdivMesh
void LoopSubdiv::Topology(std::vector<int32_t> *startFace, std::vector<uint8_t> *vflags, std::vector<int32_t> *faceV,
                          std::vector<int32_t> *faceF) const {
    Mesh cur;
    ControlTopology(indices, P, &cur);
    startFace->resize(cur.V.size());
    vflags->resize(cur.V.size());
    for (size_t i = 0; i < cur.V.size(); ++i) {
        (*startFace)[i] = cur.V[i].startFace;
        (*vflags)[i] = (uint8_t)((cur.V[i].boundary ? 1 : 0) | (cur.V[i].regular ? 2 : 0));
    }
    faceV->resize(3 * cur.F.size());
    faceF->resize(3 * cur.F.size());
    for (size_t j = 0; j < cur.F.size(); ++j)
        for (int k = 0; k < 3; ++k) { (*faceV)[3 * j + k] = cur.F[j].v[k]; (*faceF)[3 * j + k] = cur.F[j].f[k]; }
}

// LoopSubdiv::Refine on the device (include/pbrtgpu.h: pbrtgpu_loopsubdiv_refine): same TriangleMesh
TriangleMesh *LoopSubdiv::RefineOnDevice(std::string *err) const {
    std::vector<int32_t> startFace, faceV, faceF;
    std::vector<uint8_t> vflags;
    Topology(&startFace, &vflags, &faceV, &faceF);
    std::vector<float> Pin(3 * P.size());
    for (size_t i = 0; i < P.size(); ++i) { Pin[3 * i] = P[i].x; Pin[3 * i + 1] = P[i].y; Pin[3 * i + 2] = P[i].z; }
    PbrtGpuSubdivMesh m;
    m.nVertices = (uint32_t)P.size(); m.nFaces = (uint32_t)(indices.size() / 3);
    m.P = Pin.data(); m.startFace = startFace.data(); m.vflags = vflags.data(); m.faceV = faceV.data(); m.faceF = faceF.data();
    uint32_t nv = 0, nt = 0;
    if (pbrtgpu_loopsubdiv_counts(&m, nLevels, &nv, &nt) != PBRTGPU_OK) { *err = pbrtgpu_last_error(); return nullptr; }
    std::vector<float> Pl(3 * (size_t)nv), Nl(3 * (size_t)nv);
    std::vector<int32_t> verts(3 * (size_t)nt);
    if (pbrtgpu_loopsubdiv_refine(&m, nLevels, Pl.data(), Nl.data(), verts.data()) != PBRTGPU_OK) { *err = pbrtgpu_last_error(); return nullptr; }
    std::vector<V3> Pv(nv), Nv(nv);
    for (uint32_t i = 0; i < nv; ++i) { Pv[i] = V3(Pl[3 * i], Pl[3 * i + 1], Pl[3 * i + 2]); Nv[i] = V3(Nl[3 * i], Nl[3 * i + 1], Nl[3 * i + 2]); }
    return new TriangleMesh(ObjectToWorld, WorldToObject, ReverseOrientation, (int)nt, (int)nv, verts.data(), Pv.data(), Nv.data(), nullptr);
}

TriangleMesh *LoopSubdiv::Refine() const {
    // "string refiner" "gpu" on the shape, or PBRTGPU_SUBDIV=gpu: the levels run on the device.  No silent fallback.
    if (refineOnDevice) {
        std::string err;
        TriangleMesh *mesh = RefineOnDevice(&err);
        if (!mesh) refineError = "LoopSubdiv on the device failed: " + err;
        return mesh;   // null on failure: the accelerator's build reports refineError
    }
    // ---- control mesh topology, loopsubdiv.cpp:156-216 ----
    Mesh cur;
    ControlTopology(indices, P, &cur);

    // ---- subdivision levels, loopsubdiv.cpp:240-345 ----
    for (int level = 0; level < nLevels; ++level) {
        Mesh nxt;
        int nv = (int)cur.V.size(), nf = (int)cur.F.size();
        nxt.V.resize(nv);
        for (int j = 0; j < nv; ++j) {
            cur.V[j].child = j;
            nxt.V[j].regular = cur.V[j].regular;
            nxt.V[j].boundary = cur.V[j].boundary;
        }
        nxt.F.resize(4 * nf);
        for (int j = 0; j < nf; ++j)
            for (int k = 0; k < 4; ++k) cur.F[j].children[k] = 4 * j + k;
        // even vertices
        for (int j = 0; j < nv; ++j) {
            if (cur.V[j].startFace < 0) { nxt.V[j].P = cur.V[j].P; continue; }
            if (!cur.V[j].boundary) {
                if (cur.V[j].regular) nxt.V[j].P = cur.weightOneRing(j, 1.f / 16.f);
                else nxt.V[j].P = cur.weightOneRing(j, LoopBeta(cur.valence(j)));
            } else {
                nxt.V[j].P = cur.weightBoundary(j, 1.f / 8.f);
            }
        }
        // odd (edge) vertices
        std::map<EdgeKey, int> edgeVerts;
        for (int j = 0; j < nf; ++j)
            for (int k = 0; k < 3; ++k) {
                EdgeKey e = MakeEdge(cur.F[j].v[k], cur.F[j].v[NEXT(k)]);
                if (edgeVerts.count(e)) continue;
                SDV vert;
                vert.regular = true;
                vert.boundary = (cur.F[j].f[k] == -1);
                vert.startFace = cur.F[j].children[3];
                const V3 &P0 = cur.V[e.first].P, &P1 = cur.V[e.second].P;
                if (vert.boundary) {
                    vert.P = 0.5f * P0;
                    vert.P = vert.P + 0.5f * P1;
                } else {
                    vert.P = 3.f / 8.f * P0;
                    vert.P = vert.P + 3.f / 8.f * P1;
                    vert.P = vert.P + 1.f / 8.f * cur.V[cur.otherVert(j, e.first, e.second)].P;
                    vert.P = vert.P + 1.f / 8.f * cur.V[cur.otherVert(cur.F[j].f[k], e.first, e.second)].P;
                }
                nxt.V.push_back(vert);
                edgeVerts[e] = (int)nxt.V.size() - 1;
            }
        // even vertex face pointers
        for (int j = 0; j < nv; ++j) {
            if (cur.V[j].startFace < 0) continue;
            int vertNum = cur.vnum(cur.V[j].startFace, j);
            nxt.V[j].startFace = cur.F[cur.V[j].startFace].children[vertNum];
        }
        // face neighbour pointers
        for (int j = 0; j < nf; ++j) {
            const SDF &face = cur.F[j];
            for (int k = 0; k < 3; ++k) {
                nxt.F[face.children[3]].f[k] = face.children[NEXT(k)];
                nxt.F[face.children[k]].f[NEXT(k)] = face.children[3];
                int f2 = face.f[k];
                nxt.F[face.children[k]].f[k] = f2 != -1 ? cur.F[f2].children[cur.vnum(f2, face.v[k])] : -1;
                f2 = face.f[PREV(k)];
                nxt.F[face.children[k]].f[PREV(k)] = f2 != -1 ? cur.F[f2].children[cur.vnum(f2, face.v[k])] : -1;
            }
        }
        // face vertex pointers
        for (int j = 0; j < nf; ++j) {
            const SDF &face = cur.F[j];
            for (int k = 0; k < 3; ++k) {
                nxt.F[face.children[k]].v[k] = cur.V[face.v[k]].child;
                int vert = edgeVerts[MakeEdge(face.v[k], face.v[NEXT(k)])];
                nxt.F[face.children[k]].v[NEXT(k)] = vert;
                nxt.F[face.children[NEXT(k)]].v[k] = vert;
                nxt.F[face.children[3]].v[k] = vert;
            }
        }
        cur = std::move(nxt);
    }

    // ---- limit surface positions, loopsubdiv.cpp:347-356 ----
    int nv = (int)cur.V.size();
    std::vector<V3> Plimit(nv);
    for (int i = 0; i < nv; ++i) {
        if (cur.V[i].startFace < 0) { Plimit[i] = cur.V[i].P; continue; }
        if (cur.V[i].boundary) Plimit[i] = cur.weightBoundary(i, 1.f / 5.f);
        else Plimit[i] = cur.weightOneRing(i, LoopGamma(cur.valence(i)));
    }
    for (int i = 0; i < nv; ++i) cur.V[i].P = Plimit[i];

    // ---- limit surface normals, loopsubdiv.cpp:358-402 ----
    std::vector<V3> Ns;
    Ns.reserve(nv);
    std::vector<V3> Pring(16);
    for (int i = 0; i < nv; ++i) {
        V3 S(0, 0, 0), T(0, 0, 0);
        if (cur.V[i].startFace < 0) { Ns.push_back(V3(0, 0, 0)); continue; }
        int valence = cur.valence(i);
        if (valence > (int)Pring.size()) Pring.resize(valence);
        cur.oneRing(i, Pring.data());
        if (!cur.V[i].boundary) {
            for (int k = 0; k < valence; ++k) {
                S = S + cosf(2.f * kPi * k / valence) * Pring[k];
                T = T + sinf(2.f * kPi * k / valence) * Pring[k];
            }
        } else {
            S = Pring[valence - 1] - Pring[0];
            if (valence == 2) T = Pring[0] + Pring[1] - 2 * cur.V[i].P;
            else if (valence == 3) T = Pring[1] - cur.V[i].P;
            else if (valence == 4)
                T = -1 * Pring[0] + 2 * Pring[1] + 2 * Pring[2] + -1 * Pring[3] + -2 * cur.V[i].P;
            else {
                float theta = kPi / float(valence - 1);
                T = sinf(theta) * (Pring[0] + Pring[valence - 1]);
                for (int k = 1; k < valence - 1; ++k) {
                    float wt = (2 * cosf(theta) - 2) * sinf((k)*theta);
                    T = T + wt * Pring[k];
                }
                T = -T;
            }
        }
        Ns.push_back(Cross(S, T));
    }

    // ---- emit the TriangleMesh, loopsubdiv.cpp:404-422 ----
    int ntris = (int)cur.F.size();
    std::vector<int> verts(3 * ntris);
    for (int i = 0; i < ntris; ++i)
        for (int j = 0; j < 3; ++j) verts[3 * i + j] = cur.F[i].v[j];
    return new TriangleMesh(ObjectToWorld, WorldToObject, ReverseOrientation, ntris, nv, verts.data(),
                            Plimit.data(), Ns.data(), nullptr);
}

}  // namespace pb
