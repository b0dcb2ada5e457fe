// oracle.cpp -- TEST INFRASTRUCTURE: CPU restatement of pbrt-v2's path-tracing
// hot path over the flattened scene (include/pbrtgpu.h PODs).
//
// This file is the CHECKER for the CUDA path.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline leg may load it; the product (libpbrtgpu.so,
// libpbrthost.so, pbrt_v2_b200/*.py) never does.  It is pinned against the
// UNMODIFIED reference binary (oracle/_ref/bin/pbrt, built by oracle/Makefile):
// tests/test_oracle_vs_reference.py requires bit-identical PFM output at equal
// nTasks on the bunny and killeroo config scenes.
//
// Structure follows the reference function by function (file:line cited at each);
// it is scalar, recursive where the reference is, and uses the host libm.
// Build: g++ -O2 -ffp-contract=off (no -march=native: no FMA).
#include "oracle_math.h"
#include "../../include/pbrtgpu.h"

#include <atomic>
#include <cstdio>
#include <thread>
#include <vector>

using namespace orc;

namespace {

// =============================================================== RNG (MT19937)
// core/rng.h:44-59, core/rng.cpp:43-107
struct RNG {
    static const int N = 624;
    uint32_t mt[N];
    int mti;
    uint64_t draws;
    explicit RNG(uint32_t seed) { Seed(seed); draws = 0; }
    void Seed(uint32_t seed) {
        mt[0] = seed;
        for (mti = 1; mti < N; mti++)
            mt[mti] = 1812433253u * (mt[mti - 1] ^ (mt[mti - 1] >> 30)) + (uint32_t)mti;
    }
    uint32_t RandomUInt() {
        const int M = 397;
        if (mti >= N) {
            int kk;
            uint32_t y;
            for (kk = 0; kk < N - M; kk++) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            for (; kk < N - 1; kk++) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            mti = 0;
        }
        uint32_t y = mt[mti++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        ++draws;
        return y;
    }
    float RandomFloat() { return (RandomUInt() & 0xffffff) / float(1 << 24); }
};

// ========================================================= low-discrepancy tables
// core/montecarlo.h:274-327
inline float VanDerCorput(uint32_t n, uint32_t scramble) {
    n = (n << 16) | (n >> 16);
    n = ((n & 0x00ff00ff) << 8) | ((n & 0xff00ff00) >> 8);
    n = ((n & 0x0f0f0f0f) << 4) | ((n & 0xf0f0f0f0) >> 4);
    n = ((n & 0x33333333) << 2) | ((n & 0xcccccccc) >> 2);
    n = ((n & 0x55555555) << 1) | ((n & 0xaaaaaaaa) >> 1);
    n ^= scramble;
    return std::min(((n >> 8) & 0xffffff) / float(1 << 24), OneMinusEpsilon);
}
inline float Sobol2(uint32_t n, uint32_t scramble) {
    for (uint32_t v = 1u << 31; n != 0; n >>= 1, v ^= v >> 1)
        if (n & 0x1) scramble ^= v;
    return std::min(((scramble >> 8) & 0xffffff) / float(1 << 24), OneMinusEpsilon);
}
void Shuffle(float *samp, uint32_t count, uint32_t dims, RNG &rng) {   // montecarlo.h:174-181
    for (uint32_t i = 0; i < count; ++i) {
        uint32_t other = i + (rng.RandomUInt() % (count - i));
        for (uint32_t j = 0; j < dims; ++j) std::swap(samp[dims * i + j], samp[dims * other + j]);
    }
}
void LDShuffleScrambled1D(int nSamples, int nPixel, float *samples, RNG &rng) {
    uint32_t scramble = rng.RandomUInt();
    for (int i = 0; i < nSamples * nPixel; ++i) samples[i] = VanDerCorput(i, scramble);
    for (int i = 0; i < nPixel; ++i) Shuffle(samples + i * nSamples, nSamples, 1, rng);
    Shuffle(samples, nPixel, nSamples, rng);
}
void LDShuffleScrambled2D(int nSamples, int nPixel, float *samples, RNG &rng) {
    uint32_t s0 = rng.RandomUInt();
    uint32_t s1 = rng.RandomUInt();
    for (int i = 0; i < nSamples * nPixel; ++i) {
        samples[2 * i] = VanDerCorput(i, s0);
        samples[2 * i + 1] = Sobol2(i, s1);
    }
    for (int i = 0; i < nPixel; ++i) Shuffle(samples + 2 * i * nSamples, nSamples, 2, rng);
    Shuffle(samples, nPixel, 2 * nSamples, rng);
}

// struct Sample (core/sampler.h:93-116) with the integrators' requested arrays.  PathIntegrator +
// EmissionIntegrator (path.cpp:43-48, light.cpp:64-68, reflection.cpp:502-506, emission.cpp:41-42):
//   1-D per depth d<3: [4d+0] light.component [4d+1] lightNum [4d+2] bsdf.component [4d+3] path.component
//       [12] emission tau  [13] emission scatter          2-D per depth d<3: [3d+0] light.pos [3d+1] bsdf.dir [3d+2] path.dir
// DirectLightingIntegrator (directlighting.cpp:53-78), strategy "all", light i with n_i = RoundSize(nSamples):
//   1-D [2i] light.component x n_i  [2i+1] bsdf.component x n_i, then tau, scatter;  2-D [2i] light.pos x n_i  [2i+1] bsdf.dir x n_i
// strategy "one": 1-D [0] light.component [1] lightNum [2] bsdf.component, tau, scatter;  2-D [0] light.pos [1] bsdf.dir
enum { MAX_ARRAYS = PBRTGPU_MAX_SAMPLE_ARRAYS };
struct SampleLayout {
    int n1D, n2D;
    uint32_t size1D[MAX_ARRAYS], size2D[MAX_ARRAYS];
    void FromDesc(const PbrtGpuRenderDesc *rd) {
        n1D = rd->n1D; n2D = rd->n2D;
        for (int i = 0; i < n1D; ++i) size1D[i] = rd->size1D[i];
        for (int i = 0; i < n2D; ++i) size2D[i] = rd->size2D[i];
    }
    size_t Floats() const {   // per pixel sample, camera part included
        size_t n = 5;
        for (int i = 0; i < n1D; ++i) n += size1D[i];
        for (int i = 0; i < n2D; ++i) n += 2 * size2D[i];
        return n;
    }
};
struct Sample {
    float imageX, imageY, lensU, lensV, time;
    float *oneD[MAX_ARRAYS];
    float *twoD[MAX_ARRAYS];
};

// LDPixelSample, core/montecarlo.cpp:200-252.  `store`: Floats()-5 floats per pixel sample behind oneD/twoD.
void LDPixelSample(const SampleLayout &lay, int xPos, int yPos, float shutterOpen, float shutterClose, int nPixelSamples,
                   Sample *samples, float *store, float *buf, RNG &rng) {
    float *imageSamples = buf; buf += 2 * nPixelSamples;
    float *lensSamples = buf;  buf += 2 * nPixelSamples;
    float *timeSamples = buf;  buf += nPixelSamples;
    float *oneDSamples[MAX_ARRAYS], *twoDSamples[MAX_ARRAYS];
    for (int i = 0; i < lay.n1D; ++i) { oneDSamples[i] = buf; buf += lay.size1D[i] * nPixelSamples; }
    for (int i = 0; i < lay.n2D; ++i) { twoDSamples[i] = buf; buf += 2 * lay.size2D[i] * nPixelSamples; }
    LDShuffleScrambled2D(1, nPixelSamples, imageSamples, rng);
    LDShuffleScrambled2D(1, nPixelSamples, lensSamples, rng);
    LDShuffleScrambled1D(1, nPixelSamples, timeSamples, rng);
    for (int i = 0; i < lay.n1D; ++i) LDShuffleScrambled1D(lay.size1D[i], nPixelSamples, oneDSamples[i], rng);
    for (int i = 0; i < lay.n2D; ++i) LDShuffleScrambled2D(lay.size2D[i], nPixelSamples, twoDSamples[i], rng);
    const size_t per = lay.Floats() - 5;
    for (int i = 0; i < nPixelSamples; ++i) {
        samples[i].imageX = xPos + imageSamples[2 * i];
        samples[i].imageY = yPos + imageSamples[2 * i + 1];
        samples[i].time = Lerp(timeSamples[i], shutterOpen, shutterClose);
        samples[i].lensU = lensSamples[2 * i];
        samples[i].lensV = lensSamples[2 * i + 1];
        float *mem = store + per * i;
        for (int j = 0; j < lay.n1D; ++j) {
            samples[i].oneD[j] = mem;
            int startSamp = lay.size1D[j] * i;
            for (uint32_t k = 0; k < lay.size1D[j]; ++k) *mem++ = oneDSamples[j][startSamp + k];
        }
        for (int j = 0; j < lay.n2D; ++j) {
            samples[i].twoD[j] = mem;
            int startSamp = 2 * lay.size2D[j] * i;
            for (uint32_t k = 0; k < 2 * lay.size2D[j]; ++k) *mem++ = twoDSamples[j][startSamp + k];
        }
    }
}

// ================================================================== geometry
struct Ray {
    Vec o, d;
    mutable float mint, maxt;
    float time;
    Ray() : mint(0.f), maxt(INF_F), time(0.f) {}
    Ray(const Vec &oo, const Vec &dd, float start, float end = INF_F, float t = 0.f)
        : o(oo), d(dd), mint(start), maxt(end), time(t) {}
    Vec operator()(float t) const { return o + d * t; }   // geometry.h:333
};

struct DiffGeom {   // core/diffgeom.h:44-65 (screen-space differentials omitted: they
    Vec p, nn, dpdu, dpdv;   // only feed textures, constant in scope)
    float u, v;
};
// DifferentialGeometry ctor, core/diffgeom.cpp:40-55
inline void InitDG(DiffGeom *dg, const Vec &P, const Vec &DPDU, const Vec &DPDV, float uu, float vv,
                   uint32_t shapeFlags) {
    dg->p = P; dg->dpdu = DPDU; dg->dpdv = DPDV;
    dg->nn = Normalize(Cross(DPDU, DPDV));
    dg->u = uu; dg->v = vv;
    bool ro = (shapeFlags & PBRTGPU_SHAPE_REVERSE_ORIENT) != 0, sh = (shapeFlags & PBRTGPU_SHAPE_SWAPS_HANDED) != 0;
    if (ro ^ sh) dg->nn = dg->nn * -1.f;
}

struct Counters {
    uint64_t closestRays = 0, shadowRays = 0, nodesVisited = 0, primTests = 0, shadedN = 0, rngDraws = 0;
    uint64_t closestNodes = 0, closestPrims = 0;
};

struct Intersection {   // core/intersection.h:45-62
    DiffGeom dg;
    uint32_t prim;
    float rayEpsilon;
    bool instanced;   // hit under a TransformedPrimitive whose interpolated transform is not the identity
    Xf w2p;           // ... that transform (isect->WorldToObject = WorldToObject * w2p, primitive.cpp:106)
    Intersection() : prim(0), rayEpsilon(0.f), instanced(false) {}
};

struct SceneView {
    const PbrtGpuSceneDesc *d;

    void TriVerts(const PbrtGpuTriangle &t, Vec *p1, Vec *p2, Vec *p3) const {
        *p1 = Vec(d->P + 3 * t.v[0]); *p2 = Vec(d->P + 3 * t.v[1]); *p3 = Vec(d->P + 3 * t.v[2]);
    }
    void GetUVs(const PbrtGpuTriangle &t, float uv[3][2]) const {   // trianglemesh.h:86-101
        if (d->meshes[t.mesh].flags & PBRTGPU_MESH_HAS_UV) {
            for (int i = 0; i < 3; ++i) { uv[i][0] = d->uv[2 * t.v[i]]; uv[i][1] = d->uv[2 * t.v[i] + 1]; }
        } else {
            uv[0][0] = 0.; uv[0][1] = 0.; uv[1][0] = 1.; uv[1][1] = 0.; uv[2][0] = 1.; uv[2][1] = 1.;
        }
    }

    // Triangle::Intersect, shapes/trianglemesh.cpp:127-208
    bool TriIntersect(const PbrtGpuTriangle &tri, const Ray &ray, float *tHit, float *rayEpsilon, DiffGeom *dg,
                      float *ob1 = nullptr, float *ob2 = nullptr) const {
        Vec p1, p2, p3;
        TriVerts(tri, &p1, &p2, &p3);
        Vec e1 = p2 - p1;
        Vec e2 = p3 - p1;
        Vec s1 = Cross(ray.d, e2);
        float divisor = Dot(s1, e1);
        if (divisor == 0.) return false;
        float invDivisor = 1.f / divisor;
        Vec s = ray.o - p1;
        float b1 = Dot(s, s1) * invDivisor;
        if (b1 < 0. || b1 > 1.) return false;
        Vec s2 = Cross(s, e1);
        float b2 = Dot(ray.d, s2) * invDivisor;
        if (b2 < 0. || b1 + b2 > 1.) return false;
        float t = Dot(e2, s2) * invDivisor;
        if (t < ray.mint || t > ray.maxt) return false;
        if (dg) {
            Vec dpdu, dpdv;
            float uvs[3][2];
            GetUVs(tri, uvs);
            float du1 = uvs[0][0] - uvs[2][0];
            float du2 = uvs[1][0] - uvs[2][0];
            float dv1 = uvs[0][1] - uvs[2][1];
            float dv2 = uvs[1][1] - uvs[2][1];
            Vec dp1 = p1 - p3, dp2 = p2 - p3;
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
            InitDG(dg, ray(t), dpdu, dpdv, tu, tv, d->meshes[tri.mesh].flags);
        }
        *tHit = t;
        if (rayEpsilon) *rayEpsilon = 1e-3f * *tHit;
        if (ob1) { *ob1 = b1; *ob2 = b2; }
        return true;
    }

    // Sphere::Intersect / IntersectP, shapes/sphere.cpp:58-209
    bool SphereIntersect(const PbrtGpuQuadric &q, const Ray &r, float *tHit, float *rayEpsilon, DiffGeom *dg) const {
        const PbrtGpuXform &x = d->xforms[q.xform];
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
        Vec phit = ray(thit);
        if (phit.x == 0.f && phit.y == 0.f) phit.x = 1e-5f * radius;
        float phi = atan2f(phit.y, phit.x);
        if (phi < 0.) phi += 2.f * PI_F;
        if ((zmin > -radius && phit.z < zmin) || (zmax < radius && phit.z > zmax) || phi > phiMax) {
            if (thit == t1) return false;
            if (t1 > ray.maxt) return false;
            thit = t1;
            phit = ray(thit);
            if (phit.x == 0.f && phit.y == 0.f) phit.x = 1e-5f * radius;
            phi = atan2f(phit.y, phit.x);
            if (phi < 0.) phi += 2.f * PI_F;
            if ((zmin > -radius && phit.z < zmin) || (zmax < radius && phit.z > zmax) || phi > phiMax) return false;
        }
        if (dg) {
            float u = phi / phiMax;
            float theta = acosf(Clampf(phit.z / radius, -1.f, 1.f));
            float v = (theta - q.thetaMin) / (q.thetaMax - q.thetaMin);
            float zradius = sqrtf(phit.x * phit.x + phit.y * phit.y);
            float invzradius = 1.f / zradius;
            float cosphi = phit.x * invzradius;
            float sinphi = phit.y * invzradius;
            Vec dpdu(-phiMax * phit.y, phiMax * phit.x, 0);
            Vec dpdv = (q.thetaMax - q.thetaMin) * Vec(phit.z * cosphi, phit.z * sinphi, -radius * sinf(theta));
            // dndu/dndv (sphere.cpp:116-136) only feed bump/texture code: omitted
            InitDG(dg, XfPoint(x.m, phit), XfVector(x.m, dpdu), XfVector(x.m, dpdv), u, v, q.flags);
        }
        *tHit = thit;
        if (rayEpsilon) *rayEpsilon = 5e-4f * *tHit;
        return true;
    }

    // Disk::Intersect / IntersectP, shapes/disk.cpp:56-130
    bool DiskIntersect(const PbrtGpuQuadric &q, const Ray &r, float *tHit, float *rayEpsilon, DiffGeom *dg) const {
        const PbrtGpuXform &x = d->xforms[q.xform];
        Ray ray;
        ray.o = XfPoint(x.mInv, r.o); ray.d = XfVector(x.mInv, r.d);
        ray.mint = r.mint; ray.maxt = r.maxt;
        if (fabsf(ray.d.z) < 1e-7) return false;
        float thit = (q.height - ray.o.z) / ray.d.z;
        if (thit < ray.mint || thit > ray.maxt) return false;
        Vec phit = ray(thit);
        float dist2 = phit.x * phit.x + phit.y * phit.y;
        if (dist2 > q.radius * q.radius || dist2 < q.innerRadius * q.innerRadius) return false;
        float phi = atan2f(phit.y, phit.x);
        if (phi < 0) phi += 2. * PI_F;   // NB double arithmetic (disk.cpp:77)
        if (phi > q.phiMax) return false;
        if (dg) {
            float u = phi / q.phiMax;
            float R = sqrtf(dist2);
            float oneMinusV = ((R - q.innerRadius) / (q.radius - q.innerRadius));
            float v = 1.f - oneMinusV;
            Vec dpdu(-q.phiMax * phit.y, q.phiMax * phit.x, 0.);
            Vec dpdv(phit.x, phit.y, 0.);
            dpdv = dpdv * ((q.radius - q.innerRadius) / R);
            InitDG(dg, XfPoint(x.m, phit), XfVector(x.m, dpdu), XfVector(x.m, dpdv), u, v, q.flags);
        }
        *tHit = thit;
        if (rayEpsilon) *rayEpsilon = 5e-4f * *tHit;
        return true;
    }

    bool QuadricIntersect(const PbrtGpuQuadric &q, const Ray &r, float *tHit, float *rayEps, DiffGeom *dg) const {
        return q.type == PBRTGPU_PRIM_SPHERE ? SphereIntersect(q, r, tHit, rayEps, dg) : DiskIntersect(q, r, tHit, rayEps, dg);
    }

    // Slerp, core/quaternion.cpp:99-111; quaternion = (v.x, v.y, v.z, w)
    static void Slerp(float t, const float *q1, const float *q2, float *out) {
        auto qdot = [](const float *a, const float *b) { return (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) + a[3] * b[3]; };
        auto qnormalize = [&](float *q) {   // quaternion.h:109-111: v /= f multiplies by 1/f, w /= f divides
            float f = sqrtf(qdot(q, q));
            float inv = 1.f / f;
            q[0] *= inv; q[1] *= inv; q[2] *= inv;
            q[3] /= f;
        };
        float cosTheta = qdot(q1, q2);
        if (cosTheta > .9995f) {
            for (int i = 0; i < 4; ++i) out[i] = q1[i] * (1.f - t) + q2[i] * t;
            qnormalize(out);
        } else {
            float theta = acosf(Clampf(cosTheta, -1.f, 1.f));
            float thetap = theta * t;
            float qperp[4];
            for (int i = 0; i < 4; ++i) qperp[i] = q2[i] - q1[i] * cosTheta;
            qnormalize(qperp);
            float c = cosf(thetap), sn = sinf(thetap);
            for (int i = 0; i < 4; ++i) out[i] = q1[i] * c + qperp[i] * sn;
        }
    }
    // AnimatedTransform::Interpolate, core/transform.cpp:371-396
    static Xf Interpolate(const PbrtGpuAnimated &a, float time) {
        Xf r;
        if (!a.actuallyAnimated || time <= a.startTime) { memcpy(&r, &a.start, sizeof(r)); return r; }
        if (time >= a.endTime) { memcpy(&r, &a.end, sizeof(r)); return r; }
        float dt = (time - a.startTime) / (a.endTime - a.startTime);
        float trans[3];
        for (int i = 0; i < 3; ++i) trans[i] = (1.f - dt) * a.T[0][i] + dt * a.T[1][i];
        float q[4];
        Slerp(dt, a.R[0], a.R[1], q);
        Xf scale;
        MatIdentity(scale.m);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) scale.m[4 * i + j] = Lerp(dt, a.S[0][4 * i + j], a.S[1][4 * i + j]);
        MatInverse(scale.m, scale.mInv);   // Transform(const Matrix4x4 &), transform.h:99-101
        Xf tr;                             // Translate(), transform.cpp:160-170
        MatIdentity(tr.m); MatIdentity(tr.mInv);
        tr.m[3] = trans[0]; tr.m[7] = trans[1]; tr.m[11] = trans[2];
        tr.mInv[3] = -trans[0]; tr.mInv[7] = -trans[1]; tr.mInv[11] = -trans[2];
        Xf rot;                            // Quaternion::ToTransform, quaternion.cpp:39-58
        {
            const float x = q[0], y = q[1], z = q[2], w = q[3];
            float xx = x * x, yy = y * y, zz = z * z;
            float xy = x * y, xz = x * z, yz = y * z;
            float wx = x * w, wy = y * w, wz = z * w;
            float m[16];
            MatIdentity(m);
            m[0] = 1.f - 2.f * (yy + zz);
            m[1] = 2.f * (xy + wz);
            m[2] = 2.f * (xz - wy);
            m[4] = 2.f * (xy - wz);
            m[5] = 1.f - 2.f * (xx + zz);
            m[6] = 2.f * (yz + wx);
            m[8] = 2.f * (xz + wy);
            m[9] = 2.f * (yz - wx);
            m[10] = 1.f - 2.f * (xx + yy);
            MatTranspose(m, rot.m);
            memcpy(rot.mInv, m, sizeof(m));
        }
        return XfCompose(XfCompose(tr, rot), scale);
    }

    // TransformedPrimitive::Intersect, core/primitive.cpp:95-119
    bool AnimatedIntersect(const PbrtGpuAnimated &a, const Ray &r, Intersection *isect, Counters *cnt, float *b1, float *b2) const {
        Xf w2p = Interpolate(a, r.time);
        Ray ray = r;
        ray.o = XfPoint(w2p.m, r.o);
        ray.d = XfVector(w2p.m, r.d);
        bool hit;
        if (a.nNodes > 0) hit = IntersectBVH(a.firstNode, a.firstPrimitive, ray, isect, cnt, b1, b2);
        else hit = PrimIntersect(a.firstPrimitive, ray, isect, cnt, b1, b2);
        if (!hit) return false;
        r.maxt = ray.maxt;
        isect->instanced = false;
        if (!XfIsIdentity(w2p)) {
            isect->instanced = true;
            isect->w2p = w2p;
            // PrimitiveToWorld = Inverse(w2p): points/vectors by w2p.mInv, normals by the transpose of w2p.m
            isect->dg.p = XfPoint(w2p.mInv, isect->dg.p);
            isect->dg.nn = Normalize(XfNormal(w2p.m, isect->dg.nn));
            isect->dg.dpdu = XfVector(w2p.mInv, isect->dg.dpdu);
            isect->dg.dpdv = XfVector(w2p.mInv, isect->dg.dpdv);
        }
        return true;
    }
    // TransformedPrimitive::IntersectP, core/primitive.cpp:122-124
    bool AnimatedIntersectP(const PbrtGpuAnimated &a, const Ray &r, Counters *cnt) const {
        Xf w2p = Interpolate(a, r.time);   // AnimatedTransform::operator()(Ray): same three cases (transform.cpp:416-427)
        Ray ray = r;
        ray.o = XfPoint(w2p.m, r.o);
        ray.d = XfVector(w2p.m, r.d);
        if (a.nNodes > 0) return IntersectPBVH(a.firstNode, a.firstPrimitive, ray, cnt);
        return PrimIntersectP(a.firstPrimitive, ray, cnt);
    }

    // GeometricPrimitive::Intersect, core/primitive.cpp:163-177
    bool PrimIntersect(uint32_t pi, const Ray &r, Intersection *isect, Counters *cnt, float *b1 = nullptr, float *b2 = nullptr) const {
        const PbrtGpuPrimitive &p = d->primitives[pi];
        if (p.type == PBRTGPU_PRIM_ANIMATED) return AnimatedIntersect(d->animated[p.index], r, isect, cnt, b1, b2);
        float thit, rayEpsilon;
        bool hit;
        if (p.type == PBRTGPU_PRIM_TRIANGLE) hit = TriIntersect(d->triangles[p.index], r, &thit, &rayEpsilon, &isect->dg, b1, b2);
        else {
            hit = QuadricIntersect(d->quadrics[p.index], r, &thit, &rayEpsilon, &isect->dg);
            if (hit && b1) { *b1 = 0.f; *b2 = 0.f; }
        }
        if (!hit) return false;
        isect->prim = pi;
        isect->rayEpsilon = rayEpsilon;
        isect->instanced = false;
        r.maxt = thit;
        return true;
    }
    bool PrimIntersectP(uint32_t pi, const Ray &r, Counters *cnt) const {
        const PbrtGpuPrimitive &p = d->primitives[pi];
        if (p.type == PBRTGPU_PRIM_ANIMATED) return AnimatedIntersectP(d->animated[p.index], r, cnt);
        float thit;
        if (p.type == PBRTGPU_PRIM_TRIANGLE) return TriIntersect(d->triangles[p.index], r, &thit, nullptr, nullptr);
        return QuadricIntersect(d->quadrics[p.index], r, &thit, nullptr, nullptr);
    }

    // ::IntersectP(BBox, Ray, invDir, dirIsNeg), accelerators/bvh.cpp:126-148
    static inline bool BoxTest(const PbrtGpuBVHNode &n, const Ray &ray, const Vec &invDir, const uint32_t dirIsNeg[3]) {
        const float *b0 = n.bmin, *b1 = n.bmax;
        float tmin = ((dirIsNeg[0] ? b1 : b0)[0] - ray.o.x) * invDir.x;
        float tmax = ((dirIsNeg[0] ? b0 : b1)[0] - ray.o.x) * invDir.x;
        float tymin = ((dirIsNeg[1] ? b1 : b0)[1] - ray.o.y) * invDir.y;
        float tymax = ((dirIsNeg[1] ? b0 : b1)[1] - ray.o.y) * invDir.y;
        if ((tmin > tymax) || (tymin > tmax)) return false;
        if (tymin > tmin) tmin = tymin;
        if (tymax < tmax) tmax = tymax;
        float tzmin = ((dirIsNeg[2] ? b1 : b0)[2] - ray.o.z) * invDir.z;
        float tzmax = ((dirIsNeg[2] ? b0 : b1)[2] - ray.o.z) * invDir.z;
        if ((tmin > tzmax) || (tzmin > tmax)) return false;
        if (tzmin > tmin) tmin = tzmin;
        if (tzmax < tmax) tmax = tzmax;
        return (tmin < ray.maxt) && (tmax > ray.mint);
    }

    // BVHAccel::Intersect, accelerators/bvh.cpp:403-454
    bool Intersect(const Ray &ray, Intersection *isect, Counters *cnt, float *b1 = nullptr, float *b2 = nullptr) const {
        if (cnt) cnt->closestRays++;
        return IntersectBVH(0, 0, ray, isect, cnt, b1, b2);
    }
    bool IntersectP(const Ray &ray, Counters *cnt) const {
        if (cnt) cnt->shadowRays++;
        return IntersectPBVH(0, 0, ray, cnt);
    }
    // nodeBase / primBase: where this BVHAccel's nodes[] and primitives[] start in the flattened arrays
    bool IntersectBVH(uint32_t nodeBase, uint32_t primBase, const Ray &ray, Intersection *isect, Counters *cnt,
                      float *b1, float *b2) const {
        bool hit = false;
        Vec invDir(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);
        uint32_t dirIsNeg[3] = {invDir.x < 0, invDir.y < 0, invDir.z < 0};
        uint32_t todoOffset = 0, nodeNum = 0;
        uint32_t todo[64];
        while (true) {
            const PbrtGpuBVHNode *node = &d->nodes[nodeBase + nodeNum];
            if (cnt) { cnt->nodesVisited++; cnt->closestNodes++; }
            if (BoxTest(*node, ray, invDir, dirIsNeg)) {
                if (node->nPrimitives > 0) {
                    for (uint32_t i = 0; i < node->nPrimitives; ++i) {
                        if (cnt) { cnt->primTests++; cnt->closestPrims++; }
                        if (PrimIntersect(primBase + node->offset + i, ray, isect, cnt, b1, b2)) hit = true;
                    }
                    if (todoOffset == 0) break;
                    nodeNum = todo[--todoOffset];
                } else {
                    if (dirIsNeg[node->axis]) {
                        todo[todoOffset++] = nodeNum + 1;
                        nodeNum = node->offset;
                    } else {
                        todo[todoOffset++] = node->offset;
                        nodeNum = nodeNum + 1;
                    }
                }
            } else {
                if (todoOffset == 0) break;
                nodeNum = todo[--todoOffset];
            }
        }
        return hit;
    }
    // BVHAccel::IntersectP, accelerators/bvh.cpp:457-503
    bool IntersectPBVH(uint32_t nodeBase, uint32_t primBase, const Ray &ray, Counters *cnt) const {
        Vec invDir(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);
        uint32_t dirIsNeg[3] = {invDir.x < 0, invDir.y < 0, invDir.z < 0};
        uint32_t todo[64];
        uint32_t todoOffset = 0, nodeNum = 0;
        while (true) {
            const PbrtGpuBVHNode *node = &d->nodes[nodeBase + nodeNum];
            if (cnt) cnt->nodesVisited++;
            if (BoxTest(*node, ray, invDir, dirIsNeg)) {
                if (node->nPrimitives > 0) {
                    for (uint32_t i = 0; i < node->nPrimitives; ++i) {
                        if (cnt) cnt->primTests++;
                        if (PrimIntersectP(primBase + node->offset + i, ray, cnt)) return true;
                    }
                    if (todoOffset == 0) break;
                    nodeNum = todo[--todoOffset];
                } else {
                    if (dirIsNeg[node->axis]) {
                        todo[todoOffset++] = nodeNum + 1;
                        nodeNum = node->offset;
                    } else {
                        todo[todoOffset++] = node->offset;
                        nodeNum = nodeNum + 1;
                    }
                }
            } else {
                if (todoOffset == 0) break;
                nodeNum = todo[--todoOffset];
            }
        }
        return false;
    }
};

// ====================================================================== BSDF
// core/reflection.h / reflection.cpp
enum {
    BSDF_REFLECTION = 1 << 0, BSDF_TRANSMISSION = 1 << 1, BSDF_DIFFUSE = 1 << 2, BSDF_GLOSSY = 1 << 3,
    BSDF_SPECULAR = 1 << 4,
    BSDF_ALL_TYPES = BSDF_DIFFUSE | BSDF_GLOSSY | BSDF_SPECULAR,
    BSDF_ALL_REFLECTION = BSDF_REFLECTION | BSDF_ALL_TYPES,
    BSDF_ALL_TRANSMISSION = BSDF_TRANSMISSION | BSDF_ALL_TYPES,
    BSDF_ALL = BSDF_ALL_REFLECTION | BSDF_ALL_TRANSMISSION
};
inline float CosTheta(const Vec &w) { return w.z; }
inline float AbsCosTheta(const Vec &w) { return fabsf(w.z); }
inline float SinTheta2(const Vec &w) { return std::max(0.f, 1.f - CosTheta(w) * CosTheta(w)); }
inline float SinTheta(const Vec &w) { return sqrtf(SinTheta2(w)); }
inline bool SameHemisphere(const Vec &w, const Vec &wp) { return w.z * wp.z > 0.f; }
inline float SphericalTheta(const Vec &v) { return acosf(Clampf(v.z, -1.f, 1.f)); }   // geometry.h:640-642
inline float SphericalPhi(const Vec &v) {   // geometry.h:644-647
    float p = atan2f(v.y, v.x);
    return (p < 0.f) ? p + 2.f * PI_F : p;
}
inline Vec SphericalDirection(float sintheta, float costheta, float phi) {
    return Vec(sintheta * cosf(phi), sintheta * sinf(phi), costheta);
}

void ConcentricSampleDisk(float u1, float u2, float *dx, float *dy) {   // core/montecarlo.cpp:306-348
    float r, theta;
    float sx = 2 * u1 - 1;
    float sy = 2 * u2 - 1;
    if (sx == 0.0 && sy == 0.0) { *dx = 0.0; *dy = 0.0; return; }
    if (sx >= -sy) {
        if (sx > sy) {
            r = sx;
            if (sy > 0.0) theta = sy / r;
            else theta = 8.0f + sy / r;
        } else {
            r = sy;
            theta = 2.0f - sx / r;
        }
    } else {
        if (sx <= sy) {
            r = -sx;
            theta = 4.0f - sy / r;
        } else {
            r = -sy;
            theta = 6.0f + sx / r;
        }
    }
    theta *= PI_F / 4.f;
    *dx = r * cosf(theta);
    *dy = r * sinf(theta);
}
inline Vec CosineSampleHemisphere(float u1, float u2) {   // montecarlo.h:128-133
    Vec ret;
    ConcentricSampleDisk(u1, u2, &ret.x, &ret.y);
    ret.z = sqrtf(std::max(0.f, 1.f - ret.x * ret.x - ret.y * ret.y));
    return ret;
}
inline Vec UniformSampleSphere(float u1, float u2) {   // montecarlo.cpp:283-290
    float z = 1.f - 2.f * u1;
    float r = sqrtf(std::max(0.f, 1.f - z * z));
    float phi = 2.f * PI_F * u2;
    return Vec(r * cosf(phi), r * sinf(phi), z);
}
inline float UniformConePdf(float cosThetaMax) { return 1.f / (2.f * PI_F * (1.f - cosThetaMax)); }
inline Vec UniformSampleCone(float u1, float u2, float costhetamax, const Vec &x, const Vec &y, const Vec &z) {
    // montecarlo.cpp:423-430
    float costheta = Lerp(u1, costhetamax, 1.f);
    float sintheta = sqrtf(1.f - costheta * costheta);
    float phi = u2 * 2.f * PI_F;
    return cosf(phi) * sintheta * x + sinf(phi) * sintheta * y + costheta * z;
}
inline float PowerHeuristic(int nf, float fPdf, int ng, float gPdf) {   // montecarlo.h:266-269
    float f = nf * fPdf, g = ng * gPdf;
    return (f * f) / (f * f + g * g);
}

enum BxDFKind { BX_LAMBERTIAN, BX_MICROFACET_BLINN, BX_IRREG_ISO,
                BX_SPECULAR_REFLECTION,     // exponent = eta_t of its FresnelDielectric(1, eta_t), or 0 for FresnelNoOp
                BX_SPECULAR_TRANSMISSION }; // exponent = eta_t (etai = 1)
struct BxDF {
    BxDFKind kind;
    int type;         // BxDFType flags
    Spec R;
    float exponent;   // Blinn
    int measured;     // IrregIsotropicBRDF kd-tree
    bool conductor = false;        // Microfacet's Fresnel term: FresnelConductor(eta, k) (metal) instead of FresnelDielectric(1.5, 1)
    Spec eta = Spec(0.f), k = Spec(0.f);
};

struct BSDFSample { float uDir[2], uComponent; };
struct LightSample { float uPos[2], uComponent; };

struct KdLookup {   // IrregIsoProc, reflection.cpp:40-55
    Spec v;
    float sumWeights = 0.f;
    int nFound = 0;
};

struct BSDF {
    const SceneView *sv;
    Vec p;               // dgShading.p
    Vec nn, ng, sn, tn;  // shading normal, geometric normal, tangents
    int nBxDFs;
    BxDF bxdfs[2];

    Vec WorldToLocal(const Vec &v) const { return Vec(Dot(v, sn), Dot(v, tn), Dot(v, nn)); }
    Vec LocalToWorld(const Vec &v) const {
        return Vec(sn.x * v.x + tn.x * v.y + nn.x * v.z, sn.y * v.x + tn.y * v.y + nn.y * v.z,
                   sn.z * v.x + tn.z * v.y + nn.z * v.z);
    }
    static bool Matches(const BxDF &b, int flags) { return (b.type & flags) == b.type; }
    int NumComponents(int flags) const {
        int n = 0;
        for (int i = 0; i < nBxDFs; ++i) if (Matches(bxdfs[i], flags)) ++n;
        return n;
    }

    // KdTree::privateLookup, core/kdtree.h:157-183 (recursive, children before node)
    void KdLookupRec(const PbrtGpuMeasured &m, uint32_t nodeNum, const Vec &p, KdLookup &proc, float &maxDistSquared) const {
        const PbrtGpuKdNode &node = sv->d->kdNodes[m.firstNode + nodeNum];
        int axis = node.bits & 3u;
        uint32_t hasLeft = (node.bits >> 2) & 1u, rightChild = node.bits >> 3;
        if (axis != 3) {
            float dist2 = (p[axis] - node.splitPos) * (p[axis] - node.splitPos);
            if (p[axis] <= node.splitPos) {
                if (hasLeft) KdLookupRec(m, nodeNum + 1, p, proc, maxDistSquared);
                if (dist2 < maxDistSquared && rightChild < m.nNodes) KdLookupRec(m, rightChild, p, proc, maxDistSquared);
            } else {
                if (rightChild < m.nNodes) KdLookupRec(m, rightChild, p, proc, maxDistSquared);
                if (dist2 < maxDistSquared && hasLeft) KdLookupRec(m, nodeNum + 1, p, proc, maxDistSquared);
            }
        }
        const PbrtGpuBrdfSample &s = sv->d->kdData[m.firstNode + nodeNum];
        float dist2 = DistanceSquared(Vec(s.p), p);
        if (dist2 < maxDistSquared) {   // IrregIsoProc::operator()
            float weight = expf(-100.f * dist2);
            proc.v += weight * Spec(s.v);
            proc.sumWeights += weight;
            ++proc.nFound;
        }
    }

    // ---- individual BxDFs ----
    static Spec FrDiel(float cosi, float cost, const Spec &etai, const Spec &etat) {   // reflection.cpp:60-67
        Spec Rparl = ((etat * cosi) - (etai * cost)) / ((etat * cosi) + (etai * cost));
        Spec Rperp = ((etai * cosi) - (etat * cost)) / ((etai * cosi) + (etat * cost));
        return (Rparl * Rparl + Rperp * Rperp) / 2.f;
    }
    static Spec FrCond(float cosi, const Spec &eta, const Spec &k) {   // reflection.cpp:70-79
        Spec tmp = (eta * eta + k * k) * cosi * cosi;
        Spec Rparl2 = (tmp - (2.f * eta * cosi) + Spec(1.f)) / (tmp + (2.f * eta * cosi) + Spec(1.f));
        Spec tmp_f = eta * eta + k * k;
        Spec Rperp2 = (tmp_f - (2.f * eta * cosi) + Spec(cosi * cosi)) / (tmp_f + (2.f * eta * cosi) + Spec(cosi * cosi));
        return (Rparl2 + Rperp2) / 2.f;
    }
    static Spec FresnelDielectricEvaluate(float cosi, float eta_i, float eta_t) {   // reflection.cpp:115-135
        cosi = Clampf(cosi, -1.f, 1.f);
        bool entering = cosi > 0.;
        float ei = eta_i, et = eta_t;
        if (!entering) std::swap(ei, et);
        float sint = ei / et * sqrtf(std::max(0.f, 1.f - cosi * cosi));
        if (sint >= 1.) return Spec(1.);
        float cost = sqrtf(std::max(0.f, 1.f - sint * sint));
        return FrDiel(fabsf(cosi), cost, Spec(ei), Spec(et));
    }
    static float BlinnD(float exponent, const Vec &wh) {   // reflection.h:427-430
        float costhetah = AbsCosTheta(wh);
        return (exponent + 2) * INV_TWOPI_F * powf(costhetah, exponent);
    }
    static float MicrofacetG(const Vec &wo, const Vec &wi, const Vec &wh) {   // reflection.h:403-410
        float NdotWh = AbsCosTheta(wh);
        float NdotWo = AbsCosTheta(wo);
        float NdotWi = AbsCosTheta(wi);
        float WOdotWh = AbsDot(wo, wh);
        return std::min(1.f, std::min((2.f * NdotWh * NdotWo / WOdotWh), (2.f * NdotWh * NdotWi / WOdotWh)));
    }
    static Vec BRDFRemap(const Vec &wo, const Vec &wi) {   // reflection.cpp:247-256
        float cosi = CosTheta(wi), coso = CosTheta(wo);
        float sini = SinTheta(wi), sino = SinTheta(wo);
        float phii = SphericalPhi(wi), phio = SphericalPhi(wo);
        float dphi = phii - phio;
        if (dphi < 0.) dphi += 2.f * PI_F;
        if (dphi > 2.f * PI_F) dphi -= 2.f * PI_F;
        if (dphi > PI_F) dphi = 2.f * PI_F - dphi;
        return Vec(sini * sino, dphi / PI_F, cosi * coso);
    }

    Spec BxF(const BxDF &b, const Vec &wo, const Vec &wi) const {
        switch (b.kind) {
        case BX_SPECULAR_REFLECTION:    // reflection.h:317-319
        case BX_SPECULAR_TRANSMISSION:  // reflection.h:339-341
            return Spec(0.f);
        case BX_LAMBERTIAN: return b.R * INV_PI_F;   // reflection.cpp:173-175
        case BX_MICROFACET_BLINN: {                  // Microfacet::f, reflection.cpp:211-222
            float cosThetaO = AbsCosTheta(wo);
            float cosThetaI = AbsCosTheta(wi);
            if (cosThetaI == 0.f || cosThetaO == 0.f) return Spec(0.f);
            Vec wh = wi + wo;
            if (wh.x == 0. && wh.y == 0. && wh.z == 0.) return Spec(0.f);
            wh = Normalize(wh);
            float cosThetaH = Dot(wi, wh);
            // plastic.cpp:61: FresnelDielectric(1.5, 1); metal.cpp:63-65: FresnelConductor(eta, k), Evaluate = FrCond(|cosi|, ..)
            Spec F = b.conductor ? FrCond(fabsf(cosThetaH), b.eta, b.k) : FresnelDielectricEvaluate(cosThetaH, 1.5f, 1.f);
            return b.R * BlinnD(b.exponent, wh) * MicrofacetG(wo, wi, wh) * F / (4.f * cosThetaI * cosThetaO);
        }
        case BX_IRREG_ISO: {                         // IrregIsotropicBRDF::f, reflection.cpp:259-272
            Vec m = BRDFRemap(wo, wi);
            float lastMaxDist2 = .001f;
            const PbrtGpuMeasured &md = sv->d->measured[b.measured];
            while (true) {
                KdLookup proc;
                float maxDist2 = lastMaxDist2;
                KdLookupRec(md, 0, m, proc, maxDist2);
                if (proc.nFound > 2 || lastMaxDist2 > 1.5f) return proc.v.Clamp() / proc.sumWeights;
                lastMaxDist2 *= 2.f;
            }
        }
        }
        return Spec(0.f);
    }
    float BxPdf(const BxDF &b, const Vec &wo, const Vec &wi) const {
        if (b.kind == BX_SPECULAR_REFLECTION || b.kind == BX_SPECULAR_TRANSMISSION) return 0.f;   // reflection.h:322-324,344-346
        if (b.kind == BX_MICROFACET_BLINN) {   // Microfacet::Pdf + Blinn::Pdf, reflection.cpp:340-343,365-374
            if (!SameHemisphere(wo, wi)) return 0.f;
            Vec wh = Normalize(wo + wi);
            float costheta = AbsCosTheta(wh);
            float blinn_pdf = ((b.exponent + 1.f) * powf(costheta, b.exponent)) / (2.f * PI_F * 4.f * Dot(wo, wh));
            if (Dot(wo, wh) <= 0.f) blinn_pdf = 0.f;
            return blinn_pdf;
        }
        // BxDF::Pdf, reflection.cpp:321-323
        return SameHemisphere(wo, wi) ? AbsCosTheta(wi) * INV_PI_F : 0.f;
    }
    Spec BxSample_f(const BxDF &b, const Vec &wo, Vec *wi, float u1, float u2, float *pdf) const {
        if (b.kind == BX_SPECULAR_REFLECTION) {   // SpecularReflection::Sample_f, reflection.cpp:138-144
            *wi = Vec(-wo.x, -wo.y, wo.z);
            *pdf = 1.f;
            Spec F = b.exponent == 0.f ? Spec(1.f) : FresnelDielectricEvaluate(CosTheta(wo), 1.f, b.exponent);
            return F * b.R / AbsCosTheta(*wi);
        }
        if (b.kind == BX_SPECULAR_TRANSMISSION) {   // SpecularTransmission::Sample_f, reflection.cpp:147-170
            bool entering = CosTheta(wo) > 0.;
            float ei = 1.f, et = b.exponent;
            if (!entering) std::swap(ei, et);
            float sini2 = SinTheta2(wo);
            float eta = ei / et;
            float sint2 = eta * eta * sini2;
            if (sint2 >= 1.) return Spec(0.f);   // total internal reflection (*pdf stays as the caller set it: 0)
            float cost = sqrtf(std::max(0.f, 1.f - sint2));
            if (entering) cost = -cost;
            float sintOverSini = eta;
            *wi = Vec(sintOverSini * -wo.x, sintOverSini * -wo.y, cost);
            *pdf = 1.f;
            Spec F = FresnelDielectricEvaluate(CosTheta(wo), 1.f, b.exponent);
            return (Spec(1.f) - F) * b.R / AbsCosTheta(*wi);
        }
        if (b.kind == BX_MICROFACET_BLINN) {   // Microfacet::Sample_f + Blinn::Sample_f, reflection.cpp:332-363
            float costheta = powf(u1, 1.f / (b.exponent + 1));
            float sintheta = sqrtf(std::max(0.f, 1.f - costheta * costheta));
            float phi = u2 * 2.f * PI_F;
            Vec wh = SphericalDirection(sintheta, costheta, phi);
            if (!SameHemisphere(wo, wh)) wh = -wh;
            *wi = -wo + 2.f * Dot(wo, wh) * wh;
            float blinn_pdf = ((b.exponent + 1.f) * powf(costheta, b.exponent)) / (2.f * PI_F * 4.f * Dot(wo, wh));
            if (Dot(wo, wh) <= 0.f) blinn_pdf = 0.f;
            *pdf = blinn_pdf;
            if (!SameHemisphere(wo, *wi)) return Spec(0.f);
            return BxF(b, wo, *wi);
        }
        // BxDF::Sample_f (cosine hemisphere), reflection.cpp:311-318
        *wi = CosineSampleHemisphere(u1, u2);
        if (wo.z < 0.) wi->z *= -1.f;
        *pdf = BxPdf(b, wo, *wi);
        return BxF(b, wo, *wi);
    }

    // BSDF::f, reflection.cpp:612-626
    Spec f(const Vec &woW, const Vec &wiW, int flags) const {
        Vec wi = WorldToLocal(wiW), wo = WorldToLocal(woW);
        if (Dot(wiW, ng) * Dot(woW, ng) > 0) flags &= ~BSDF_TRANSMISSION;
        else flags &= ~BSDF_REFLECTION;
        Spec f(0.);
        for (int i = 0; i < nBxDFs; ++i)
            if (Matches(bxdfs[i], flags)) f += BxF(bxdfs[i], wo, wi);
        return f;
    }
    // BSDF::Pdf, reflection.cpp:583-598
    float Pdf(const Vec &woW, const Vec &wiW, int flags) const {
        if (nBxDFs == 0.) return 0.;
        Vec wo = WorldToLocal(woW), wi = WorldToLocal(wiW);
        float pdf = 0.f;
        int matchingComps = 0;
        for (int i = 0; i < nBxDFs; ++i)
            if (Matches(bxdfs[i], flags)) {
                ++matchingComps;
                pdf += BxPdf(bxdfs[i], wo, wi);
            }
        return matchingComps > 0 ? pdf / matchingComps : 0.f;
    }
    // BSDF::Sample_f, reflection.cpp:522-580
    Spec Sample_f(const Vec &woW, Vec *wiW, const BSDFSample &bsdfSample, float *pdf, int flags, int *sampledType) const {
        int matchingComps = NumComponents(flags);
        if (matchingComps == 0) {
            *pdf = 0.f;
            if (sampledType) *sampledType = 0;
            return Spec(0.f);
        }
        int which = std::min(Floor2Int(bsdfSample.uComponent * matchingComps), matchingComps - 1);
        const BxDF *bxdf = nullptr;
        int count = which;
        for (int i = 0; i < nBxDFs; ++i)
            if (Matches(bxdfs[i], flags) && count-- == 0) { bxdf = &bxdfs[i]; break; }
        Vec wo = WorldToLocal(woW);
        Vec wi;
        *pdf = 0.f;
        Spec f = BxSample_f(*bxdf, wo, &wi, bsdfSample.uDir[0], bsdfSample.uDir[1], pdf);
        if (*pdf == 0.f) {
            if (sampledType) *sampledType = 0;
            return Spec(0.f);
        }
        if (sampledType) *sampledType = bxdf->type;
        *wiW = LocalToWorld(wi);
        if (!(bxdf->type & BSDF_SPECULAR) && matchingComps > 1)
            for (int i = 0; i < nBxDFs; ++i)
                if (&bxdfs[i] != bxdf && Matches(bxdfs[i], flags)) *pdf += BxPdf(bxdfs[i], wo, wi);
        if (matchingComps > 1) *pdf /= matchingComps;
        if (!(bxdf->type & BSDF_SPECULAR)) {
            f = Spec(0.);
            if (Dot(*wiW, ng) * Dot(woW, ng) > 0) flags &= ~BSDF_TRANSMISSION;
            else flags &= ~BSDF_REFLECTION;
            for (int i = 0; i < nBxDFs; ++i)
                if (Matches(bxdfs[i], flags)) f += BxF(bxdfs[i], wo, wi);
        }
        return f;
    }
};

// ================================================================ the renderer
struct Renderer {
    SceneView sv;
    const PbrtGpuRenderDesc *rd;
    int xPixelStart, yPixelStart, xPixelCount, yPixelCount;   // film extent, image.cpp:48-51
    int sxStart, sxEnd, syStart, syEnd;                       // sampler extent, image.cpp:157-166
    PbrtGpuPixel *pixels;

    void Init(const PbrtGpuSceneDesc *d, const PbrtGpuRenderDesc *r, PbrtGpuPixel *out) {
        sv.d = d; rd = r; pixels = out;
        xPixelStart = Ceil2Int(rd->xResolution * rd->cropWindow[0]);
        xPixelCount = std::max(1, Ceil2Int(rd->xResolution * rd->cropWindow[1]) - xPixelStart);
        yPixelStart = Ceil2Int(rd->yResolution * rd->cropWindow[2]);
        yPixelCount = std::max(1, Ceil2Int(rd->yResolution * rd->cropWindow[3]) - yPixelStart);
        sxStart = Floor2Int(xPixelStart + 0.5f - rd->filterXWidth);
        sxEnd = Ceil2Int(xPixelStart - 0.5f + xPixelCount + rd->filterXWidth);
        syStart = Floor2Int(yPixelStart + 0.5f - rd->filterYWidth);
        syEnd = Ceil2Int(yPixelStart - 0.5f + yPixelCount + rd->filterYWidth);
    }

    // Sampler::ComputeSubWindow, core/sampler.cpp:55-74
    void ComputeSubWindow(int num, int count, int *x0, int *x1, int *y0, int *y1) const {
        int dx = sxEnd - sxStart, dy = syEnd - syStart;
        int nx = count, ny = 1;
        while ((nx & 0x1) == 0 && 2 * dx * ny < dy * nx) { nx >>= 1; ny <<= 1; }
        int xo = num % nx, yo = num / nx;
        float tx0 = float(xo) / float(nx), tx1 = float(xo + 1) / float(nx);
        float ty0 = float(yo) / float(ny), ty1 = float(yo + 1) / float(ny);
        *x0 = Floor2Int(Lerp(tx0, sxStart, sxEnd));
        *x1 = Floor2Int(Lerp(tx1, sxStart, sxEnd));
        *y0 = Floor2Int(Lerp(ty0, syStart, syEnd));
        *y1 = Floor2Int(Lerp(ty1, syStart, syEnd));
    }

    // PerspectiveCamera::GenerateRayDifferential, cameras/perspective.cpp:81-138 (main ray only:
    // the auxiliary rx/ry rays feed texture filtering, which no in-scope material uses)
    void GenerateRay(const Sample &sample, Ray *ray) const {
        Vec Pras(sample.imageX, sample.imageY, 0);
        Vec Pcamera = XfPoint(rd->rasterToCamera, Pras);
        Vec dir = Normalize(Vec(Pcamera.x, Pcamera.y, Pcamera.z));
        *ray = Ray(Vec(0, 0, 0), dir, 0.f, INF_F);
        if (rd->lensRadius > 0.) {
            float lensU, lensV;
            ConcentricSampleDisk(sample.lensU, sample.lensV, &lensU, &lensV);
            lensU *= rd->lensRadius;
            lensV *= rd->lensRadius;
            float ft = rd->focalDistance / ray->d.z;
            Vec Pfocus = (*ray)(ft);
            ray->o = Vec(lensU, lensV, 0.f);
            ray->d = Normalize(Pfocus - ray->o);
        }
        ray->time = sample.time;
        ray->o = XfPoint(rd->cameraToWorld, ray->o);   // CameraToWorld(*ray, ray), transform.h:253-262
        ray->d = XfVector(rd->cameraToWorld, ray->d);
    }

    // Triangle::GetShadingGeometry, shapes/trianglemesh.cpp:293-368; Shape::GetShadingGeometry shape.h:58-62
    void ShadingGeometry(const Intersection &isect, DiffGeom *dgs, Counters *cnt) const {
        const PbrtGpuPrimitive &p = sv.d->primitives[isect.prim];
        *dgs = isect.dg;
        if (p.type != PBRTGPU_PRIM_TRIANGLE) return;
        const PbrtGpuTriangle &tri = sv.d->triangles[p.index];
        const PbrtGpuMesh &mesh = sv.d->meshes[tri.mesh];
        if (!(mesh.flags & PBRTGPU_MESH_HAS_N)) return;
        if (cnt) cnt->shadedN++;
        const DiffGeom &dg = isect.dg;
        float b[3];
        float uv[3][2];
        sv.GetUVs(tri, uv);
        float A[2][2] = {{uv[1][0] - uv[0][0], uv[2][0] - uv[0][0]}, {uv[1][1] - uv[0][1], uv[2][1] - uv[0][1]}};
        float C[2] = {dg.u - uv[0][0], dg.v - uv[0][1]};
        {   // SolveLinearSystem2x2, core/transform.cpp:40-50
            float det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
            bool ok = !(fabsf(det) < 1e-10f);
            if (ok) {
                b[1] = (A[1][1] * C[0] - A[0][1] * C[1]) / det;
                b[2] = (A[0][0] * C[1] - A[1][0] * C[0]) / det;
                if (std::isnan(b[1]) || std::isnan(b[2])) ok = false;
            }
            if (!ok) b[0] = b[1] = b[2] = 1.f / 3.f;
            else b[0] = 1.f - b[1] - b[2];
        }
        // obj2world = isect.ObjectToWorld; under a TransformedPrimitive that is Inverse(WorldToObject * w2p)
        // (primitive.cpp:106-107), so its mInv -- what a Normal is transformed with -- is WorldToObject.m * w2p.m
        PbrtGpuXform o2w = sv.d->xforms[mesh.xform];
        if (isect.instanced) {
            float wm[16];
            MatMul(o2w.mInv, isect.w2p.m, wm);
            memcpy(o2w.mInv, wm, sizeof(wm));
        }
        Vec n0(sv.d->N + 3 * tri.v[0]), n1(sv.d->N + 3 * tri.v[1]), n2(sv.d->N + 3 * tri.v[2]);
        Vec ns = Normalize(XfNormal(o2w.mInv, b[0] * n0 + b[1] * n1 + b[2] * n2));
        Vec ss = Normalize(dg.dpdu);
        Vec ts = Cross(ss, ns);
        if (LengthSquared(ts) > 0.f) {
            ts = Normalize(ts);
            ss = Cross(ts, ns);
        } else
            CoordinateSystem(ns, &ss, &ts);
        // dndu/dndv are texture-only: omitted.  DifferentialGeometry ctor recomputes nn from (ss, ts).
        InitDG(dgs, dg.p, ss, ts, dg.u, dg.v, mesh.flags);
    }

    // Material::GetBSDF: matte.cpp:42-63, plastic.cpp:42-66, measured.cpp:194-210; BSDF ctor reflection.cpp:601-609
    void GetBSDF(const Intersection &isect, BSDF *bsdf, Counters *cnt) const {
        DiffGeom dgs;
        ShadingGeometry(isect, &dgs, cnt);
        bsdf->sv = &sv;
        bsdf->p = dgs.p;
        bsdf->ng = isect.dg.nn;
        bsdf->nn = dgs.nn;
        bsdf->sn = Normalize(dgs.dpdu);
        bsdf->tn = Cross(bsdf->nn, bsdf->sn);
        bsdf->nBxDFs = 0;
        const PbrtGpuMaterial &m = sv.d->materials[sv.d->primitives[isect.prim].material];
        if (m.type == PBRTGPU_MAT_MATTE) {
            Spec r = Spec(m.Kd).Clamp();
            if (!r.IsBlack()) bsdf->bxdfs[bsdf->nBxDFs++] = BxDF{BX_LAMBERTIAN, BSDF_REFLECTION | BSDF_DIFFUSE, r, 0.f, -1};
        } else if (m.type == PBRTGPU_MAT_PLASTIC) {
            Spec kd = Spec(m.Kd).Clamp();
            if (!kd.IsBlack()) bsdf->bxdfs[bsdf->nBxDFs++] = BxDF{BX_LAMBERTIAN, BSDF_REFLECTION | BSDF_DIFFUSE, kd, 0.f, -1};
            Spec ks = Spec(m.Ks).Clamp();
            if (!ks.IsBlack()) {
                float e = 1.f / m.roughness;
                if (e > 10000.f || std::isnan(e)) e = 10000.f;   // Blinn ctor, reflection.h:424-425
                bsdf->bxdfs[bsdf->nBxDFs++] = BxDF{BX_MICROFACET_BLINN, BSDF_REFLECTION | BSDF_GLOSSY, ks, e, -1};
            }
        } else if (m.type == PBRTGPU_MAT_METAL) {    // metal.cpp:51-67: Microfacet(1., FresnelConductor(eta, k), Blinn(1 / rough))
            float e = 1.f / m.roughness;
            if (e > 10000.f || std::isnan(e)) e = 10000.f;   // Blinn ctor, reflection.h:424-425
            BxDF b{BX_MICROFACET_BLINN, BSDF_REFLECTION | BSDF_GLOSSY, Spec(1.f), e, -1};
            b.conductor = true; b.eta = Spec(m.Kd); b.k = Spec(m.Ks);
            bsdf->bxdfs[bsdf->nBxDFs++] = b;
        } else if (m.type == PBRTGPU_MAT_MIRROR) {   // mirror.cpp:41-53: SpecularReflection(R, FresnelNoOp)
            Spec r = Spec(m.Ks).Clamp();
            if (!r.IsBlack()) bsdf->bxdfs[bsdf->nBxDFs++] = BxDF{BX_SPECULAR_REFLECTION, BSDF_REFLECTION | BSDF_SPECULAR, r, 0.f, -1};
        } else if (m.type == PBRTGPU_MAT_GLASS) {    // glass.cpp:41-57
            float ior = m.roughness;
            Spec r = Spec(m.Ks).Clamp(), t = Spec(m.Kd).Clamp();
            if (!r.IsBlack()) bsdf->bxdfs[bsdf->nBxDFs++] = BxDF{BX_SPECULAR_REFLECTION, BSDF_REFLECTION | BSDF_SPECULAR, r, ior, -1};
            if (!t.IsBlack()) bsdf->bxdfs[bsdf->nBxDFs++] = BxDF{BX_SPECULAR_TRANSMISSION, BSDF_TRANSMISSION | BSDF_SPECULAR, t, ior, -1};
        } else {
            bsdf->bxdfs[bsdf->nBxDFs++] = BxDF{BX_IRREG_ISO, BSDF_REFLECTION | BSDF_GLOSSY, Spec(0.f), 0.f, m.measured};
        }
    }

    // Intersection::Le, core/intersection.cpp:61-64 + DiffuseAreaLight::L, lights/diffuse.h:51-53
    Spec Le(const Intersection &isect, const Vec &w) const {
        int al = sv.d->primitives[isect.prim].areaLight;
        if (al < 0) return Spec(0.);
        return Dot(isect.dg.nn, w) > 0.f ? Spec(sv.d->lights[al].L) : Spec(0.f);
    }

    // ---- shape sampling for area lights ----
    // Shape::Pdf(p, wi), core/shape.cpp:86-99
    float ShapePdf(const PbrtGpuQuadric &q, const Vec &p, const Vec &wi) const {
        DiffGeom dgLight;
        Ray ray(p, wi, 1e-3f);
        float thit, rayEpsilon;
        if (!sv.QuadricIntersect(q, ray, &thit, &rayEpsilon, &dgLight)) return 0.;
        float pdf = DistanceSquared(p, ray(thit)) / (AbsDot(dgLight.nn, -wi) * q.area);
        if (std::isinf(pdf)) pdf = 0.f;
        return pdf;
    }
    // Sphere::Sample(u1,u2), sphere.cpp:226-231
    Vec SphereSampleUniform(const PbrtGpuQuadric &q, float u1, float u2, Vec *ns) const {
        const PbrtGpuXform &x = sv.d->xforms[q.xform];
        Vec p = Vec(0, 0, 0) + q.radius * UniformSampleSphere(u1, u2);
        *ns = Normalize(XfNormal(x.mInv, Vec(p.x, p.y, p.z)));
        if (q.flags & PBRTGPU_SHAPE_REVERSE_ORIENT) *ns = *ns * -1.f;
        return XfPoint(x.m, p);
    }
    // Sphere::Sample(p,u1,u2), sphere.cpp:236-261
    Vec SphereSample(const PbrtGpuQuadric &q, const Vec &p, float u1, float u2, Vec *ns) const {
        const PbrtGpuXform &x = sv.d->xforms[q.xform];
        Vec Pcenter = XfPoint(x.m, Vec(0, 0, 0));
        Vec wc = Normalize(Pcenter - p);
        Vec wcX, wcY;
        CoordinateSystem(wc, &wcX, &wcY);
        if (DistanceSquared(p, Pcenter) - q.radius * q.radius < 1e-4f) return SphereSampleUniform(q, u1, u2, ns);
        float sinThetaMax2 = q.radius * q.radius / DistanceSquared(p, Pcenter);
        float cosThetaMax = sqrtf(std::max(0.f, 1.f - sinThetaMax2));
        DiffGeom dgSphere;
        float thit, rayEpsilon;
        Vec ps;
        Ray r(p, UniformSampleCone(u1, u2, cosThetaMax, wcX, wcY, wc), 1e-3f);
        if (!sv.SphereIntersect(q, r, &thit, &rayEpsilon, &dgSphere)) thit = Dot(Pcenter - p, Normalize(r.d));
        ps = r(thit);
        *ns = Normalize(ps - Pcenter);
        if (q.flags & PBRTGPU_SHAPE_REVERSE_ORIENT) *ns = *ns * -1.f;
        return ps;
    }
    // Sphere::Pdf(p,wi), sphere.cpp:264-274
    float SpherePdf(const PbrtGpuQuadric &q, const Vec &p, const Vec &wi) const {
        const PbrtGpuXform &x = sv.d->xforms[q.xform];
        Vec Pcenter = XfPoint(x.m, Vec(0, 0, 0));
        if (DistanceSquared(p, Pcenter) - q.radius * q.radius < 1e-4f) return ShapePdf(q, p, wi);
        float sinThetaMax2 = q.radius * q.radius / DistanceSquared(p, Pcenter);
        float cosThetaMax = sqrtf(std::max(0.f, 1.f - sinThetaMax2));
        return UniformConePdf(cosThetaMax);
    }
    // Disk::Sample, disk.cpp:148-156 (Shape::Sample(p,...) forwards to it, shape.h:77-80)
    Vec DiskSample(const PbrtGpuQuadric &q, float u1, float u2, Vec *Ns) const {
        const PbrtGpuXform &x = sv.d->xforms[q.xform];
        Vec p;
        ConcentricSampleDisk(u1, u2, &p.x, &p.y);
        p.x *= q.radius;
        p.y *= q.radius;
        p.z = q.height;
        *Ns = Normalize(XfNormal(x.mInv, Vec(0, 0, 1)));
        if (q.flags & PBRTGPU_SHAPE_REVERSE_ORIENT) *Ns = *Ns * -1.f;
        return XfPoint(x.m, p);
    }

    // ---- InfiniteAreaLight (lights/infinite.cpp): radiance map level 0 + Distribution2D arrive ready-made in envMaps[] ----
    // (what the constructor computes, infinite.cpp:67-104, is host front-end work: pbrt_v2_b200/csrc/host/envmap.cpp)
    // MIPMap::Lookup(s, t, width = 0) (core/mipmap.h:239-255): level = nLevels - 1 + Log2(1e-8) < 0 -> triangle(0, s, t)
    // (:259-269): bilinear over level 0, TEXTURE_REPEAT wraps every index (Texel, :203-228; Mod, pbrt.h:237-242)
    static int Mod(int a, int b) {
        int n = int(a / b);
        a -= n * b;
        if (a < 0) a += b;
        return a;
    }
    Spec EnvTexel(const PbrtGpuEnvMap &m, int s, int t) const {
        s = Mod(s, (int)m.width);
        t = Mod(t, (int)m.height);
        return Spec(sv.d->envData + m.texels + 3 * ((size_t)t * m.width + s));
    }
    Spec InfiniteLookup(const PbrtGpuLight &l, float s, float t) const {
        const PbrtGpuEnvMap &m = sv.d->envMaps[l.nShapes];
        s = s * m.width - 0.5f;
        t = t * m.height - 0.5f;
        int s0 = Floor2Int(s), t0 = Floor2Int(t);
        float ds = s - s0, dt = t - t0;
        return (1.f - ds) * (1.f - dt) * EnvTexel(m, s0, t0) + (1.f - ds) * dt * EnvTexel(m, s0, t0 + 1) +
               ds * (1.f - dt) * EnvTexel(m, s0 + 1, t0) + ds * dt * EnvTexel(m, s0 + 1, t0 + 1);
    }
    // Light::Le(ray) (core/light.cpp:59-61: black) / InfiniteAreaLight::Le (infinite.cpp:114-119)
    Spec LightLe(const PbrtGpuLight &l, const Ray &r) const {
        if (l.type != PBRTGPU_LIGHT_INFINITE) return Spec(0.f);
        Vec wh = Normalize(XfVector(sv.d->xforms[l.firstShape].mInv, r.d));   // WorldToLight(r.d)
        float s = SphericalPhi(wh) * INV_TWOPI_F;
        float t = SphericalTheta(wh) * INV_PI_F;
        return InfiniteLookup(l, s, t);
    }
    Spec EscapedLe(const Ray &r) const {   // "for every light: L += light->Le(ray)" (samplerrenderer.cpp:239-242, path.cpp:114-116)
        Spec L(0.f);
        for (uint32_t i = 0; i < sv.d->nLights; ++i) L += LightLe(sv.d->lights[i], r);
        return L;
    }
    // Distribution1D::SampleContinuous(u, &pdf, &off), montecarlo.h:81-97 (std::upper_bound: first cdf entry > u)
    static float SampleContinuous1D(const float *func, const float *cdf, float funcInt, int count, float u, float *pdf, int *off) {
        const float *ptr = std::upper_bound(cdf, cdf + count + 1, u);
        int offset = std::max(0, int(ptr - cdf - 1));
        if (off) *off = offset;
        float du = (u - cdf[offset]) / (cdf[offset + 1] - cdf[offset]);
        if (pdf) *pdf = func[offset] / funcInt;
        return (offset + du) / count;
    }
    // Distribution2D::SampleContinuous(u0, u1, uv, &pdf), montecarlo.h:146-153
    void InfiniteSampleContinuous(const PbrtGpuLight &l, float u0, float u1, float uv[2], float *pdf) const {
        const PbrtGpuEnvMap &m = sv.d->envMaps[l.nShapes];
        const float *e = sv.d->envData;
        float pdfs[2];
        int v;
        uv[1] = SampleContinuous1D(e + m.funcInt, e + m.margCdf, m.margFuncInt, (int)m.nv, u1, &pdfs[1], &v);
        uv[0] = SampleContinuous1D(e + m.func + (size_t)v * m.nu, e + m.cdf + (size_t)v * (m.nu + 1), e[m.funcInt + v], (int)m.nu, u0,
                                   &pdfs[0], nullptr);
        *pdf = pdfs[0] * pdfs[1];
    }
    // Distribution2D::Pdf(u, v), montecarlo.h:154-162 (Float2Int truncates, pbrt.h:296-298)
    float InfiniteMapPdf(const PbrtGpuLight &l, float u, float v) const {
        const PbrtGpuEnvMap &m = sv.d->envMaps[l.nShapes];
        const float *e = sv.d->envData;
        int iu = std::min(std::max((int)(u * (int)m.nu), 0), (int)m.nu - 1);
        int iv = std::min(std::max((int)(v * (int)m.nv), 0), (int)m.nv - 1);
        if (e[m.funcInt + iv] * m.margFuncInt == 0.f) return 0.f;
        return (e[m.func + (size_t)iv * m.nu + iu] * e[m.funcInt + iv]) / (e[m.funcInt + iv] * m.margFuncInt);
    }

    // ShapeSet::Pdf(p, wi), core/light.cpp:157-162 (one shape); InfiniteAreaLight::Pdf, infinite.cpp:229-239
    float LightPdf(const PbrtGpuLight &l, const Vec &p, const Vec &wi) const {
        if (l.type == PBRTGPU_LIGHT_POINT) return 0.;
        if (l.type == PBRTGPU_LIGHT_INFINITE) {
            Vec w = XfVector(sv.d->xforms[l.firstShape].mInv, wi);
            float theta = SphericalTheta(w), phi = SphericalPhi(w);
            float sintheta = sinf(theta);
            if (sintheta == 0.f) return 0.f;
            return InfiniteMapPdf(l, phi * INV_TWOPI_F, theta * INV_PI_F) / (2.f * PI_F * PI_F * sintheta);
        }
        float pdf = 0.f;
        for (uint32_t i = 0; i < l.nShapes; ++i) {
            const PbrtGpuLightShape &ls = sv.d->lightShapes[l.firstShape + i];
            const PbrtGpuQuadric &q = sv.d->quadrics[ls.index];
            pdf += ls.area * (q.type == PBRTGPU_PRIM_SPHERE ? SpherePdf(q, p, wi) : ShapePdf(q, p, wi));
        }
        return pdf / l.sumArea;
    }

    // Light::Sample_L: PointLight point.cpp:50-57; DiffuseAreaLight diffuse.cpp:69-81; VisibilityTester light.h:87-92
    Spec Sample_L(const PbrtGpuLight &l, const Vec &p, float pEpsilon, const LightSample &ls, float time, Vec *wi,
                  float *pdf, Ray *vis) const {
        if (l.type == PBRTGPU_LIGHT_POINT) {
            Vec lightPos(l.pos);
            *wi = Normalize(lightPos - p);
            *pdf = 1.f;
            float dist = Distance(p, lightPos);
            *vis = Ray(p, (lightPos - p) / dist, pEpsilon, dist * (1.f - 0.f), time);
            return Spec(l.L) / DistanceSquared(lightPos, p);
        }
        if (l.type == PBRTGPU_LIGHT_INFINITE) {   // infinite.cpp:195-226
            float uv[2], mapPdf;
            InfiniteSampleContinuous(l, ls.uPos[0], ls.uPos[1], uv, &mapPdf);
            if (mapPdf == 0.f) { *pdf = 0.f; return Spec(0.f); }
            float theta = uv[1] * PI_F, phi = uv[0] * 2.f * PI_F;
            float costheta = cosf(theta), sintheta = sinf(theta);
            float sinphi = sinf(phi), cosphi = cosf(phi);
            *wi = XfVector(sv.d->xforms[l.firstShape].m, Vec(sintheta * cosphi, sintheta * sinphi, costheta));
            *pdf = mapPdf / (2.f * PI_F * PI_F * sintheta);
            if (sintheta == 0.f) *pdf = 0.f;
            *vis = Ray(p, *wi, pEpsilon, INF_F, time);   // VisibilityTester::SetRay, light.h:93-96
            return InfiniteLookup(l, uv[0], uv[1]);
        }
        Vec ns;
        // ShapeSet::Sample(p, ls, Ns), light.cpp:142-146: SampleDiscrete over one shape is always 0
        const PbrtGpuLightShape &lsh = sv.d->lightShapes[l.firstShape];
        const PbrtGpuQuadric &q = sv.d->quadrics[lsh.index];
        Vec ps = q.type == PBRTGPU_PRIM_SPHERE ? SphereSample(q, p, ls.uPos[0], ls.uPos[1], &ns)
                                               : DiskSample(q, ls.uPos[0], ls.uPos[1], &ns);
        *wi = Normalize(ps - p);
        *pdf = LightPdf(l, p, *wi);
        float dist = Distance(p, ps);
        *vis = Ray(p, (ps - p) / dist, pEpsilon, dist * (1.f - 1e-3f), time);
        return Dot(ns, -*wi) > 0.f ? Spec(l.L) : Spec(0.f);
    }

    // EstimateDirect, core/integrator.cpp:117-174
    Spec EstimateDirect(int lightIdx, const Vec &p, const Vec &n, const Vec &wo, float rayEpsilon, float time,
                        const BSDF &bsdf, const LightSample &lightSample, const BSDFSample &bsdfSample, int flags,
                        Counters *cnt) const {
        const PbrtGpuLight &light = sv.d->lights[lightIdx];
        Spec Ld(0.);
        Vec wi;
        float lightPdf, bsdfPdf;
        Ray visibility;
        Spec Li = Sample_L(light, p, rayEpsilon, lightSample, time, &wi, &lightPdf, &visibility);
        if (lightPdf > 0. && !Li.IsBlack()) {
            Spec f = bsdf.f(wo, wi, flags);
            if (!f.IsBlack() && !sv.IntersectP(visibility, cnt)) {
                if (light.type == PBRTGPU_LIGHT_POINT)
                    Ld += f * Li * (AbsDot(wi, n) / lightPdf);
                else {
                    bsdfPdf = bsdf.Pdf(wo, wi, flags);
                    float weight = PowerHeuristic(1, lightPdf, 1, bsdfPdf);
                    Ld += f * Li * (AbsDot(wi, n) * weight / lightPdf);
                }
            }
        }
        if (light.type != PBRTGPU_LIGHT_POINT) {
            int sampledType;
            Spec f = bsdf.Sample_f(wo, &wi, bsdfSample, &bsdfPdf, flags, &sampledType);
            if (!f.IsBlack() && bsdfPdf > 0.) {
                float weight = 1.f;
                if (!(sampledType & BSDF_SPECULAR)) {
                    lightPdf = LightPdf(light, p, wi);
                    if (lightPdf == 0.) return Ld;
                    weight = PowerHeuristic(1, bsdfPdf, 1, lightPdf);
                }
                Intersection lightIsect;
                Spec Li(0.f);
                Ray ray(p, wi, rayEpsilon, INF_F, time);
                if (sv.Intersect(ray, &lightIsect, cnt)) {
                    if (sv.d->primitives[lightIsect.prim].areaLight == lightIdx) Li = Le(lightIsect, -wi);
                } else
                    Li = LightLe(light, ray);   // black for area lights
                if (!Li.IsBlack()) Ld += f * Li * AbsDot(wi, n) * weight / bsdfPdf;
            }
        }
        return Ld;
    }

    // UniformSampleOneLight, core/integrator.cpp:82-114.  Table slots (indices into Sample::oneD / twoD, element 0)
    // or -1 for all of them: LightSample(rng) / BSDFSample(rng) and rng.RandomFloat() for the light number
    struct OneLightSlots { int lightNum, lightComp, lightPos, bsdfComp, bsdfDir; };
    Spec UniformSampleOneLight(const Vec &p, const Vec &n, const Vec &wo, float rayEpsilon, float time, const BSDF &bsdf,
                               const Sample *sample, RNG &rng, const OneLightSlots *slots, Counters *cnt) const {
        int nLights = int(sv.d->nLights);
        if (nLights == 0) return Spec(0.);
        int lightNum;
        if (slots) lightNum = Floor2Int(sample->oneD[slots->lightNum][0] * nLights);
        else lightNum = Floor2Int(rng.RandomFloat() * nLights);
        lightNum = std::min(lightNum, nLights - 1);
        LightSample lightSample;
        BSDFSample bsdfSample;
        if (slots) {
            lightSample.uPos[0] = sample->twoD[slots->lightPos][0];
            lightSample.uPos[1] = sample->twoD[slots->lightPos][1];
            lightSample.uComponent = sample->oneD[slots->lightComp][0];
            bsdfSample.uDir[0] = sample->twoD[slots->bsdfDir][0];
            bsdfSample.uDir[1] = sample->twoD[slots->bsdfDir][1];
            bsdfSample.uComponent = sample->oneD[slots->bsdfComp][0];
        } else {
            lightSample.uPos[0] = rng.RandomFloat();   // LightSample(RNG&), light.h:117-121
            lightSample.uPos[1] = rng.RandomFloat();
            lightSample.uComponent = rng.RandomFloat();
            bsdfSample.uDir[0] = rng.RandomFloat();    // BSDFSample(RNG&), reflection.h:135-139
            bsdfSample.uDir[1] = rng.RandomFloat();
            bsdfSample.uComponent = rng.RandomFloat();
        }
        return (float)nLights * EstimateDirect(lightNum, p, n, wo, rayEpsilon, time, bsdf, lightSample, bsdfSample,
                                               BSDF_ALL & ~BSDF_SPECULAR, cnt);
    }

    // UniformSampleAllLights, core/integrator.cpp:47-79 (with sample offsets: directlighting "all")
    Spec UniformSampleAllLights(const Vec &p, const Vec &n, const Vec &wo, float rayEpsilon, float time, const BSDF &bsdf,
                                const Sample *sample, Counters *cnt) const {
        Spec L(0.);
        for (uint32_t i = 0; i < sv.d->nLights; ++i) {
            int nSamples = (int)rd->size1D[2 * i];   // lightSampleOffsets[i].nSamples
            Spec Ld(0.);
            for (int j = 0; j < nSamples; ++j) {
                LightSample lightSample;
                BSDFSample bsdfSample;
                lightSample.uPos[0] = sample->twoD[2 * i][2 * j];
                lightSample.uPos[1] = sample->twoD[2 * i][2 * j + 1];
                lightSample.uComponent = sample->oneD[2 * i][j];
                bsdfSample.uDir[0] = sample->twoD[2 * i + 1][2 * j];
                bsdfSample.uDir[1] = sample->twoD[2 * i + 1][2 * j + 1];
                bsdfSample.uComponent = sample->oneD[2 * i + 1][j];
                Ld += EstimateDirect((int)i, p, n, wo, rayEpsilon, time, bsdf, lightSample, bsdfSample,
                                     BSDF_ALL & ~BSDF_SPECULAR, cnt);
            }
            L += Ld / (float)nSamples;   // Spectrum / float with an int argument, spectrum.h:170
        }
        return L;
    }

    // SpecularReflect / SpecularTransmit, core/integrator.cpp:177-258 (ray differentials only feed texture filtering:
    // omitted).  BSDFSample(rng) draws its three numbers BEFORE BSDF::Sample_f looks for a specular component.
    Spec SpecularBounce(const Ray &ray, const BSDF &bsdf, RNG &rng, const Intersection &isect, const Sample *sample,
                        int depth, int flags, Counters *cnt) const {
        Vec wo = -ray.d, wi;
        float pdf;
        const Vec &p = bsdf.p;
        const Vec &n = bsdf.nn;
        BSDFSample bs;
        bs.uDir[0] = rng.RandomFloat();   // reflection.h:135-139
        bs.uDir[1] = rng.RandomFloat();
        bs.uComponent = rng.RandomFloat();
        Spec f = bsdf.Sample_f(wo, &wi, bs, &pdf, flags, nullptr);
        Spec L(0.f);
        if (pdf > 0.f && !f.IsBlack() && AbsDot(wi, n) != 0.f) {
            Ray rd(p, wi, isect.rayEpsilon, INF_F, ray.time);   // RayDifferential(p, wi, parent, eps): depth + 1
            // SamplerRenderer::Li, samplerrenderer.cpp:225-247 (no volume: T = 1, Lvi = 0; area/point lights' Le(ray) = 0)
            Intersection next;
            Spec Li(0.f);
            if (sv.Intersect(rd, &next, cnt)) Li = DirectLi(rd, next, sample, rng, cnt, depth + 1);
            else Li += EscapedLe(rd);   // samplerrenderer.cpp:238-242
            Li = Spec(1.f) * Li + Spec(0.f);
            L = f * Li * AbsDot(wi, n) / pdf;
        }
        return L;
    }

    // DirectLightingIntegrator::Li, integrators/directlighting.cpp:81-117
    Spec DirectLi(const Ray &ray, const Intersection &isect, const Sample *sample, RNG &rng, Counters *cnt, int depth = 0) const {
        Spec L(0.f);
        BSDF bsdf;
        GetBSDF(isect, &bsdf, cnt);
        Vec wo = -ray.d;
        const Vec &p = bsdf.p;
        const Vec &n = bsdf.nn;
        L += Le(isect, wo);
        if (sv.d->nLights > 0) {
            if (rd->lightStrategy == PBRTGPU_STRATEGY_ALL)
                L += UniformSampleAllLights(p, n, wo, isect.rayEpsilon, ray.time, bsdf, sample, cnt);
            else {
                const OneLightSlots slots = {1, 0, 0, 2, 1};
                L += UniformSampleOneLight(p, n, wo, isect.rayEpsilon, ray.time, bsdf, sample, rng, &slots, cnt);
            }
        }
        if (depth + 1 < rd->maxDepth) {
            L += SpecularBounce(ray, bsdf, rng, isect, sample, depth, BSDF_REFLECTION | BSDF_SPECULAR, cnt);
            L += SpecularBounce(ray, bsdf, rng, isect, sample, depth, BSDF_TRANSMISSION | BSDF_SPECULAR, cnt);
        }
        return L;
    }

    // PathIntegrator::Li, integrators/path.cpp:52-123
    Spec PathLi(const Ray &r, const Intersection &isect, const Sample *sample, RNG &rng, Counters *cnt) const {
        const int SAMPLE_DEPTH = 3;
        Spec pathThroughput(1.), L(0.);
        Ray ray(r);
        bool specularBounce = false;
        Intersection localIsect;
        const Intersection *isectp = &isect;
        for (int bounces = 0;; ++bounces) {
            if (bounces == 0 || specularBounce) L += pathThroughput * Le(*isectp, -ray.d);
            BSDF bsdf;
            GetBSDF(*isectp, &bsdf, cnt);
            const Vec &p = bsdf.p;
            const Vec &n = bsdf.nn;
            Vec wo = -ray.d;
            if (bounces < SAMPLE_DEPTH) {
                const OneLightSlots slots = {4 * bounces + 1, 4 * bounces + 0, 3 * bounces + 0, 4 * bounces + 2, 3 * bounces + 1};
                L += pathThroughput * UniformSampleOneLight(p, n, wo, isectp->rayEpsilon, ray.time, bsdf, sample, rng, &slots, cnt);
            } else
                L += pathThroughput * UniformSampleOneLight(p, n, wo, isectp->rayEpsilon, ray.time, bsdf, sample, rng, nullptr, cnt);
            BSDFSample outgoingBSDFSample;
            if (bounces < SAMPLE_DEPTH) {
                outgoingBSDFSample.uDir[0] = sample->twoD[3 * bounces + 2][0];
                outgoingBSDFSample.uDir[1] = sample->twoD[3 * bounces + 2][1];
                outgoingBSDFSample.uComponent = sample->oneD[4 * bounces + 3][0];
            } else {
                outgoingBSDFSample.uDir[0] = rng.RandomFloat();
                outgoingBSDFSample.uDir[1] = rng.RandomFloat();
                outgoingBSDFSample.uComponent = rng.RandomFloat();
            }
            Vec wi;
            float pdf;
            int flags;
            Spec f = bsdf.Sample_f(wo, &wi, outgoingBSDFSample, &pdf, BSDF_ALL, &flags);
            if (f.IsBlack() || pdf == 0.) break;
            specularBounce = (flags & BSDF_SPECULAR) != 0;
            pathThroughput *= f * AbsDot(wi, n) / pdf;
            ray = Ray(p, wi, isectp->rayEpsilon, INF_F, ray.time);
            if (bounces > 3) {
                float continueProbability = std::min(.5f, pathThroughput.y());
                if (rng.RandomFloat() > continueProbability) break;
                pathThroughput /= continueProbability;
            }
            if (bounces == rd->maxDepth) break;
            if (!sv.Intersect(ray, &localIsect, cnt)) {
                if (specularBounce)
                    for (uint32_t i = 0; i < sv.d->nLights; ++i) L += pathThroughput * LightLe(sv.d->lights[i], ray);   // path.cpp:113-117
                break;
            }
            isectp = &localIsect;
        }
        return L;
    }

    // ImageFilm::AddSample, film/image.cpp:77-137 (box filter: every table weight is 1)
    void AddSample(const Sample &sample, const Spec &L) const {
        float dimageX = sample.imageX - 0.5f;
        float dimageY = sample.imageY - 0.5f;
        int x0 = Ceil2Int(dimageX - rd->filterXWidth);
        int x1 = Floor2Int(dimageX + rd->filterXWidth);
        int y0 = Ceil2Int(dimageY - rd->filterYWidth);
        int y1 = Floor2Int(dimageY + rd->filterYWidth);
        x0 = std::max(x0, xPixelStart);
        x1 = std::min(x1, xPixelStart + xPixelCount - 1);
        y0 = std::max(y0, yPixelStart);
        y1 = std::min(y1, yPixelStart + yPixelCount - 1);
        if ((x1 - x0) < 0 || (y1 - y0) < 0) return;
        float xyz[3];
        xyz[0] = 0.412453f * L.c[0] + 0.357580f * L.c[1] + 0.180423f * L.c[2];
        xyz[1] = 0.212671f * L.c[0] + 0.715160f * L.c[1] + 0.072169f * L.c[2];
        xyz[2] = 0.019334f * L.c[0] + 0.119193f * L.c[1] + 0.950227f * L.c[2];
        // filter table offsets, image.cpp:101-112 (Filter::invXWidth = 1.f / xWidth, filter.h:47-49)
        const float invXWidth = 1.f / rd->filterXWidth, invYWidth = 1.f / rd->filterYWidth;
        int ifx[64], ify[64];
        if (x1 - x0 >= 64 || y1 - y0 >= 64) return;   // (filters wider than 32 pixels: not a configuration anyone renders)
        for (int x = x0; x <= x1; ++x) {
            float fx = fabsf((x - dimageX) * invXWidth * 16);
            ifx[x - x0] = std::min(Floor2Int(fx), 16 - 1);
        }
        for (int y = y0; y <= y1; ++y) {
            float fy = fabsf((y - dimageY) * invYWidth * 16);
            ify[y - y0] = std::min(Floor2Int(fy), 16 - 1);
        }
        for (int y = y0; y <= y1; ++y)
            for (int x = x0; x <= x1; ++x) {
                const float filterWt = rd->hasFilterTable ? rd->filterTable[ify[y - y0] * 16 + ifx[x - x0]] : 1.f;
                PbrtGpuPixel &pixel = pixels[(size_t)(y - yPixelStart) * xPixelCount + (x - xPixelStart)];
                pixel.Lxyz[0] += filterWt * xyz[0];   // non-atomic, like the reference (image.cpp:122-127)
                pixel.Lxyz[1] += filterWt * xyz[1];
                pixel.Lxyz[2] += filterWt * xyz[2];
                pixel.weightSum += filterWt;
            }
    }

    // SamplerRendererTask::Run, renderers/samplerrenderer.cpp:60-164
    void RunTask(int taskNum, Counters *cnt) const {
        int x0, x1, y0, y1;
        ComputeSubWindow(taskNum, rd->nTasks, &x0, &x1, &y0, &y1);
        if (x0 == x1 || y0 == y1) return;   // GetSubSampler returns NULL, lowdiscrepancy.cpp:62
        RNG rng((uint32_t)taskNum);
        const int spp = rd->samplesPerPixel;
        SampleLayout lay;
        lay.FromDesc(rd);
        std::vector<Sample> samples(spp);
        std::vector<float> buf(lay.Floats() * spp), store(lay.Floats() * spp);
        std::vector<Spec> Ls(spp);
        const bool direct = rd->integrator == PBRTGPU_INTEGRATOR_DIRECTLIGHTING;
        for (int yPos = y0; yPos < y1; ++yPos)
            for (int xPos = x0; xPos < x1; ++xPos) {
                LDPixelSample(lay, xPos, yPos, rd->shutterOpen, rd->shutterClose, spp, samples.data(), store.data(), buf.data(), rng);
                for (int i = 0; i < spp; ++i) {
                    Ray ray;
                    GenerateRay(samples[i], &ray);
                    // SamplerRenderer::Li, samplerrenderer.cpp:225-247 (volume integrator: T=1, Lvi=0)
                    Intersection isect;
                    Spec Li(0.f);
                    if (sv.Intersect(ray, &isect, cnt))
                        Li = direct ? DirectLi(ray, isect, &samples[i], rng, cnt) : PathLi(ray, isect, &samples[i], rng, cnt);
                    else
                        Li += EscapedLe(ray);   // "Handle ray that doesn't intersect any geometry"
                    Ls[i] = 1.f * (Spec(1.f) * Li + Spec(0.f));
                    if (Ls[i].HasNaNs()) Ls[i] = Spec(0.f);
                    else if (Ls[i].y() < -1e-5) Ls[i] = Spec(0.f);
                    else if (std::isinf(Ls[i].y())) Ls[i] = Spec(0.f);
                }
                for (int i = 0; i < spp; ++i) AddSample(samples[i], Ls[i]);
            }
        if (cnt) cnt->rngDraws += rng.draws;
    }
};

}  // namespace

// ================================================================== C exports
extern "C" {

struct OracleCounters {
    uint64_t cameraSamples, closestRays, shadowRays, nodesVisited, primTests, shadedHitsWithNormals, rngDraws;
};

// Render tasks [taskBegin, taskEnd) with `nThreads` host threads (tasks are
// independent units: own RNG stream, own pixels -- samplerrenderer.cpp:72-73).
int oracle_render(const PbrtGpuSceneDesc *scene, const PbrtGpuRenderDesc *rd, PbrtGpuPixel *pixels, int nThreads,
                  OracleCounters *counters) {
    Renderer R;
    R.Init(scene, rd, pixels);
    memset(pixels, 0, sizeof(PbrtGpuPixel) * (size_t)R.xPixelCount * R.yPixelCount);
    if (nThreads < 1) nThreads = 1;
    std::atomic<int> next(rd->taskBegin);
    std::vector<Counters> cnts(nThreads);
    std::vector<std::thread> threads;
    for (int t = 0; t < nThreads; ++t)
        threads.emplace_back([&, t]() {
            for (;;) {
                // same enqueue order as the reference: nTasks-1-i (samplerrenderer.cpp:212)
                int i = next.fetch_add(1);
                if (i >= rd->taskEnd) break;
                const int task = rd->taskBegin + (rd->taskEnd - 1 - i);
                if (rd->shardCount > 1 && (task / std::max(1, rd->shardBlock)) % rd->shardCount != rd->shardIndex) continue;
                R.RunTask(task, &cnts[t]);
            }
        });
    for (auto &th : threads) th.join();
    if (counters) {
        memset(counters, 0, sizeof(*counters));
        for (const Counters &c : cnts) {
            counters->closestRays += c.closestRays; counters->shadowRays += c.shadowRays;
            counters->nodesVisited += c.nodesVisited; counters->primTests += c.primTests;
            counters->shadedHitsWithNormals += c.shadedN; counters->rngDraws += c.rngDraws;
        }
        // camera samples = pixels of the rendered tasks x spp
        uint64_t px = 0;
        for (int t = rd->taskBegin; t < rd->taskEnd; ++t) {
            if (rd->shardCount > 1 && (t / std::max(1, rd->shardBlock)) % rd->shardCount != rd->shardIndex) continue;
            int x0, x1, y0, y1;
            R.ComputeSubWindow(t, rd->nTasks, &x0, &x1, &y0, &y1);
            px += (uint64_t)std::max(0, x1 - x0) * std::max(0, y1 - y0);
        }
        counters->cameraSamples = px * rd->samplesPerPixel;
    }
    return 0;
}

int oracle_intersect(const PbrtGpuSceneDesc *scene, const PbrtGpuRay *rays, int n, PbrtGpuHit *hits) {
    SceneView sv{scene};
    for (int i = 0; i < n; ++i) {
        Ray r(Vec(rays[i].o), Vec(rays[i].d), rays[i].mint, rays[i].maxt, rays[i].time);
        Intersection isect;
        float b1 = 0, b2 = 0;
        memset(&hits[i], 0, sizeof(PbrtGpuHit));
        if (sv.Intersect(r, &isect, nullptr, &b1, &b2)) {
            hits[i].hit = 1;
            hits[i].primitive = isect.prim;
            hits[i].t = r.maxt;
            hits[i].b1 = b1; hits[i].b2 = b2;
            hits[i].p[0] = isect.dg.p.x; hits[i].p[1] = isect.dg.p.y; hits[i].p[2] = isect.dg.p.z;
            hits[i].n[0] = isect.dg.nn.x; hits[i].n[1] = isect.dg.nn.y; hits[i].n[2] = isect.dg.nn.z;
        } else {
            hits[i].t = r.maxt;
        }
    }
    return 0;
}

// AggregateTest::Render, renderers/aggregatetest.cpp:60-119 -- the reference's own accelerator check: random rays
// (RNG with its default seed; origins in the grown scene box or at the last hit point, directions uniform on the sphere
// with occasional axis-aligned ones, ray epsilon 0 / last hit's / 1e-3), the accelerator against an exhaustive loop over
// the refined primitives, each gated by its world bound (BBox::IntersectP, geometry.cpp:68-86); a primitive that is hit
// although its bound is missed makes the case "inconsistent" and it is not reported.  The rays and the exhaustive
// answers (hit flag, ray.maxt) are returned so that the same set can be sent through other implementations; the return
// value is the number of "Disagreement" warnings the reference would print (static scenes only: no MotionBounds here).
struct Box3 { Vec lo, hi; };
static Box3 UnionPt(const Box3 &b, const Vec &p) {
    Box3 r;
    r.lo = Vec(std::min(b.lo.x, p.x), std::min(b.lo.y, p.y), std::min(b.lo.z, p.z));
    r.hi = Vec(std::max(b.hi.x, p.x), std::max(b.hi.y, p.y), std::max(b.hi.z, p.z));
    return r;
}
static Box3 PrimWorldBound(const PbrtGpuSceneDesc *d, uint32_t pi) {
    const PbrtGpuPrimitive &p = d->primitives[pi];
    Box3 b;
    if (p.type == PBRTGPU_PRIM_TRIANGLE) {   // Triangle::WorldBound, trianglemesh.cpp:118-124: Union(BBox(p1, p2), p3)
        const PbrtGpuTriangle &t = d->triangles[p.index];
        Vec p1(d->P + 3 * t.v[0]), p2(d->P + 3 * t.v[1]), p3(d->P + 3 * t.v[2]);
        b.lo = Vec(std::min(p1.x, p2.x), std::min(p1.y, p2.y), std::min(p1.z, p2.z));
        b.hi = Vec(std::max(p1.x, p2.x), std::max(p1.y, p2.y), std::max(p1.z, p2.z));
        return UnionPt(b, p3);
    }
    const PbrtGpuQuadric &q = d->quadrics[p.index];   // Shape::WorldBound = ObjectToWorld(ObjectBound()), shape.cpp:52-54
    Vec lo, hi;
    if (q.type == PBRTGPU_PRIM_SPHERE) { lo = Vec(-q.radius, -q.radius, q.zmin); hi = Vec(q.radius, q.radius, q.zmax); }
    else { lo = Vec(-q.radius, -q.radius, q.height); hi = Vec(q.radius, q.radius, q.height); }
    const float *m = d->xforms[q.xform].m;
    auto M = [&](float x, float y, float z) {   // Transform::operator()(const Point &): the returning version
        float xp = m[0] * x + m[1] * y + m[2] * z + m[3], yp = m[4] * x + m[5] * y + m[6] * z + m[7];
        float zp = m[8] * x + m[9] * y + m[10] * z + m[11], wp = m[12] * x + m[13] * y + m[14] * z + m[15];
        if (wp == 1.) return Vec(xp, yp, zp);
        return Vec(xp, yp, zp) / wp;
    };
    Vec c0 = M(lo.x, lo.y, lo.z);   // Transform::operator()(const BBox &), transform.cpp:283-294
    b.lo = c0; b.hi = c0;
    b = UnionPt(b, M(hi.x, lo.y, lo.z)); b = UnionPt(b, M(lo.x, hi.y, lo.z)); b = UnionPt(b, M(lo.x, lo.y, hi.z));
    b = UnionPt(b, M(lo.x, hi.y, hi.z)); b = UnionPt(b, M(hi.x, hi.y, lo.z)); b = UnionPt(b, M(hi.x, lo.y, hi.z));
    b = UnionPt(b, M(hi.x, hi.y, hi.z));
    return b;
}
static bool BoxIntersectP(const Box3 &b, const Ray &ray) {   // BBox::IntersectP, geometry.cpp:68-86
    float t0 = ray.mint, t1 = ray.maxt;
    for (int i = 0; i < 3; ++i) {
        float invRayDir = 1.f / ray.d[i];
        float tNear = (b.lo[i] - ray.o[i]) * invRayDir;
        float tFar = (b.hi[i] - ray.o[i]) * invRayDir;
        if (tNear > tFar) std::swap(tNear, tFar);
        t0 = tNear > t0 ? tNear : t0;
        t1 = tFar < t1 ? tFar : t1;
        if (t0 > t1) return false;
    }
    return true;
}

int oracle_aggregate_test(const PbrtGpuSceneDesc *scene, int nIterations, PbrtGpuRay *raysOut, uint8_t *hitOut, float *tOut) {
    SceneView sv{scene};
    if (scene->nAnimated) return -1;
    RNG rng(5489u);   // RNG(uint32_t seed = 5489UL), core/rng.h:47
    const uint32_t nTop = scene->nPrimitives;
    std::vector<Box3> bboxes(nTop);
    for (uint32_t j = 0; j < nTop; ++j) bboxes[j] = PrimWorldBound(scene, j);
    Vec bmin(scene->nodes[0].bmin), bmax(scene->nodes[0].bmax);   // scene->WorldBound() = BVHAccel::WorldBound
    {   // bbox.Expand(extent of the largest axis), aggregatetest.cpp:65-66
        Vec d = bmax - bmin;
        int ax = (d.x > d.y && d.x > d.z) ? 0 : (d.y > d.z ? 1 : 2);
        float delta = bmax[ax] - bmin[ax];
        bmin = bmin - Vec(delta, delta, delta);
        bmax = bmax + Vec(delta, delta, delta);
    }
    Vec lastHit(0, 0, 0);
    float lastEps = 0.f;
    int disagreements = 0;
    for (int i = 0; i < nIterations; ++i) {
        // the reference draws inside argument lists (aggregatetest.cpp:74-80); g++ on x86-64 evaluates them right to
        // left, so z is drawn first and UniformSampleSphere's u2 before u1 -- pinned by the reference binary's own output
        const float rz = rng.RandomFloat(), ry = rng.RandomFloat(), rx = rng.RandomFloat();
        Vec org(Lerp(rx, bmin.x, bmax.x), Lerp(ry, bmin.y, bmax.y), Lerp(rz, bmin.z, bmax.z));
        if ((rng.RandomUInt() % 4) == 0) org = lastHit;
        const float u2 = rng.RandomFloat(), u1 = rng.RandomFloat();
        Vec dir = UniformSampleSphere(u1, u2);
        if ((rng.RandomUInt() % 32) == 0) dir.x = dir.y = 0.f;
        else if ((rng.RandomUInt() % 32) == 0) dir.x = dir.z = 0.f;
        else if ((rng.RandomUInt() % 32) == 0) dir.y = dir.z = 0.f;
        float eps = 0.f;
        if (rng.RandomFloat() < .25) eps = lastEps;
        else if (rng.RandomFloat() < .25) eps = 1e-3f;
        Ray rayAccel(org, dir, eps), rayAll(org, dir, eps);
        Intersection isectAccel, isectAll;
        bool hitAccel = sv.Intersect(rayAccel, &isectAccel, nullptr);
        bool hitAll = false, inconsistentBounds = false;
        for (uint32_t j = 0; j < nTop; ++j) {
            if (BoxIntersectP(bboxes[j], rayAll)) hitAll |= sv.PrimIntersect(j, rayAll, &isectAll, nullptr);
            else if (sv.PrimIntersect(j, rayAll, &isectAll, nullptr)) inconsistentBounds = true;
        }
        if (!inconsistentBounds && (hitAccel != hitAll || rayAccel.maxt != rayAll.maxt)) ++disagreements;
        if (raysOut) {
            PbrtGpuRay &r = raysOut[i];
            r.o[0] = org.x; r.o[1] = org.y; r.o[2] = org.z; r.d[0] = dir.x; r.d[1] = dir.y; r.d[2] = dir.z;
            r.mint = eps; r.maxt = INF_F; r.time = 0.f;
            hitOut[i] = hitAll ? 1 : 0;
            tOut[i] = rayAll.maxt;
        }
        if (hitAll) {
            lastHit = rayAll(rayAll.maxt);
            lastEps = isectAll.rayEpsilon;
        }
    }
    return disagreements;
}

int oracle_intersect_p(const PbrtGpuSceneDesc *scene, const PbrtGpuRay *rays, int n, uint8_t *occluded) {
    SceneView sv{scene};
    for (int i = 0; i < n; ++i) {
        Ray r(Vec(rays[i].o), Vec(rays[i].d), rays[i].mint, rays[i].maxt, rays[i].time);
        occluded[i] = sv.IntersectP(r, nullptr) ? 1 : 0;
    }
    return 0;
}

// Shading quantities at the first hit of each ray, for function-level comparison with the device code compiled for the
// host (tests/native/device_on_host.cpp).  u: 6 floats per ray (light uPos0 uPos1, bsdf uDir0 uDir1 uComponent, light
// pick).  out: 36 floats per ray, zeros on a miss:
//   0-2 p  3-5 nn  6-8 sn  9-11 tn  12-14 ng  15 rayEpsilon | light sample: 16-18 wi 19 pdf 20-22 Li 23 visibility maxt
//   24-26 f(wo, wi, ALL & ~SPECULAR) 27 Pdf | BSDF::Sample_f(ALL): 28-30 f 31-33 wi 34 pdf | 35 light pdf of that wi
int oracle_shade_test(const PbrtGpuSceneDesc *scene, const PbrtGpuRay *rays, int n, const float *u, float *out) {
    Renderer R;
    PbrtGpuRenderDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.xResolution = rd.yResolution = 1; rd.cropWindow[1] = rd.cropWindow[3] = 1.f; rd.nTasks = 1;
    R.Init(scene, &rd, nullptr);
    const int nLights = (int)scene->nLights;
    for (int i = 0; i < n; ++i) {
        float *o = out + 36 * (size_t)i;
        for (int k = 0; k < 36; ++k) o[k] = 0.f;
        Ray ray(Vec(rays[i].o), Vec(rays[i].d), rays[i].mint, rays[i].maxt, rays[i].time);
        Intersection isect;
        if (!R.sv.Intersect(ray, &isect, nullptr)) continue;
        BSDF bsdf;
        R.GetBSDF(isect, &bsdf, nullptr);
        Vec wo = -ray.d;
        auto put = [&](int at, const Vec &v) { o[at] = v.x; o[at + 1] = v.y; o[at + 2] = v.z; };
        put(0, bsdf.p); put(3, bsdf.nn); put(6, bsdf.sn); put(9, bsdf.tn); put(12, bsdf.ng);
        o[15] = isect.rayEpsilon;
        if (nLights == 0) continue;
        const float *uu = u + 6 * (size_t)i;
        const int lightNum = std::min(Floor2Int(uu[5] * nLights), nLights - 1);
        const PbrtGpuLight &light = scene->lights[lightNum];
        LightSample ls;
        ls.uPos[0] = uu[0]; ls.uPos[1] = uu[1]; ls.uComponent = 0.f;
        Vec wi;
        float lightPdf;
        Ray vis;
        Spec Li = R.Sample_L(light, bsdf.p, isect.rayEpsilon, ls, ray.time, &wi, &lightPdf, &vis);
        put(16, wi); o[19] = lightPdf; o[20] = Li.c[0]; o[21] = Li.c[1]; o[22] = Li.c[2]; o[23] = vis.maxt;
        Spec f = bsdf.f(wo, wi, BSDF_ALL & ~BSDF_SPECULAR);
        o[24] = f.c[0]; o[25] = f.c[1]; o[26] = f.c[2];
        o[27] = bsdf.Pdf(wo, wi, BSDF_ALL & ~BSDF_SPECULAR);
        BSDFSample bs;
        bs.uDir[0] = uu[2]; bs.uDir[1] = uu[3]; bs.uComponent = uu[4];
        Vec wi2;
        float pdf2 = 0.f;
        int sampled = 0;
        Spec f2 = bsdf.Sample_f(wo, &wi2, bs, &pdf2, BSDF_ALL, &sampled);
        if (pdf2 != 0.f) {
            o[28] = f2.c[0]; o[29] = f2.c[1]; o[30] = f2.c[2];
            put(31, wi2); o[34] = pdf2;
            o[35] = R.LightPdf(light, bsdf.p, wi2);
        }
    }
    return 0;
}

// The BxDF restatements evaluated on the direction pairs `refdump bxdf` (the reference's own classes) prints:
// out = n rows of 15 floats: lambert f, pdf | blinn f, pdf | blinn Sample_f wi.xyz, pdf, f | lambert Sample_f wi.xyz, pdf |
// FresnelDielectric(1.5, 1)(wo.z)
int oracle_bxdf_test(float exponent, int n, const float *wo, const float *wi, const float *u, float *out) {
    BSDF bsdf;
    bsdf.sv = nullptr;
    if (exponent > 10000.f || std::isnan(exponent)) exponent = 10000.f;   // Blinn ctor, reflection.h:424-425
    BxDF lambert{BX_LAMBERTIAN, BSDF_REFLECTION | BSDF_DIFFUSE, Spec(0.5f), 0.f, -1};
    BxDF blinn{BX_MICROFACET_BLINN, BSDF_REFLECTION | BSDF_GLOSSY, Spec(0.5f), exponent, -1};
    for (int i = 0; i < n; ++i) {
        Vec o(wo + 3 * i), w(wi + 3 * i);
        float *r = out + 15 * i;
        r[0] = bsdf.BxF(lambert, o, w).c[0]; r[1] = bsdf.BxPdf(lambert, o, w);
        r[2] = bsdf.BxF(blinn, o, w).c[0];   r[3] = bsdf.BxPdf(blinn, o, w);
        Vec ws;
        float pdf = 0.f;
        Spec f = bsdf.BxSample_f(blinn, o, &ws, u[2 * i], u[2 * i + 1], &pdf);
        r[4] = ws.x; r[5] = ws.y; r[6] = ws.z; r[7] = pdf; r[8] = f.c[0];
        pdf = 0.f;
        bsdf.BxSample_f(lambert, o, &ws, u[2 * i], u[2 * i + 1], &pdf);
        r[9] = ws.x; r[10] = ws.y; r[11] = ws.z; r[12] = pdf;
        r[13] = BSDF::FresnelDielectricEvaluate(o.z, 1.5f, 1.f).c[0];
        r[14] = 0.f;
    }
    return 0;
}

// spp rows of sample values for pixel number `pixelIndex` (row-major in the task's window) of task `taskNum`, the task's RNG
// having served LDPixelSample for the pixels before it and nothing else (no integrator draws in between) -- what the LD
// table kernels produce when run round after round on their own
int oracle_ld_samples_seq(const PbrtGpuRenderDesc *rd, int taskNum, int pixelIndex, float *out) {
    Renderer R;
    PbrtGpuSceneDesc dummy;
    memset(&dummy, 0, sizeof(dummy));
    R.Init(&dummy, rd, nullptr);
    int x0, x1, y0, y1;
    R.ComputeSubWindow(taskNum, rd->nTasks, &x0, &x1, &y0, &y1);
    if (x0 == x1 || y0 == y1) return 1;
    RNG rng((uint32_t)taskNum);
    const int spp = rd->samplesPerPixel;
    SampleLayout lay;
    lay.FromDesc(rd);
    const size_t nf = lay.Floats();
    std::vector<Sample> samples(spp);
    std::vector<float> buf(nf * spp), store(nf * spp);
    if (pixelIndex < 0 || pixelIndex >= (x1 - x0) * (y1 - y0)) return 2;
    for (int p = 0; p <= pixelIndex; ++p)
        LDPixelSample(lay, x0 + p % (x1 - x0), y0 + p / (x1 - x0), rd->shutterOpen, rd->shutterClose, spp, samples.data(),
                      store.data(), buf.data(), rng);
    for (int i = 0; i < spp; ++i) {
        float *o = out + nf * i;
        o[0] = samples[i].imageX; o[1] = samples[i].imageY; o[2] = samples[i].lensU; o[3] = samples[i].lensV;
        o[4] = samples[i].time;
        o += 5;
        for (int j = 0; j < lay.n1D; ++j)
            for (uint32_t k = 0; k < lay.size1D[j]; ++k) *o++ = samples[i].oneD[j][k];
        for (int j = 0; j < lay.n2D; ++j)
            for (uint32_t k = 0; k < 2 * lay.size2D[j]; ++k) *o++ = samples[i].twoD[j][k];
    }
    return 0;
}

// ... for the first pixel of the task
int oracle_ld_samples(const PbrtGpuRenderDesc *rd, int taskNum, float *out) { return oracle_ld_samples_seq(rd, taskNum, 0, out); }

int oracle_task_window(const PbrtGpuRenderDesc *rd, int taskNum, int *win) {
    Renderer R;
    PbrtGpuSceneDesc dummy;
    memset(&dummy, 0, sizeof(dummy));
    R.Init(&dummy, rd, nullptr);
    R.ComputeSubWindow(taskNum, rd->nTasks, &win[0], &win[1], &win[2], &win[3]);
    return 0;
}

}  // extern "C"
