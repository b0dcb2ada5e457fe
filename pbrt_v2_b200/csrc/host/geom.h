// geom.h -- host-side float geometry with pbrt-v2's exact arithmetic.
//
// Operation order follows the reference so that host-side scene preparation
// (vertex transforms, BVH bounds, light areas, kd-tree points) produces the same
// bits as pbrt-v2 built with g++ -O2 (x86-64 SSE2, no FMA):
//   Vector/Point/Normal ops     /root/reference/src/core/geometry.h:40-330
//   Cross() in double           core/geometry.h:477-503
//   Normalize = v * (1/len)     core/geometry.h:93-97,507
//   Matrix4x4 / Transform       core/transform.h:49-330, core/transform.cpp
// Compile WITHOUT -ffast-math and without -march=native (no FMA contraction).
#ifndef PBRT_B200_HOST_GEOM_H
#define PBRT_B200_HOST_GEOM_H

#include <cmath>
#include <cstring>
#include <algorithm>
#include <limits>

namespace pb {

// core/pbrt.h:188-191 -- M_PI is redefined as a FLOAT literal in pbrt-v2
static const float kPi = 3.14159265358979323846f;
static const float kInvPi = 0.31830988618379067154f;
static const float kInvTwoPi = 0.15915494309189533577f;
static const float kInf = std::numeric_limits<float>::infinity();

struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return (&x)[i]; }
    float &operator[](int i) { return (&x)[i]; }
};
inline V3 operator+(const V3 &a, const V3 &b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(const V3 &a, const V3 &b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(const V3 &a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(const V3 &a, float f) { return V3(f * a.x, f * a.y, f * a.z); }
inline V3 operator*(float f, const V3 &a) { return V3(f * a.x, f * a.y, f * a.z); }
// geometry.h:93-97: division multiplies by the reciprocal
inline V3 operator/(const V3 &a, float f) { float inv = 1.f / f; return V3(a.x * inv, a.y * inv, a.z * inv); }
inline float Dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float AbsDot(const V3 &a, const V3 &b) { return fabsf(Dot(a, b)); }
inline float LengthSquared(const V3 &a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline float Length(const V3 &a) { return sqrtf(LengthSquared(a)); }
inline V3 Normalize(const V3 &a) { return a / Length(a); }
// geometry.h:477-484: products and the subtraction are done in double
inline V3 Cross(const V3 &a, const V3 &b) {
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return V3(float((ay * bz) - (az * by)), float((az * bx) - (ax * bz)), float((ax * by) - (ay * bx)));
}
inline void CoordinateSystem(const V3 &v1, V3 *v2, V3 *v3) {   // geometry.h:508-518
    if (fabsf(v1.x) > fabsf(v1.y)) {
        float invLen = 1.f / sqrtf(v1.x * v1.x + v1.z * v1.z);
        *v2 = V3(-v1.z * invLen, 0.f, v1.x * invLen);
    } else {
        float invLen = 1.f / sqrtf(v1.y * v1.y + v1.z * v1.z);
        *v2 = V3(0.f, v1.z * invLen, -v1.y * invLen);
    }
    *v3 = Cross(v1, *v2);
}
inline float Lerp(float t, float v1, float v2) { return (1.f - t) * v1 + t * v2; }   // pbrt.h:219
inline float Clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline float Radians(float deg) { return ((float)kPi / 180.f) * deg; }              // pbrt.h:243

struct BBox {   // geometry.h:383-470
    V3 pMin, pMax;
    BBox() : pMin(kInf, kInf, kInf), pMax(-kInf, -kInf, -kInf) {}
    explicit BBox(const V3 &p) : pMin(p), pMax(p) {}
    BBox(const V3 &p1, const V3 &p2)
        : pMin(std::min(p1.x, p2.x), std::min(p1.y, p2.y), std::min(p1.z, p2.z)),
          pMax(std::max(p1.x, p2.x), std::max(p1.y, p2.y), std::max(p1.z, p2.z)) {}
    float SurfaceArea() const {
        V3 d = pMax - pMin;
        return 2.f * (d.x * d.y + d.x * d.z + d.y * d.z);
    }
    int MaximumExtent() const {
        V3 d = pMax - pMin;
        if (d.x > d.y && d.x > d.z) return 0;
        else if (d.y > d.z) return 1;
        else return 2;
    }
};
inline BBox Union(const BBox &b, const V3 &p) {   // core/geometry.cpp:40-49
    BBox r;
    r.pMin = V3(std::min(b.pMin.x, p.x), std::min(b.pMin.y, p.y), std::min(b.pMin.z, p.z));
    r.pMax = V3(std::max(b.pMax.x, p.x), std::max(b.pMax.y, p.y), std::max(b.pMax.z, p.z));
    return r;
}
inline BBox Union(const BBox &b, const BBox &b2) {   // core/geometry.cpp:52-61
    BBox r;
    r.pMin = V3(std::min(b.pMin.x, b2.pMin.x), std::min(b.pMin.y, b2.pMin.y), std::min(b.pMin.z, b2.pMin.z));
    r.pMax = V3(std::max(b.pMax.x, b2.pMax.x), std::max(b.pMax.y, b2.pMax.y), std::max(b.pMax.z, b2.pMax.z));
    return r;
}

struct Matrix4x4 {   // core/transform.h:49-83
    float m[4][4];
    Matrix4x4() {
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m[i][j] = (i == j) ? 1.f : 0.f;
    }
    Matrix4x4(float t00, float t01, float t02, float t03, float t10, float t11, float t12, float t13,
              float t20, float t21, float t22, float t23, float t30, float t31, float t32, float t33) {
        m[0][0] = t00; m[0][1] = t01; m[0][2] = t02; m[0][3] = t03;
        m[1][0] = t10; m[1][1] = t11; m[1][2] = t12; m[1][3] = t13;
        m[2][0] = t20; m[2][1] = t21; m[2][2] = t22; m[2][3] = t23;
        m[3][0] = t30; m[3][1] = t31; m[3][2] = t32; m[3][3] = t33;
    }
    bool operator==(const Matrix4x4 &o) const { return memcmp(m, o.m, sizeof(m)) == 0 || eq(o); }
    bool eq(const Matrix4x4 &o) const {
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) if (m[i][j] != o.m[i][j]) return false;
        return true;
    }
    static Matrix4x4 Mul(const Matrix4x4 &m1, const Matrix4x4 &m2) {   // transform.h:69-77
        Matrix4x4 r;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                r.m[i][j] = m1.m[i][0] * m2.m[0][j] + m1.m[i][1] * m2.m[1][j] +
                            m1.m[i][2] * m2.m[2][j] + m1.m[i][3] * m2.m[3][j];
        return r;
    }
};
Matrix4x4 Transpose(const Matrix4x4 &m);
Matrix4x4 Inverse(const Matrix4x4 &m);   // Gauss-Jordan, core/transform.cpp:76-136

struct Transform {   // core/transform.h:86-178
    Matrix4x4 m, mInv;
    Transform() {}
    explicit Transform(const Matrix4x4 &mat) : m(mat), mInv(Inverse(mat)) {}
    Transform(const Matrix4x4 &mat, const Matrix4x4 &minv) : m(mat), mInv(minv) {}
    bool operator==(const Transform &t) const { return m.eq(t.m) && mInv.eq(t.mInv); }
    bool operator!=(const Transform &t) const { return !(*this == t); }
    bool operator<(const Transform &t2) const {   // transform.h:116-123 (TransformCache key order)
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                if (m.m[i][j] < t2.m.m[i][j]) return true;
                if (m.m[i][j] > t2.m.m[i][j]) return false;
            }
        return false;
    }
    V3 Pt(const V3 &pt) const {   // transform.h:204-212
        float x = pt.x, y = pt.y, z = pt.z;
        float xp = m.m[0][0] * x + m.m[0][1] * y + m.m[0][2] * z + m.m[0][3];
        float yp = m.m[1][0] * x + m.m[1][1] * y + m.m[1][2] * z + m.m[1][3];
        float zp = m.m[2][0] * x + m.m[2][1] * y + m.m[2][2] * z + m.m[2][3];
        float wp = m.m[3][0] * x + m.m[3][1] * y + m.m[3][2] * z + m.m[3][3];
        if (wp == 1.) return V3(xp, yp, zp);
        else return V3(xp, yp, zp) / wp;
    }
    V3 Vec(const V3 &v) const {   // transform.h:225-230
        float x = v.x, y = v.y, z = v.z;
        return V3(m.m[0][0] * x + m.m[0][1] * y + m.m[0][2] * z,
                  m.m[1][0] * x + m.m[1][1] * y + m.m[1][2] * z,
                  m.m[2][0] * x + m.m[2][1] * y + m.m[2][2] * z);
    }
    V3 Nrm(const V3 &n) const {   // transform.h:241-246 (inverse transpose)
        float x = n.x, y = n.y, z = n.z;
        return V3(mInv.m[0][0] * x + mInv.m[1][0] * y + mInv.m[2][0] * z,
                  mInv.m[0][1] * x + mInv.m[1][1] * y + mInv.m[2][1] * z,
                  mInv.m[0][2] * x + mInv.m[1][2] * y + mInv.m[2][2] * z);
    }
    BBox Box(const BBox &b) const;                  // transform.cpp:283-294
    Transform operator*(const Transform &t2) const {   // transform.cpp:297-301
        return Transform(Matrix4x4::Mul(m, t2.m), Matrix4x4::Mul(t2.mInv, mInv));
    }
    bool SwapsHandedness() const;                   // transform.cpp:304-315
    bool HasScale() const;
};
inline Transform Inverse(const Transform &t) { return Transform(t.mInv, t.m); }
Transform Translate(const V3 &delta);
Transform Scale(float x, float y, float z);
Transform Rotate(float angle, const V3 &axis);
Transform LookAt(const V3 &pos, const V3 &look, const V3 &up, bool *ok);
Transform Perspective(float fov, float znear, float zfar);
bool SolveLinearSystem2x2(const float A[2][2], const float B[2], float *x0, float *x1);

struct Quaternion {   // core/quaternion.h:44-92
    V3 v;
    float w;
    Quaternion() : v(0, 0, 0), w(1.f) {}
    Quaternion(const V3 &vv, float ww) : v(vv), w(ww) {}
    explicit Quaternion(const Matrix4x4 &m);   // quaternion.cpp:61-96
    Transform ToTransform() const;             // quaternion.cpp:39-58
};
Quaternion Slerp(float t, const Quaternion &q1, const Quaternion &q2);   // quaternion.cpp:99-111

// core/transform.h:299-327, transform.cpp:328-413.  The per-ray Interpolate() of the hot path runs on
// the device (device/danim.cuh); this host copy feeds Decompose's outputs to it and bounds the motion.
struct AnimatedTransform {
    float startTime, endTime;
    Transform startTransform, endTransform;
    bool actuallyAnimated;
    V3 T[2];
    Quaternion R[2];
    Matrix4x4 S[2];
    AnimatedTransform(const Transform &t1, float time1, const Transform &t2, float time2);
    static void Decompose(const Matrix4x4 &m, V3 *T, Quaternion *R, Matrix4x4 *S);
    void Interpolate(float time, Transform *t) const;
    BBox MotionBounds(const BBox &b, bool useInverse) const;
};

}  // namespace pb
#endif
