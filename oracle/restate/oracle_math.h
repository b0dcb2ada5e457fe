// oracle_math.h -- TEST INFRASTRUCTURE (see oracle/README.md).
//
// Scalar float helpers for the CPU restatement of pbrt-v2's path-tracing hot
// path.  Operation order mirrors the reference exactly (x86-64 SSE2, no FMA):
//   core/geometry.h:40-330,477-518   Vector/Point/Normal, Cross in double
//   core/pbrt.h:188-191,219-330       M_PI as FLOAT, Lerp, Clamp, Quadratic
// Transcendentals are the host libm's (glibc), the same ones the reference calls.
#ifndef ORACLE_MATH_H
#define ORACLE_MATH_H

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <algorithm>

namespace orc {

static const float PI_F = 3.14159265358979323846f;
static const float INV_PI_F = 0.31830988618379067154f;
static const float INV_TWOPI_F = 0.15915494309189533577f;
static const float INF_F = std::numeric_limits<float>::infinity();
static const float OneMinusEpsilon = 0x1.fffffep-1;   // core/montecarlo.h:50

struct Vec {
    float x, y, z;
    Vec() : x(0), y(0), z(0) {}
    Vec(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit Vec(const float *p) : x(p[0]), y(p[1]), z(p[2]) {}
    float operator[](int i) const { return (&x)[i]; }
};
inline Vec operator+(const Vec &a, const Vec &b) { return Vec(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec operator-(const Vec &a, const Vec &b) { return Vec(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec operator-(const Vec &a) { return Vec(-a.x, -a.y, -a.z); }
inline Vec operator*(const Vec &a, float f) { return Vec(f * a.x, f * a.y, f * a.z); }
inline Vec operator*(float f, const Vec &a) { return Vec(f * a.x, f * a.y, f * a.z); }
inline Vec operator/(const Vec &a, float f) { float inv = 1.f / f; return Vec(a.x * inv, a.y * inv, a.z * inv); }
inline float Dot(const Vec &a, const Vec &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float AbsDot(const Vec &a, const Vec &b) { return fabsf(Dot(a, b)); }
inline float LengthSquared(const Vec &a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline float Length(const Vec &a) { return sqrtf(LengthSquared(a)); }
inline Vec Normalize(const Vec &a) { return a / Length(a); }
inline float DistanceSquared(const Vec &a, const Vec &b) { return LengthSquared(a - b); }
inline float Distance(const Vec &a, const Vec &b) { return Length(a - b); }
inline Vec Cross(const Vec &a, const Vec &b) {   // geometry.h:477-484
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return Vec(float((ay * bz) - (az * by)), float((az * bx) - (ax * bz)), float((ax * by) - (ay * bx)));
}
inline void CoordinateSystem(const Vec &v1, Vec *v2, Vec *v3) {   // geometry.h:508-518
    if (fabsf(v1.x) > fabsf(v1.y)) {
        float invLen = 1.f / sqrtf(v1.x * v1.x + v1.z * v1.z);
        *v2 = Vec(-v1.z * invLen, 0.f, v1.x * invLen);
    } else {
        float invLen = 1.f / sqrtf(v1.y * v1.y + v1.z * v1.z);
        *v2 = Vec(0.f, v1.z * invLen, -v1.y * invLen);
    }
    *v3 = Cross(v1, *v2);
}
inline float Lerp(float t, float v1, float v2) { return (1.f - t) * v1 + t * v2; }
inline float Clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline int Floor2Int(float v) { return (int)floorf(v); }
inline int Ceil2Int(float v) { return (int)ceilf(v); }

inline bool Quadratic(float A, float B, float C, float *t0, float *t1) {   // pbrt.h:309-323
    float discrim = B * B - 4.f * A * C;
    if (discrim < 0.) return false;
    float rootDiscrim = sqrtf(discrim);
    float q;
    if (B < 0) q = -.5f * (B - rootDiscrim);
    else q = -.5f * (B + rootDiscrim);
    *t0 = q / A;
    *t1 = C / q;
    if (*t0 > *t1) std::swap(*t0, *t1);
    return true;
}

// RGBSpectrum, core/spectrum.h:88-260,388-452: division by a float is a TRUE
// division per channel (:170-175), not a reciprocal multiply.
struct Spec {
    float c[3];
    Spec() { c[0] = c[1] = c[2] = 0.f; }
    explicit Spec(float v) { c[0] = c[1] = c[2] = v; }
    Spec(float r, float g, float b) { c[0] = r; c[1] = g; c[2] = b; }
    explicit Spec(const float *p) { c[0] = p[0]; c[1] = p[1]; c[2] = p[2]; }
    bool IsBlack() const { return !(c[0] != 0. || c[1] != 0. || c[2] != 0.); }
    bool HasNaNs() const { return std::isnan(c[0]) || std::isnan(c[1]) || std::isnan(c[2]); }
    float y() const { return 0.212671f * c[0] + 0.715160f * c[1] + 0.072169f * c[2]; }
    Spec Clamp() const {
        return Spec(Clampf(c[0], 0, INF_F), Clampf(c[1], 0, INF_F), Clampf(c[2], 0, INF_F));
    }
};
inline Spec operator+(const Spec &a, const Spec &b) { return Spec(a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2]); }
inline Spec operator-(const Spec &a, const Spec &b) { return Spec(a.c[0] - b.c[0], a.c[1] - b.c[1], a.c[2] - b.c[2]); }
inline Spec operator*(const Spec &a, const Spec &b) { return Spec(a.c[0] * b.c[0], a.c[1] * b.c[1], a.c[2] * b.c[2]); }
inline Spec operator/(const Spec &a, const Spec &b) { return Spec(a.c[0] / b.c[0], a.c[1] / b.c[1], a.c[2] / b.c[2]); }
inline Spec operator*(const Spec &a, float f) { return Spec(a.c[0] * f, a.c[1] * f, a.c[2] * f); }
inline Spec operator*(float f, const Spec &a) { return a * f; }
inline Spec operator/(const Spec &a, float f) { return Spec(a.c[0] / f, a.c[1] / f, a.c[2] / f); }
inline Spec &operator+=(Spec &a, const Spec &b) { a.c[0] += b.c[0]; a.c[1] += b.c[1]; a.c[2] += b.c[2]; return a; }
inline Spec &operator*=(Spec &a, const Spec &b) { a.c[0] *= b.c[0]; a.c[1] *= b.c[1]; a.c[2] *= b.c[2]; return a; }
inline Spec &operator/=(Spec &a, float f) { a.c[0] /= f; a.c[1] /= f; a.c[2] /= f; return a; }

// Transform application, core/transform.h:204-262 (row-major float[16])
inline Vec XfPoint(const float *m, const Vec &pt) {   // void version: *ptrans /= w  (:213-220)
    float x = pt.x, y = pt.y, z = pt.z;
    Vec r(m[0] * x + m[1] * y + m[2] * z + m[3], m[4] * x + m[5] * y + m[6] * z + m[7],
          m[8] * x + m[9] * y + m[10] * z + m[11]);
    float w = m[12] * x + m[13] * y + m[14] * z + m[15];
    if (w != 1.) { float inv = 1.f / w; r.x *= inv; r.y *= inv; r.z *= inv; }
    return r;
}
inline Vec XfVector(const float *m, const Vec &v) {
    float x = v.x, y = v.y, z = v.z;
    return Vec(m[0] * x + m[1] * y + m[2] * z, m[4] * x + m[5] * y + m[6] * z, m[8] * x + m[9] * y + m[10] * z);
}
inline Vec XfNormal(const float *mInv, const Vec &n) {   // transform.h:241-246
    float x = n.x, y = n.y, z = n.z;
    return Vec(mInv[0] * x + mInv[4] * y + mInv[8] * z, mInv[1] * x + mInv[5] * y + mInv[9] * z,
               mInv[2] * x + mInv[6] * y + mInv[10] * z);
}


// ---- Matrix4x4 / Transform as row-major float[16] (core/transform.h:49-83,86-178) -- used by the
// TransformedPrimitive restatement (AnimatedTransform::Interpolate runs per ray)
struct Xf { float m[16], mInv[16]; };
inline void MatIdentity(float *r) { for (int i = 0; i < 16; ++i) r[i] = (i % 5 == 0) ? 1.f : 0.f; }
inline void MatMul(const float *m1, const float *m2, float *r) {   // Matrix4x4::Mul, transform.h:69-77
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r[4 * i + j] = m1[4 * i] * m2[j] + m1[4 * i + 1] * m2[4 + j] + m1[4 * i + 2] * m2[8 + j] + m1[4 * i + 3] * m2[12 + j];
}
inline void MatTranspose(const float *m, float *r) {
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r[4 * i + j] = m[4 * j + i];
}
inline void MatInverse(const float *m, float *out) {   // Inverse(Matrix4x4), core/transform.cpp:76-136
    int indxc[4], indxr[4];
    int ipiv[4] = {0, 0, 0, 0};
    float minv[4][4];
    for (int i = 0; i < 16; ++i) minv[i / 4][i % 4] = m[i];
    for (int i = 0; i < 4; i++) {
        int irow = -1, icol = -1;
        float big = 0.;
        for (int j = 0; j < 4; j++) {
            if (ipiv[j] != 1) {
                for (int k = 0; k < 4; k++) {
                    if (ipiv[k] == 0) {
                        if (fabsf(minv[j][k]) >= big) {
                            big = float(fabsf(minv[j][k]));
                            irow = j;
                            icol = k;
                        }
                    }
                }
            }
        }
        if (icol < 0) break;   // NaN matrix: the reference would index out of bounds here
        ++ipiv[icol];
        if (irow != icol)
            for (int k = 0; k < 4; ++k) std::swap(minv[irow][k], minv[icol][k]);
        indxr[i] = irow;
        indxc[i] = icol;
        float pivinv = 1.f / minv[icol][icol];
        minv[icol][icol] = 1.f;
        for (int j = 0; j < 4; j++) minv[icol][j] *= pivinv;
        for (int j = 0; j < 4; j++) {
            if (j != icol) {
                float save = minv[j][icol];
                minv[j][icol] = 0;
                for (int k = 0; k < 4; k++) minv[j][k] -= minv[icol][k] * save;
            }
        }
    }
    for (int j = 3; j >= 0; j--) {
        if (indxr[j] != indxc[j])
            for (int k = 0; k < 4; k++) std::swap(minv[k][indxr[j]], minv[k][indxc[j]]);
    }
    for (int i = 0; i < 16; ++i) out[i] = minv[i / 4][i % 4];
}
inline Xf XfCompose(const Xf &a, const Xf &b) {   // Transform::operator*, transform.cpp:297-301
    Xf r;
    MatMul(a.m, b.m, r.m);
    MatMul(b.mInv, a.mInv, r.mInv);
    return r;
}
inline Xf XfInverse(const Xf &a) {
    Xf r;
    memcpy(r.m, a.mInv, sizeof(r.m));
    memcpy(r.mInv, a.m, sizeof(r.m));
    return r;
}
inline bool XfIsIdentity(const Xf &a) {   // transform.h:124-133 (checks m only)
    for (int i = 0; i < 16; ++i) if (a.m[i] != ((i % 5 == 0) ? 1.f : 0.f)) return false;
    return true;
}

}  // namespace orc
#endif
