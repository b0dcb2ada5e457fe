// dmath.cuh -- device-side float arithmetic with pbrt-v2's exact operation order.
//
// The whole translation unit is compiled with -fmad=false -prec-div=true
// -prec-sqrt=true -ftz=false: the reference is x86-64 SSE2 code without FMA
// (src/Makefile:62-70), and a contraction that flips one hit/miss or one
// Russian-roulette comparison would shift that task's RNG stream for the rest
// of its pixels.  min/max are written as the ternaries
// std::min/std::max compile to (they differ from fminf/fmaxf on NaN and +-0).
//
// Reference semantics:  core/geometry.h:40-330 (Vector/Point/Normal),
// :477-518 (Cross in double, Normalize = v * (1/len), CoordinateSystem),
// core/pbrt.h:188-191 (M_PI is a FLOAT literal), :219-330 (Lerp, Clamp,
// Quadratic), core/spectrum.h:88-260 (RGBSpectrum: '/' is a true division).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "refmath.h"

#define DEV __device__ __forceinline__
// big, rarely-hot routines: one copy per kernel keeps the instruction footprint in cache
#define DEV_NOINLINE __device__ __noinline__
#ifndef PBRT_NI_LEVEL
#define PBRT_NI_LEVEL 1   /* measured on B200: 1 (math only) beats 0, 2 and 3 */
#endif
#if PBRT_NI_LEVEL >= 1
#define DEV_NI1 DEV_NOINLINE
#else
#define DEV_NI1 DEV
#endif
#if PBRT_NI_LEVEL >= 2
#define DEV_NI2 DEV_NOINLINE
#else
#define DEV_NI2 DEV
#endif
#if PBRT_NI_LEVEL >= 3
#define DEV_NI3 DEV_NOINLINE
#else
#define DEV_NI3 DEV
#endif

namespace dev {

constexpr float PI_F = 3.14159265358979323846f;
constexpr float INV_PI_F = 0.31830988618379067154f;
constexpr float INV_TWOPI_F = 0.15915494309189533577f;
constexpr float ONE_MINUS_EPS = 0x1.fffffep-1f;   // core/montecarlo.h:50
#define INF_F __int_as_float(0x7f800000)

DEV float minf(float a, float b) { return (b < a) ? b : a; }   // std::min
DEV float maxf(float a, float b) { return (a < b) ? b : a; }   // std::max
DEV int mini(int a, int b) { return (b < a) ? b : a; }
DEV int maxi(int a, int b) { return (a < b) ? b : a; }

struct V3 {
    float x, y, z;
    DEV V3() {}
    DEV V3(float a, float b, float c) : x(a), y(b), z(c) {}
    DEV float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
DEV V3 operator+(const V3 &a, const V3 &b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
DEV V3 operator-(const V3 &a, const V3 &b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
DEV V3 operator-(const V3 &a) { return V3(-a.x, -a.y, -a.z); }
DEV V3 operator*(const V3 &a, float f) { return V3(f * a.x, f * a.y, f * a.z); }
DEV V3 operator*(float f, const V3 &a) { return V3(f * a.x, f * a.y, f * a.z); }
DEV V3 operator/(const V3 &a, float f) { float inv = 1.f / f; return V3(a.x * inv, a.y * inv, a.z * inv); }
DEV float Dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DEV float AbsDot(const V3 &a, const V3 &b) { return fabsf(Dot(a, b)); }
DEV float LengthSquared(const V3 &a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
DEV float Length(const V3 &a) { return sqrtf(LengthSquared(a)); }
DEV V3 Normalize(const V3 &a) { return a / Length(a); }
DEV float DistanceSquared(const V3 &a, const V3 &b) { return LengthSquared(a - b); }
DEV float Distance(const V3 &a, const V3 &b) { return Length(a - b); }
// geometry.h:477-484: products of floats are exact in double; one rounding in
// the subtraction, one in the conversion back to float.
DEV V3 Cross(const V3 &a, const V3 &b) {
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return V3(float((ay * bz) - (az * by)), float((az * bx) - (ax * bz)), float((ax * by) - (ay * bx)));
}
DEV void CoordinateSystem(const V3 &v1, V3 *v2, V3 *v3) {
    if (fabsf(v1.x) > fabsf(v1.y)) {
        float invLen = 1.f / sqrtf(v1.x * v1.x + v1.z * v1.z);
        *v2 = V3(-v1.z * invLen, 0.f, v1.x * invLen);
    } else {
        float invLen = 1.f / sqrtf(v1.y * v1.y + v1.z * v1.z);
        *v2 = V3(0.f, v1.z * invLen, -v1.y * invLen);
    }
    *v3 = Cross(v1, *v2);
}
DEV float Lerp(float t, float v1, float v2) { return (1.f - t) * v1 + t * v2; }
DEV float Clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
DEV int Floor2Int(float v) { return (int)floorf(v); }
DEV int Ceil2Int(float v) { return (int)ceilf(v); }

DEV bool Quadratic(float A, float B, float C, float *t0, float *t1) {   // pbrt.h:309-323
    float discrim = B * B - 4.f * A * C;
    if ((double)discrim < 0.) return false;
    float rootDiscrim = sqrtf(discrim);
    float q;
    if (B < 0) q = -.5f * (B - rootDiscrim);
    else q = -.5f * (B + rootDiscrim);
    *t0 = q / A;
    *t1 = C / q;
    if (*t0 > *t1) { float t = *t0; *t0 = *t1; *t1 = t; }
    return true;
}

struct Spec {   // RGBSpectrum
    float r, g, b;
    DEV Spec() {}
    DEV explicit Spec(float v) : r(v), g(v), b(v) {}
    DEV Spec(float a, float bb, float c) : r(a), g(bb), b(c) {}
    DEV bool IsBlack() const { return !(r != 0.f || g != 0.f || b != 0.f); }
    DEV bool HasNaNs() const { return isnan(r) || isnan(g) || isnan(b); }
    DEV float y() const { return 0.212671f * r + 0.715160f * g + 0.072169f * b; }
    DEV Spec Clamp() const { return Spec(Clampf(r, 0.f, INF_F), Clampf(g, 0.f, INF_F), Clampf(b, 0.f, INF_F)); }
};
DEV Spec operator+(const Spec &a, const Spec &b) { return Spec(a.r + b.r, a.g + b.g, a.b + b.b); }
DEV Spec operator-(const Spec &a, const Spec &b) { return Spec(a.r - b.r, a.g - b.g, a.b - b.b); }
DEV Spec operator*(const Spec &a, const Spec &b) { return Spec(a.r * b.r, a.g * b.g, a.b * b.b); }
DEV Spec operator/(const Spec &a, const Spec &b) { return Spec(a.r / b.r, a.g / b.g, a.b / b.b); }
DEV Spec operator*(const Spec &a, float f) { return Spec(a.r * f, a.g * f, a.b * f); }
DEV Spec operator*(float f, const Spec &a) { return Spec(a.r * f, a.g * f, a.b * f); }
DEV Spec operator/(const Spec &a, float f) { return Spec(a.r / f, a.g / f, a.b / f); }

// Transform application (row-major float[16]), core/transform.h:204-262
DEV V3 XfPoint(const float *__restrict__ m, const V3 &pt) {
    float x = pt.x, y = pt.y, z = pt.z;
    V3 r(m[0] * x + m[1] * y + m[2] * z + m[3], m[4] * x + m[5] * y + m[6] * z + m[7],
         m[8] * x + m[9] * y + m[10] * z + m[11]);
    float w = m[12] * x + m[13] * y + m[14] * z + m[15];
    if (w != 1.f) { float inv = 1.f / w; r.x *= inv; r.y *= inv; r.z *= inv; }
    return r;
}
DEV V3 XfVector(const float *__restrict__ m, const V3 &v) {
    float x = v.x, y = v.y, z = v.z;
    return V3(m[0] * x + m[1] * y + m[2] * z, m[4] * x + m[5] * y + m[6] * z, m[8] * x + m[9] * y + m[10] * z);
}
DEV V3 XfNormal(const float *__restrict__ mInv, const V3 &n) {
    float x = n.x, y = n.y, z = n.z;
    return V3(mInv[0] * x + mInv[4] * y + mInv[8] * z, mInv[1] * x + mInv[5] * y + mInv[9] * z,
              mInv[2] * x + mInv[6] * y + mInv[10] * z);
}

// ---- libm ---------------------------------------------------------------------
// The reference calls glibc's sinf/cosf/powf/expf/atan2f/acosf; none of them is
// correctly rounded, so refmath.h restates glibc 2.39's own algorithms.  Verified bit-identical to the
// host libm on every float of the used ranges (tests/test_refmath.py: 0 mismatches for sinf/cosf/acosf/
// atanf/atan2f/powf; expf differs on ~1 argument in 10^9 because the host's libm picks an FMA build).
DEV_NI1 float ref_sinf(float x) { return refmath::ref_sinf(x); }
DEV_NI1 float ref_cosf(float x) { return refmath::ref_cosf(x); }
DEV_NI1 float ref_expf(float x) { return refmath::ref_expf(x); }
DEV_NI1 float ref_powf(float x, float y) { return refmath::ref_powf(x, y); }
DEV_NI1 float ref_atan2f(float y, float x) { return refmath::ref_atan2f(y, x); }
DEV_NI1 float ref_acosf(float x) { return refmath::ref_acosf(x); }

}  // namespace dev
