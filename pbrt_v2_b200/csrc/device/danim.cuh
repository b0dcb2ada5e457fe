// danim.cuh -- TransformedPrimitive support: AnimatedTransform::Interpolate per ray, on the device.
//
// pbrtShape wraps a shape declared under an animated CTM in a TransformedPrimitive (core/api.cpp:1011-1043)
// whose Intersect/IntersectP first interpolate the world-to-primitive transform at ray.time
// (core/primitive.cpp:95-124, core/transform.cpp:371-396): translation lerp, quaternion slerp, scale-matrix
// lerp, then Translate(T) * R.ToTransform() * Transform(S) -- two 4x4 products for m, two for mInv and one
// Gauss-Jordan inverse, all in float.  The steps below are the same operations in the same order
// (compiled with -fmad=false), so the interpolated matrices carry the reference's bits.
#pragma once
#include "dmath.cuh"
#include "../../../include/pbrtgpu.h"

namespace dev {

struct DXf { float m[16], mInv[16]; };

DEV void MatIdentity(float *r) {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = (i % 5 == 0) ? 1.f : 0.f;
}
DEV void MatMul(const float *m1, const float *m2, float *r) {   // Matrix4x4::Mul, core/transform.h:69-77
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            r[4 * i + j] = m1[4 * i] * m2[j] + m1[4 * i + 1] * m2[4 + j] + m1[4 * i + 2] * m2[8 + j] + m1[4 * i + 3] * m2[12 + j];
}
// Inverse(const Matrix4x4 &), core/transform.cpp:76-136: full-pivot Gauss-Jordan, same visiting order
__device__ __noinline__ void MatInverse(const float *m, float *out) {
    // A matrix that turns into NaNs half way -- an interpolated transform that is singular at the ray's time -- leaves the
    // loop below early, and the un-scrambling loop then indexes with whatever indxr / indxc held (the reference is no better
    // off: its pivot search ends with icol = -1 and `++ipiv[icol]` writes below the array, transform.cpp:82-104).  Found by
    // scripts/fuzz_descriptors.py after the round's last GPU visit: the zero-initialisation is behind PBRT_SAFE_INVERSE
    // (on in the f4 variant and the host harnesses) until the kernels that contain it have run on a GPU -- it re-allocates
    // registers across every ANIM instantiation.  Non-finite transforms and shutter times are refused at validation.
#ifndef PBRT_SAFE_INVERSE
#define PBRT_SAFE_INVERSE 0
#endif
#if PBRT_SAFE_INVERSE
    int indxc[4] = {0, 0, 0, 0}, indxr[4] = {0, 0, 0, 0};
#else
    int indxc[4], indxr[4];
#endif
    int ipiv[4] = {0, 0, 0, 0};
    float a[4][4];
    for (int i = 0; i < 16; ++i) a[i >> 2][i & 3] = m[i];
    for (int i = 0; i < 4; i++) {
        int irow = -1, icol = -1;
        float big = 0.f;
        for (int j = 0; j < 4; j++) {
            if (ipiv[j] != 1) {
                for (int k = 0; k < 4; k++) {
                    if (ipiv[k] == 0) {
                        if (fabsf(a[j][k]) >= big) {
                            big = fabsf(a[j][k]);
                            irow = j;
                            icol = k;
                        }
                    }
                }
            }
        }
        if (icol < 0) break;   // NaN matrix (the reference would read out of bounds)
        ++ipiv[icol];
        if (irow != icol)
            for (int k = 0; k < 4; ++k) { float tmp = a[irow][k]; a[irow][k] = a[icol][k]; a[icol][k] = tmp; }
        indxr[i] = irow;
        indxc[i] = icol;
        float pivinv = 1.f / a[icol][icol];
        a[icol][icol] = 1.f;
        for (int j = 0; j < 4; j++) a[icol][j] *= pivinv;
        for (int j = 0; j < 4; j++) {
            if (j != icol) {
                float save = a[j][icol];
                a[j][icol] = 0;
                for (int k = 0; k < 4; k++) a[j][k] -= a[icol][k] * save;
            }
        }
    }
    for (int j = 3; j >= 0; j--) {
        if (indxr[j] != indxc[j])
            for (int k = 0; k < 4; k++) { float tmp = a[k][indxr[j]]; a[k][indxr[j]] = a[k][indxc[j]]; a[k][indxc[j]] = tmp; }
    }
    for (int i = 0; i < 16; ++i) out[i] = a[i >> 2][i & 3];
}

// Slerp, core/quaternion.cpp:99-111.  q = (v.x, v.y, v.z, w)
DEV float QDot(const float *a, const float *b) { return (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) + a[3] * b[3]; }
DEV void QNormalize(float *q) {   // quaternion.h:109-111: v /= f multiplies by 1/f (geometry.h:100-105), w /= f divides
    const float f = sqrtf(QDot(q, q));
    const float inv = 1.f / f;
    q[0] *= inv; q[1] *= inv; q[2] *= inv;
    q[3] /= f;
}
DEV void Slerp(float t, const float *q1, const float *q2, float *out) {
    const float cosTheta = QDot(q1, q2);
    if (cosTheta > .9995f) {
        for (int i = 0; i < 4; ++i) out[i] = q1[i] * (1.f - t) + q2[i] * t;
        QNormalize(out);
    } else {
        const float theta = ref_acosf(Clampf(cosTheta, -1.f, 1.f));
        const float thetap = theta * t;
        float qperp[4];
        for (int i = 0; i < 4; ++i) qperp[i] = q2[i] - q1[i] * cosTheta;
        QNormalize(qperp);
        const float c = ref_cosf(thetap), sn = ref_sinf(thetap);
        for (int i = 0; i < 4; ++i) out[i] = q1[i] * c + qperp[i] * sn;
    }
}

// AnimatedTransform::Interpolate, core/transform.cpp:371-396
__device__ __noinline__ void AnimInterpolate(const PbrtGpuAnimated *ap, float time, DXf *out) {
    const PbrtGpuAnimated &a = *ap;
    if (!a.actuallyAnimated || time <= a.startTime) {
        for (int i = 0; i < 16; ++i) { out->m[i] = a.start.m[i]; out->mInv[i] = a.start.mInv[i]; }
        return;
    }
    if (time >= a.endTime) {
        for (int i = 0; i < 16; ++i) { out->m[i] = a.end.m[i]; out->mInv[i] = a.end.mInv[i]; }
        return;
    }
    const float dt = (time - a.startTime) / (a.endTime - a.startTime);
    float trans[3];
    for (int i = 0; i < 3; ++i) trans[i] = (1.f - dt) * a.T[0][i] + dt * a.T[1][i];
    float q[4];
    Slerp(dt, a.R[0], a.R[1], q);
    float scale[16], scaleInv[16];
    MatIdentity(scale);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) scale[4 * i + j] = Lerp(dt, a.S[0][4 * i + j], a.S[1][4 * i + j]);
    MatInverse(scale, scaleInv);   // Transform(const Matrix4x4 &), transform.h:99-101
    float trM[16], trI[16];        // Translate(), transform.cpp:160-170
    MatIdentity(trM); MatIdentity(trI);
    trM[3] = trans[0]; trM[7] = trans[1]; trM[11] = trans[2];
    trI[3] = -trans[0]; trI[7] = -trans[1]; trI[11] = -trans[2];
    float rotM[16], rotI[16];      // Quaternion::ToTransform, quaternion.cpp:39-58: Transform(Transpose(m), m)
    {
        const float x = q[0], y = q[1], z = q[2], w = q[3];
        const float xx = x * x, yy = y * y, zz = z * z;
        const float xy = x * y, xz = x * z, yz = y * z;
        const float wx = x * w, wy = y * w, wz = z * w;
        MatIdentity(rotI);
        rotI[0] = 1.f - 2.f * (yy + zz);
        rotI[1] = 2.f * (xy + wz);
        rotI[2] = 2.f * (xz - wy);
        rotI[4] = 2.f * (xy - wz);
        rotI[5] = 1.f - 2.f * (xx + zz);
        rotI[6] = 2.f * (yz + wx);
        rotI[8] = 2.f * (xz + wy);
        rotI[9] = 2.f * (yz - wx);
        rotI[10] = 1.f - 2.f * (xx + yy);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) rotM[4 * i + j] = rotI[4 * j + i];
    }
    // Transform::operator* (transform.cpp:297-301): m = m1 * m2, mInv = m2.mInv * m1.mInv; left to right
    float trM2[16], trI2[16];
    MatMul(trM, rotM, trM2);
    MatMul(rotI, trI, trI2);
    MatMul(trM2, scale, out->m);
    MatMul(scaleInv, trI2, out->mInv);
}

DEV bool MatIsIdentity(const float *m) {   // Transform::IsIdentity, transform.h:138-147
    bool id = true;
#pragma unroll
    for (int i = 0; i < 16; ++i) id = id && (m[i] == ((i % 5 == 0) ? 1.f : 0.f));
    return id;
}

}  // namespace dev
