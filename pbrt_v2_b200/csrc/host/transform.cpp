// transform.cpp -- host-side 4x4 transforms with the reference's arithmetic
// order (see geom.h).  Each function cites the pbrt-v2 routine whose results it
// must reproduce bit-for-bit; the code itself is written for this project.
#include "geom.h"
#include <cstdio>
#include <utility>

namespace pb {

Matrix4x4 Transpose(const Matrix4x4 &a) {   // core/transform.cpp:68-73
    Matrix4x4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[j][i];
    return r;
}

// Full-pivot Gauss-Jordan elimination.  Must visit candidates, scale rows and
// eliminate in exactly the order of core/transform.cpp:76-136 so that every
// rounding step matches (the inverse feeds RasterToCamera, WorldToObject, ...).
Matrix4x4 Inverse(const Matrix4x4 &src) {
    float a[4][4];
    memcpy(a, src.m, sizeof(a));
    int pivotRow[4] = {0, 0, 0, 0}, pivotCol[4] = {0, 0, 0, 0};
    int used[4] = {0, 0, 0, 0};
    for (int step = 0; step < 4; ++step) {
        int r = -1, c = -1;
        float best = 0.f;
        for (int j = 0; j < 4; ++j) {
            if (used[j] == 1) continue;
            for (int k = 0; k < 4; ++k) {
                if (used[k] == 0) {
                    if (fabsf(a[j][k]) >= best) {
                        best = fabsf(a[j][k]);
                        r = j;
                        c = k;
                    }
                } else if (used[k] > 1) {
                    fprintf(stderr, "pbrt_b200: singular matrix in Inverse()\n");
                }
            }
        }
        if (c < 0) {   // a matrix of NaNs: no pivot candidate compares >= 0 (the reference indexes out of bounds here)
            fprintf(stderr, "pbrt_b200: singular matrix in Inverse()\n");
            break;
        }
        ++used[c];
        if (r != c)
            for (int k = 0; k < 4; ++k) std::swap(a[r][k], a[c][k]);
        pivotRow[step] = r;
        pivotCol[step] = c;
        if (a[c][c] == 0.f) fprintf(stderr, "pbrt_b200: singular matrix in Inverse()\n");
        float pinv = 1.f / a[c][c];
        a[c][c] = 1.f;
        for (int j = 0; j < 4; ++j) a[c][j] *= pinv;
        for (int j = 0; j < 4; ++j) {
            if (j == c) continue;
            float f = a[j][c];
            a[j][c] = 0;
            for (int k = 0; k < 4; ++k) a[j][k] -= a[c][k] * f;
        }
    }
    for (int j = 3; j >= 0; --j)
        if (pivotRow[j] != pivotCol[j])
            for (int k = 0; k < 4; ++k) std::swap(a[k][pivotRow[j]], a[k][pivotCol[j]]);
    Matrix4x4 out;
    memcpy(out.m, a, sizeof(a));
    return out;
}

bool SolveLinearSystem2x2(const float A[2][2], const float B[2], float *x0, float *x1) {
    // core/transform.cpp:40-50
    float det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
    if (fabsf(det) < 1e-10f) return false;
    *x0 = (A[1][1] * B[0] - A[0][1] * B[1]) / det;
    *x1 = (A[0][0] * B[1] - A[1][0] * B[0]) / det;
    if (std::isnan(*x0) || std::isnan(*x1)) return false;
    return true;
}

Transform Translate(const V3 &d) {   // core/transform.cpp:144-154
    return Transform(Matrix4x4(1, 0, 0, d.x, 0, 1, 0, d.y, 0, 0, 1, d.z, 0, 0, 0, 1),
                     Matrix4x4(1, 0, 0, -d.x, 0, 1, 0, -d.y, 0, 0, 1, -d.z, 0, 0, 0, 1));
}

Transform Scale(float x, float y, float z) {   // core/transform.cpp:157-167
    return Transform(Matrix4x4(x, 0, 0, 0, 0, y, 0, 0, 0, 0, z, 0, 0, 0, 0, 1),
                     Matrix4x4(1.f / x, 0, 0, 0, 0, 1.f / y, 0, 0, 0, 0, 1.f / z, 0, 0, 0, 0, 1));
}

Transform Rotate(float angle, const V3 &axis) {   // core/transform.cpp:206-232
    V3 a = Normalize(axis);
    float s = sinf(Radians(angle));
    float c = cosf(Radians(angle));
    Matrix4x4 r;
    r.m[0][0] = a.x * a.x + (1.f - a.x * a.x) * c;
    r.m[0][1] = a.x * a.y * (1.f - c) - a.z * s;
    r.m[0][2] = a.x * a.z * (1.f - c) + a.y * s;
    r.m[0][3] = 0;
    r.m[1][0] = a.x * a.y * (1.f - c) + a.z * s;
    r.m[1][1] = a.y * a.y + (1.f - a.y * a.y) * c;
    r.m[1][2] = a.y * a.z * (1.f - c) - a.x * s;
    r.m[1][3] = 0;
    r.m[2][0] = a.x * a.z * (1.f - c) - a.y * s;
    r.m[2][1] = a.y * a.z * (1.f - c) + a.x * s;
    r.m[2][2] = a.z * a.z + (1.f - a.z * a.z) * c;
    r.m[2][3] = 0;
    r.m[3][0] = 0; r.m[3][1] = 0; r.m[3][2] = 0; r.m[3][3] = 1;
    return Transform(r, Transpose(r));
}

Transform LookAt(const V3 &pos, const V3 &look, const V3 &up, bool *ok) {   // transform.cpp:235-270
    V3 dir = Normalize(look - pos);
    if (Length(Cross(Normalize(up), dir)) == 0) {
        if (ok) *ok = false;
        return Transform();
    }
    if (ok) *ok = true;
    V3 left = Normalize(Cross(Normalize(up), dir));
    V3 newUp = Cross(dir, left);
    Matrix4x4 c2w(left.x, newUp.x, dir.x, pos.x,
                  left.y, newUp.y, dir.y, pos.y,
                  left.z, newUp.z, dir.z, pos.z,
                  0, 0, 0, 1);
    return Transform(Inverse(c2w), c2w);
}

Transform Perspective(float fov, float n, float f) {   // core/transform.cpp:324-334
    Matrix4x4 persp(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, f / (f - n), -f * n / (f - n), 0, 0, 1, 0);
    float invTanAng = 1.f / tanf(Radians(fov) / 2.f);
    return Scale(invTanAng, invTanAng, 1) * Transform(persp);
}

BBox Transform::Box(const BBox &b) const {   // core/transform.cpp:283-294 (same corner order)
    BBox r(Pt(V3(b.pMin.x, b.pMin.y, b.pMin.z)));
    r = Union(r, Pt(V3(b.pMax.x, b.pMin.y, b.pMin.z)));
    r = Union(r, Pt(V3(b.pMin.x, b.pMax.y, b.pMin.z)));
    r = Union(r, Pt(V3(b.pMin.x, b.pMin.y, b.pMax.z)));
    r = Union(r, Pt(V3(b.pMin.x, b.pMax.y, b.pMax.z)));
    r = Union(r, Pt(V3(b.pMax.x, b.pMax.y, b.pMin.z)));
    r = Union(r, Pt(V3(b.pMax.x, b.pMin.y, b.pMax.z)));
    r = Union(r, Pt(V3(b.pMax.x, b.pMax.y, b.pMax.z)));
    return r;
}

bool Transform::SwapsHandedness() const {   // core/transform.cpp:304-315
    float det = ((m.m[0][0] * (m.m[1][1] * m.m[2][2] - m.m[1][2] * m.m[2][1])) -
                 (m.m[0][1] * (m.m[1][0] * m.m[2][2] - m.m[1][2] * m.m[2][0])) +
                 (m.m[0][2] * (m.m[1][0] * m.m[2][1] - m.m[1][1] * m.m[2][0])));
    return det < 0.f;
}

bool Transform::HasScale() const {   // core/transform.h:134-141
    float la2 = LengthSquared(Vec(V3(1, 0, 0)));
    float lb2 = LengthSquared(Vec(V3(0, 1, 0)));
    float lc2 = LengthSquared(Vec(V3(0, 0, 1)));
    auto notOne = [](float x) { return x < .999f || x > 1.001f; };
    return notOne(la2) || notOne(lb2) || notOne(lc2);
}

// ---------------------------------------------------------------- Quaternion --
Transform Quaternion::ToTransform() const {   // quaternion.cpp:39-58
    float xx = v.x * v.x, yy = v.y * v.y, zz = v.z * v.z;
    float xy = v.x * v.y, xz = v.x * v.z, yz = v.y * v.z;
    float wx = v.x * w, wy = v.y * w, wz = v.z * w;
    Matrix4x4 m;
    m.m[0][0] = 1.f - 2.f * (yy + zz);
    m.m[0][1] = 2.f * (xy + wz);
    m.m[0][2] = 2.f * (xz - wy);
    m.m[1][0] = 2.f * (xy - wz);
    m.m[1][1] = 1.f - 2.f * (xx + zz);
    m.m[1][2] = 2.f * (yz + wx);
    m.m[2][0] = 2.f * (xz + wy);
    m.m[2][1] = 2.f * (yz - wx);
    m.m[2][2] = 1.f - 2.f * (xx + yy);
    return Transform(Transpose(m), m);   // "transpose since we are left-handed"
}

Quaternion::Quaternion(const Matrix4x4 &m) {   // quaternion.cpp:61-96 (double-typed "+ 1.0" kept)
    float trace = m.m[0][0] + m.m[1][1] + m.m[2][2];
    if (trace > 0.f) {
        float s = sqrtf((float)(trace + 1.0));
        w = s / 2.0f;
        s = 0.5f / s;
        v.x = (m.m[2][1] - m.m[1][2]) * s;
        v.y = (m.m[0][2] - m.m[2][0]) * s;
        v.z = (m.m[1][0] - m.m[0][1]) * s;
    } else {
        const int nxt[3] = {1, 2, 0};
        float q[3];
        int i = 0;
        if (m.m[1][1] > m.m[0][0]) i = 1;
        if (m.m[2][2] > m.m[i][i]) i = 2;
        int j = nxt[i];
        int k = nxt[j];
        float s = sqrtf((float)((m.m[i][i] - (m.m[j][j] + m.m[k][k])) + 1.0));
        q[i] = s * 0.5f;
        if (s != 0.f) s = 0.5f / s;
        w = (m.m[k][j] - m.m[j][k]) * s;
        q[j] = (m.m[j][i] + m.m[i][j]) * s;
        q[k] = (m.m[k][i] + m.m[i][k]) * s;
        v = V3(q[0], q[1], q[2]);
    }
}

static inline float QDot(const Quaternion &a, const Quaternion &b) { return Dot(a.v, b.v) + a.w * b.w; }
static inline Quaternion QMul(const Quaternion &q, float f) { return Quaternion(V3(q.v.x * f, q.v.y * f, q.v.z * f), q.w * f); }
static inline Quaternion QAdd(const Quaternion &a, const Quaternion &b) { return Quaternion(a.v + b.v, a.w + b.w); }
static inline Quaternion QSub(const Quaternion &a, const Quaternion &b) { return Quaternion(a.v - b.v, a.w - b.w); }
static inline Quaternion QNormalize(const Quaternion &q) {   // quaternion.h:109-111: v by reciprocal, w by division
    float f = sqrtf(QDot(q, q));
    return Quaternion(q.v / f, q.w / f);
}

Quaternion Slerp(float t, const Quaternion &q1, const Quaternion &q2) {   // quaternion.cpp:99-111
    float cosTheta = QDot(q1, q2);
    if (cosTheta > .9995f) return QNormalize(QAdd(QMul(q1, 1.f - t), QMul(q2, t)));
    float theta = acosf(Clampf(cosTheta, -1.f, 1.f));
    float thetap = theta * t;
    Quaternion qperp = QNormalize(QSub(q2, QMul(q1, cosTheta)));
    return QAdd(QMul(q1, cosf(thetap)), QMul(qperp, sinf(thetap)));
}

// --------------------------------------------------------- AnimatedTransform --
AnimatedTransform::AnimatedTransform(const Transform &t1, float time1, const Transform &t2, float time2)
    : startTime(time1), endTime(time2), startTransform(t1), endTransform(t2), actuallyAnimated(t1 != t2) {
    Decompose(startTransform.m, &T[0], &R[0], &S[0]);
    Decompose(endTransform.m, &T[1], &R[1], &S[1]);
}

void AnimatedTransform::Decompose(const Matrix4x4 &m, V3 *T, Quaternion *Rquat, Matrix4x4 *S) {   // transform.cpp:328-368
    T->x = m.m[0][3];
    T->y = m.m[1][3];
    T->z = m.m[2][3];
    Matrix4x4 M = m;
    for (int i = 0; i < 3; ++i) M.m[i][3] = M.m[3][i] = 0.f;
    M.m[3][3] = 1.f;
    float norm;
    int count = 0;
    Matrix4x4 R = M;
    do {   // polar decomposition by averaging with the inverse transpose
        Matrix4x4 Rnext;
        Matrix4x4 Rit = Inverse(Transpose(R));
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) Rnext.m[i][j] = 0.5f * (R.m[i][j] + Rit.m[i][j]);
        norm = 0.f;
        for (int i = 0; i < 3; ++i) {
            float n = fabsf(R.m[i][0] - Rnext.m[i][0]) + fabsf(R.m[i][1] - Rnext.m[i][1]) + fabsf(R.m[i][2] - Rnext.m[i][2]);
            norm = std::max(norm, n);
        }
        R = Rnext;
    } while (++count < 100 && norm > .0001f);
    *Rquat = Quaternion(R);
    *S = Matrix4x4::Mul(Inverse(R), M);
}

void AnimatedTransform::Interpolate(float time, Transform *t) const {   // transform.cpp:371-396
    if (!actuallyAnimated || time <= startTime) { *t = startTransform; return; }
    if (time >= endTime) { *t = endTransform; return; }
    float dt = (time - startTime) / (endTime - startTime);
    V3 trans = (1.f - dt) * T[0] + dt * T[1];
    Quaternion rotate = Slerp(dt, R[0], R[1]);
    Matrix4x4 scale;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) scale.m[i][j] = Lerp(dt, S[0].m[i][j], S[1].m[i][j]);
    *t = Translate(trans) * rotate.ToTransform() * Transform(scale);
}

BBox AnimatedTransform::MotionBounds(const BBox &b, bool useInverse) const {   // transform.cpp:399-413
    if (!actuallyAnimated) return useInverse ? Inverse(startTransform).Box(b) : startTransform.Box(b);
    BBox ret;
    const int nSteps = 128;
    for (int i = 0; i < nSteps; ++i) {
        Transform t;
        float time = Lerp(float(i) / float(nSteps - 1), startTime, endTime);
        Interpolate(time, &t);
        if (useInverse) t = Inverse(t);
        ret = Union(ret, t.Box(b));
    }
    return ret;
}

}  // namespace pb
