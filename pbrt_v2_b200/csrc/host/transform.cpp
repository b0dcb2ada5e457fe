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
    int pivotRow[4], pivotCol[4];
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

}  // namespace pb
