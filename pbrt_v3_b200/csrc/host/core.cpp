// Host transform / matrix routines.  Arithmetic order follows src/core/transform.cpp so that the
// camera matrices and world-space vertices handed to the GPU carry the reference's bits.
#include "core.h"

#include <cstdarg>

namespace pbrt {

int g_errorCount = 0;

void Warning(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::fprintf(stderr, "Warning: ");
    std::vfprintf(stderr, fmt, ap);
    std::fprintf(stderr, "\n");
    va_end(ap);
}
void Error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::fprintf(stderr, "Error: ");
    std::vfprintf(stderr, fmt, ap);
    std::fprintf(stderr, "\n");
    va_end(ap);
    ++g_errorCount;
}

Matrix4x4 Transpose(const Matrix4x4 &m) {
    Matrix4x4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = m.m[j][i];
    return r;
}

// transform.cpp:82-141.  Gauss-Jordan elimination with full pivoting; the pivot search order and
// the ">=" comparison decide which (equal-magnitude) pivot wins, so they are kept as-is.
Matrix4x4 Inverse(const Matrix4x4 &src) {
    int colIdx[4], rowIdx[4];
    int pivoted[4] = {0, 0, 0, 0};
    Float a[4][4];
    std::memcpy(a, src.m, sizeof(a));
    for (int step = 0; step < 4; ++step) {
        int prow = 0, pcol = 0;
        Float best = 0.f;
        for (int j = 0; j < 4; ++j) {
            if (pivoted[j] == 1) continue;
            for (int k = 0; k < 4; ++k) {
                if (pivoted[k] == 0) {
                    if (std::abs(a[j][k]) >= best) {
                        best = Float(std::abs(a[j][k]));
                        prow = j;
                        pcol = k;
                    }
                } else if (pivoted[k] > 1)
                    Error("Singular matrix in MatrixInvert");
            }
        }
        ++pivoted[pcol];
        if (prow != pcol)
            for (int k = 0; k < 4; ++k) std::swap(a[prow][k], a[pcol][k]);
        rowIdx[step] = prow;
        colIdx[step] = pcol;
        if (a[pcol][pcol] == 0.f) Error("Singular matrix in MatrixInvert");
        Float pivinv = 1. / a[pcol][pcol];  // double divide, rounded to float (transform.cpp:117)
        a[pcol][pcol] = 1.;
        for (int j = 0; j < 4; ++j) a[pcol][j] *= pivinv;
        for (int j = 0; j < 4; ++j) {
            if (j == pcol) continue;
            Float save = a[j][pcol];
            a[j][pcol] = 0;
            for (int k = 0; k < 4; ++k) a[j][k] -= a[pcol][k] * save;
        }
    }
    for (int j = 3; j >= 0; --j)
        if (rowIdx[j] != colIdx[j])
            for (int k = 0; k < 4; ++k) std::swap(a[k][rowIdx[j]], a[k][colIdx[j]]);
    Matrix4x4 r;
    std::memcpy(r.m, a, sizeof(a));
    return r;
}

Bounds3f Transform::operator()(const Bounds3f &b) const {
    const Transform &M = *this;
    Bounds3f ret(M(Point3f(b.pMin.x, b.pMin.y, b.pMin.z)));
    ret = Union(ret, M(Point3f(b.pMax.x, b.pMin.y, b.pMin.z)));
    ret = Union(ret, M(Point3f(b.pMin.x, b.pMax.y, b.pMin.z)));
    ret = Union(ret, M(Point3f(b.pMin.x, b.pMin.y, b.pMax.z)));
    ret = Union(ret, M(Point3f(b.pMin.x, b.pMax.y, b.pMax.z)));
    ret = Union(ret, M(Point3f(b.pMax.x, b.pMax.y, b.pMin.z)));
    ret = Union(ret, M(Point3f(b.pMax.x, b.pMin.y, b.pMax.z)));
    ret = Union(ret, M(Point3f(b.pMax.x, b.pMax.y, b.pMax.z)));
    return ret;
}

bool Transform::SwapsHandedness() const {
    Float det = m.m[0][0] * (m.m[1][1] * m.m[2][2] - m.m[1][2] * m.m[2][1]) -
                m.m[0][1] * (m.m[1][0] * m.m[2][2] - m.m[1][2] * m.m[2][0]) +
                m.m[0][2] * (m.m[1][0] * m.m[2][1] - m.m[1][1] * m.m[2][0]);
    return det < 0;
}

bool Transform::HasScale() const {
    Float la2 = ApplyVector(Vector3f(1, 0, 0)).LengthSquared();
    Float lb2 = ApplyVector(Vector3f(0, 1, 0)).LengthSquared();
    Float lc2 = ApplyVector(Vector3f(0, 0, 1)).LengthSquared();
    auto notOne = [](Float x) { return x < .999f || x > 1.001f; };
    return notOne(la2) || notOne(lb2) || notOne(lc2);
}

Transform Translate(const Vector3f &d) {
    return Transform(Matrix4x4(1, 0, 0, d.x, 0, 1, 0, d.y, 0, 0, 1, d.z, 0, 0, 0, 1),
                     Matrix4x4(1, 0, 0, -d.x, 0, 1, 0, -d.y, 0, 0, 1, -d.z, 0, 0, 0, 1));
}

Transform Scale(Float x, Float y, Float z) {
    return Transform(Matrix4x4(x, 0, 0, 0, 0, y, 0, 0, 0, 0, z, 0, 0, 0, 0, 1),
                     Matrix4x4(1 / x, 0, 0, 0, 0, 1 / y, 0, 0, 0, 0, 1 / z, 0, 0, 0, 0, 1));
}

// transform.cpp:197-218
Transform Rotate(Float theta, const Vector3f &axis) {
    Vector3f a = Normalize(axis);
    Float s = std::sin(Radians(theta));
    Float c = std::cos(Radians(theta));
    Matrix4x4 m;
    m.m[0][0] = a.x * a.x + (1 - a.x * a.x) * c;
    m.m[0][1] = a.x * a.y * (1 - c) - a.z * s;
    m.m[0][2] = a.x * a.z * (1 - c) + a.y * s;
    m.m[0][3] = 0;
    m.m[1][0] = a.x * a.y * (1 - c) + a.z * s;
    m.m[1][1] = a.y * a.y + (1 - a.y * a.y) * c;
    m.m[1][2] = a.y * a.z * (1 - c) - a.x * s;
    m.m[1][3] = 0;
    m.m[2][0] = a.x * a.z * (1 - c) - a.y * s;
    m.m[2][1] = a.y * a.z * (1 - c) + a.x * s;
    m.m[2][2] = a.z * a.z + (1 - a.z * a.z) * c;
    m.m[2][3] = 0;
    return Transform(m, Transpose(m));
}

// transform.cpp:220-256
Transform LookAt(const Point3f &pos, const Point3f &look, const Vector3f &up) {
    Matrix4x4 c2w;
    c2w.m[0][3] = pos.x;
    c2w.m[1][3] = pos.y;
    c2w.m[2][3] = pos.z;
    c2w.m[3][3] = 1;
    Vector3f dir = Normalize(look - pos);
    if (Cross(Normalize(up), dir).Length() == 0) {
        Error("\"up\" vector (%f, %f, %f) and viewing direction (%f, %f, %f) passed to LookAt are "
              "pointing in the same direction.  Using the identity transformation.",
              up.x, up.y, up.z, dir.x, dir.y, dir.z);
        return Transform();
    }
    Vector3f right = Normalize(Cross(Normalize(up), dir));
    Vector3f newUp = Cross(dir, right);
    c2w.m[0][0] = right.x; c2w.m[1][0] = right.y; c2w.m[2][0] = right.z; c2w.m[3][0] = 0.;
    c2w.m[0][1] = newUp.x; c2w.m[1][1] = newUp.y; c2w.m[2][1] = newUp.z; c2w.m[3][1] = 0.;
    c2w.m[0][2] = dir.x;   c2w.m[1][2] = dir.y;   c2w.m[2][2] = dir.z;   c2w.m[3][2] = 0.;
    return Transform(Inverse(c2w), c2w);
}

// transform.cpp:303-311
Transform Perspective(Float fov, Float n, Float f) {
    Matrix4x4 persp(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, f / (f - n), -f * n / (f - n), 0, 0, 1, 0);
    Float invTanAng = 1 / std::tan(Radians(fov) / 2);
    return Scale(invTanAng, invTanAng, 1) * Transform(persp);
}

}  // namespace pbrt
