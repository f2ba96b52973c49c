// Host-side math used by the scene front end (parser, transforms, BVH build, film).
// Everything here runs once per scene on the CPU; the per-ray / per-sample hot path lives in
// ../device and runs on the GPU only.
//
// The arithmetic (operation order, float vs double, rounding helpers) follows the reference so
// that the flattened scene the GPU sees holds the same bits the reference would compute:
//   constants / gamma / NextFloatUp|Down   src/core/pbrt.h:196-291
//   Vector/Point/Normal/Bounds             src/core/geometry.h
//   Matrix4x4 / Transform                  src/core/transform.{h,cpp}
//   RNG (PCG32)                            src/core/rng.h:64-144
#ifndef PB2_HOST_CORE_H
#define PB2_HOST_CORE_H

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

namespace pbrt {

typedef float Float;

static constexpr Float Infinity = std::numeric_limits<Float>::infinity();
static constexpr Float MachineEpsilon = std::numeric_limits<Float>::epsilon() * 0.5;
static constexpr Float Pi = 3.14159265358979323846;

inline Float gamma(int n) { return (n * MachineEpsilon) / (1 - n * MachineEpsilon); }
inline Float Radians(Float deg) { return (Pi / 180) * deg; }
template <typename T, typename U, typename V>
inline T Clamp(T val, U low, V high) {
    if (val < low) return low;
    if (val > high) return high;
    return val;
}
template <typename T>
inline T Mod(T a, T b) {
    T result = a - (a / b) * b;
    return (T)((result < 0) ? result + b : result);
}

// ---------------------------------------------------------------- vectors
struct Vector3f {
    Float x = 0, y = 0, z = 0;
    Vector3f() {}
    Vector3f(Float x, Float y, Float z) : x(x), y(y), z(z) {}
    Float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    Float &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    Vector3f operator+(const Vector3f &v) const { return Vector3f(x + v.x, y + v.y, z + v.z); }
    Vector3f operator-(const Vector3f &v) const { return Vector3f(x - v.x, y - v.y, z - v.z); }
    Vector3f operator-() const { return Vector3f(-x, -y, -z); }
    Vector3f operator*(Float s) const { return Vector3f(s * x, s * y, s * z); }
    Vector3f operator/(Float f) const {
        Float inv = (Float)1 / f;  // geometry.h:244-248: multiply by the reciprocal
        return Vector3f(x * inv, y * inv, z * inv);
    }
    Float LengthSquared() const { return x * x + y * y + z * z; }
    Float Length() const { return std::sqrt(LengthSquared()); }
    bool operator==(const Vector3f &v) const { return x == v.x && y == v.y && z == v.z; }
};
typedef Vector3f Point3f;   // the host code never needs the point/vector distinction in types
typedef Vector3f Normal3f;

struct Point2f {
    Float x = 0, y = 0;
    Point2f() {}
    Point2f(Float x, Float y) : x(x), y(y) {}
};
struct Point2i {
    int x = 0, y = 0;
    Point2i() {}
    Point2i(int x, int y) : x(x), y(y) {}
    int operator[](int i) const { return i == 0 ? x : y; }
};
struct Bounds2i {
    Point2i pMin, pMax;
    Bounds2i() {}
    Bounds2i(Point2i a, Point2i b) : pMin(a), pMax(b) {}
    int Area() const { return (pMax.x - pMin.x) * (pMax.y - pMin.y); }
};
struct Bounds2f {
    Point2f pMin, pMax;
};

inline Vector3f operator*(Float s, const Vector3f &v) { return v * s; }
inline Float Dot(const Vector3f &a, const Vector3f &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vector3f Abs(const Vector3f &v) { return Vector3f(std::abs(v.x), std::abs(v.y), std::abs(v.z)); }
// geometry.h:957-963: cross products are evaluated in double and rounded once.
inline Vector3f Cross(const Vector3f &v1, const Vector3f &v2) {
    double v1x = v1.x, v1y = v1.y, v1z = v1.z;
    double v2x = v2.x, v2y = v2.y, v2z = v2.z;
    return Vector3f((Float)((v1y * v2z) - (v1z * v2y)), (Float)((v1z * v2x) - (v1x * v2z)),
                    (Float)((v1x * v2y) - (v1y * v2x)));
}
inline Vector3f Normalize(const Vector3f &v) { return v / v.Length(); }
// geometry.h:1020-1028
inline void CoordinateSystem(const Vector3f &v1, Vector3f *v2, Vector3f *v3) {
    if (std::abs(v1.x) > std::abs(v1.y))
        *v2 = Vector3f(-v1.z, 0, v1.x) / std::sqrt(v1.x * v1.x + v1.z * v1.z);
    else
        *v2 = Vector3f(0, v1.z, -v1.y) / std::sqrt(v1.y * v1.y + v1.z * v1.z);
    *v3 = Cross(v1, *v2);
}
inline Float Distance(const Point3f &a, const Point3f &b) { return (a - b).Length(); }

struct Bounds3f {
    Point3f pMin, pMax;
    Bounds3f() {
        Float minNum = std::numeric_limits<Float>::lowest();
        Float maxNum = std::numeric_limits<Float>::max();
        pMin = Point3f(maxNum, maxNum, maxNum);
        pMax = Point3f(minNum, minNum, minNum);
    }
    explicit Bounds3f(const Point3f &p) : pMin(p), pMax(p) {}
    Bounds3f(const Point3f &p1, const Point3f &p2)
        : pMin(std::min(p1.x, p2.x), std::min(p1.y, p2.y), std::min(p1.z, p2.z)),
          pMax(std::max(p1.x, p2.x), std::max(p1.y, p2.y), std::max(p1.z, p2.z)) {}
    Vector3f Diagonal() const { return pMax - pMin; }
    Float SurfaceArea() const {
        Vector3f d = Diagonal();
        return 2 * (d.x * d.y + d.x * d.z + d.y * d.z);
    }
    int MaximumExtent() const {
        Vector3f d = Diagonal();
        if (d.x > d.y && d.x > d.z) return 0;
        else if (d.y > d.z) return 1;
        else return 2;
    }
    Vector3f Offset(const Point3f &p) const {
        Vector3f o = p - pMin;
        if (pMax.x > pMin.x) o.x /= pMax.x - pMin.x;
        if (pMax.y > pMin.y) o.y /= pMax.y - pMin.y;
        if (pMax.z > pMin.z) o.z /= pMax.z - pMin.z;
        return o;
    }
};
inline Bounds3f Union(const Bounds3f &b, const Point3f &p) {
    Bounds3f r;
    r.pMin = Point3f(std::min(b.pMin.x, p.x), std::min(b.pMin.y, p.y), std::min(b.pMin.z, p.z));
    r.pMax = Point3f(std::max(b.pMax.x, p.x), std::max(b.pMax.y, p.y), std::max(b.pMax.z, p.z));
    return r;
}
inline Bounds3f Union(const Bounds3f &a, const Bounds3f &b) {
    Bounds3f r;
    r.pMin = Point3f(std::min(a.pMin.x, b.pMin.x), std::min(a.pMin.y, b.pMin.y), std::min(a.pMin.z, b.pMin.z));
    r.pMax = Point3f(std::max(a.pMax.x, b.pMax.x), std::max(a.pMax.y, b.pMax.y), std::max(a.pMax.z, b.pMax.z));
    return r;
}

// ---------------------------------------------------------------- matrices / transforms
struct Matrix4x4 {
    Float m[4][4];
    Matrix4x4() {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) m[i][j] = (i == j) ? 1.f : 0.f;
    }
    Matrix4x4(Float t00, Float t01, Float t02, Float t03, Float t10, Float t11, Float t12, Float t13,
              Float t20, Float t21, Float t22, Float t23, Float t30, Float t31, Float t32, Float t33) {
        Float v[16] = {t00, t01, t02, t03, t10, t11, t12, t13, t20, t21, t22, t23, t30, t31, t32, t33};
        std::memcpy(m, v, sizeof(v));
    }
    bool operator==(const Matrix4x4 &o) const { return std::memcmp(m, o.m, sizeof(m)) == 0; }
    bool operator<(const Matrix4x4 &o) const { return std::memcmp(m, o.m, sizeof(m)) < 0; }
    // transform.h:84-94
    static Matrix4x4 Mul(const Matrix4x4 &m1, const Matrix4x4 &m2) {
        Matrix4x4 r;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                r.m[i][j] = m1.m[i][0] * m2.m[0][j] + m1.m[i][1] * m2.m[1][j] +
                            m1.m[i][2] * m2.m[2][j] + m1.m[i][3] * m2.m[3][j];
        return r;
    }
};
Matrix4x4 Transpose(const Matrix4x4 &m);
Matrix4x4 Inverse(const Matrix4x4 &m);  // Gauss-Jordan with full pivoting, transform.cpp:82-141

class Transform {
  public:
    Transform() {}
    explicit Transform(const Matrix4x4 &m) : m(m), mInv(Inverse(m)) {}
    Transform(const Matrix4x4 &m, const Matrix4x4 &mInv) : m(m), mInv(mInv) {}
    const Matrix4x4 &GetMatrix() const { return m; }
    const Matrix4x4 &GetInverseMatrix() const { return mInv; }
    bool operator==(const Transform &t) const { return t.m == m && t.mInv == mInv; }
    bool operator<(const Transform &t) const { return m < t.m; }
    bool IsIdentity() const { return m == Matrix4x4(); }
    // transform.h:219-231 (divide by w unless it is exactly 1)
    Point3f operator()(const Point3f &p) const {
        Float x = p.x, y = p.y, z = p.z;
        Float xp = m.m[0][0] * x + m.m[0][1] * y + m.m[0][2] * z + m.m[0][3];
        Float yp = m.m[1][0] * x + m.m[1][1] * y + m.m[1][2] * z + m.m[1][3];
        Float zp = m.m[2][0] * x + m.m[2][1] * y + m.m[2][2] * z + m.m[2][3];
        Float wp = m.m[3][0] * x + m.m[3][1] * y + m.m[3][2] * z + m.m[3][3];
        if (wp == 1) return Point3f(xp, yp, zp);
        return Point3f(xp, yp, zp) / wp;
    }
    // transform.h:233-239
    Vector3f ApplyVector(const Vector3f &v) const {
        Float x = v.x, y = v.y, z = v.z;
        return Vector3f(m.m[0][0] * x + m.m[0][1] * y + m.m[0][2] * z,
                        m.m[1][0] * x + m.m[1][1] * y + m.m[1][2] * z,
                        m.m[2][0] * x + m.m[2][1] * y + m.m[2][2] * z);
    }
    // transform.h:241-249 (inverse transpose, not renormalised)
    Normal3f ApplyNormal(const Normal3f &n) const {
        Float x = n.x, y = n.y, z = n.z;
        return Normal3f(mInv.m[0][0] * x + mInv.m[1][0] * y + mInv.m[2][0] * z,
                        mInv.m[0][1] * x + mInv.m[1][1] * y + mInv.m[2][1] * z,
                        mInv.m[0][2] * x + mInv.m[1][2] * y + mInv.m[2][2] * z);
    }
    Bounds3f operator()(const Bounds3f &b) const;
    Transform operator*(const Transform &t2) const {
        return Transform(Matrix4x4::Mul(m, t2.m), Matrix4x4::Mul(t2.mInv, mInv));
    }
    bool SwapsHandedness() const;
    bool HasScale() const;

  private:
    Matrix4x4 m, mInv;
};
inline Transform Inverse(const Transform &t) { return Transform(t.GetInverseMatrix(), t.GetMatrix()); }
Transform Translate(const Vector3f &delta);
Transform Scale(Float x, Float y, Float z);
Transform Rotate(Float theta, const Vector3f &axis);
Transform LookAt(const Point3f &pos, const Point3f &look, const Vector3f &up);
Transform Perspective(Float fov, Float znear, Float zfar);

// ---------------------------------------------------------------- PCG32 (rng.h:64-144)
class RNG {
  public:
    RNG() : state(0x853c49e6748fea9bULL), inc(0xda3e39cb94b95bdbULL) {}
    explicit RNG(uint64_t sequenceIndex) { SetSequence(sequenceIndex); }
    void SetSequence(uint64_t initseq) {
        state = 0u;
        inc = (initseq << 1u) | 1u;
        UniformUInt32();
        state += 0x853c49e6748fea9bULL;
        UniformUInt32();
    }
    uint32_t UniformUInt32() {
        uint64_t oldstate = state;
        state = oldstate * 0x5851f42d4c957f2dULL + inc;
        uint32_t xorshifted = (uint32_t)(((oldstate >> 18u) ^ oldstate) >> 27u);
        uint32_t rot = (uint32_t)(oldstate >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    uint32_t UniformUInt32(uint32_t b) {
        uint32_t threshold = (~b + 1u) % b;
        while (true) {
            uint32_t r = UniformUInt32();
            if (r >= threshold) return r % b;
        }
    }
    Float UniformFloat() { return std::min(0x1.fffffep-1f, Float(UniformUInt32() * 0x1p-32f)); }

  private:
    uint64_t state, inc;
};

// error.cpp:62-102: user errors print and continue.
void Warning(const char *fmt, ...);
void Error(const char *fmt, ...);
extern int g_errorCount;

}  // namespace pbrt
#endif
