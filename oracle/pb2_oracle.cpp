// TEST INFRASTRUCTURE ONLY — never linked into, imported by or called from the product
// (pbrt_v3_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may load this library, and only as the checker or the timed CPU baseline.
//
// Plain C++ (CPU, scalar, object-per-concept like the reference) restatement of the reference's
// path-tracing hot path, every function citing the reference lines it follows.  It is compiled
// with the reference's own numeric environment (g++ -O2, x86-64 baseline, no FMA contraction), so
// that on the same inputs it is expected to reproduce the reference BIT FOR BIT; tests/ pin it
// against the reference's known-answer tests (src/tests/shapes.cpp, sampling.cpp) and, in the
// build container, against oracle/_ref (the reference's own sources) on rays, sampler values,
// light distributions, per-sample radiance and whole images.  PARITY PINNED: see DESIGN.md §Oracle.
//
// Structure mirrors the reference (Shape / Primitive / BVHAccel / BxDF / BSDF / Light / Sampler /
// PathIntegrator objects with virtual calls), deliberately unlike the CUDA implementation
// (flat records, per-lane state machine), so that the two are independent statements of the path.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "pb2.h"

namespace orc {

typedef float Float;
static const Float Infinity = std::numeric_limits<Float>::infinity();
static const Float MachineEpsilon = std::numeric_limits<Float>::epsilon() * 0.5;  // pbrt.h:201
static const Float ShadowEpsilon = 0.0001f;
static const Float Pi = 3.14159265358979323846;
static const Float InvPi = 0.31830988618379067154;
static const Float PiOver2 = 1.57079632679489661923;
static const Float PiOver4 = 0.78539816339744830961;
static const Float OneMinusEpsilon = 0x1.fffffep-1;

inline Float gamma(int n) { return (n * MachineEpsilon) / (1 - n * MachineEpsilon); }  // pbrt.h:289
inline Float Radians(Float deg) { return (Pi / 180) * deg; }
template <typename T, typename U, typename V>
inline T Clamp(T val, U low, V high) {
    if (val < low) return low;
    else if (val > high) return high;
    else return val;
}
inline Float Lerp(Float t, Float v1, Float v2) { return (1 - t) * v1 + t * v2; }
inline uint32_t FloatToBits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float BitsToFloat(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
// pbrt.h:241-265
inline float NextFloatUp(float v) {
    if (std::isinf(v) && v > 0.) return v;
    if (v == -0.f) v = 0.f;
    uint32_t ui = FloatToBits(v);
    if (v >= 0) ++ui; else --ui;
    return BitsToFloat(ui);
}
inline float NextFloatDown(float v) {
    if (std::isinf(v) && v < 0.) return v;
    if (v == 0.f) v = -0.f;
    uint32_t ui = FloatToBits(v);
    if (v > 0) --ui; else ++ui;
    return BitsToFloat(ui);
}

// ------------------------------------------------------------------ geometry (src/core/geometry.h)
struct Vec {
    Float x, y, z;
    Vec() : x(0), y(0), z(0) {}
    Vec(Float x, Float y, Float z) : x(x), y(y), z(z) {}
    Float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    Float &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    Vec operator+(const Vec &v) const { return Vec(x + v.x, y + v.y, z + v.z); }
    Vec operator-(const Vec &v) const { return Vec(x - v.x, y - v.y, z - v.z); }
    Vec operator-() const { return Vec(-x, -y, -z); }
    Vec operator*(Float s) const { return Vec(s * x, s * y, s * z); }
    Vec operator/(Float f) const { Float inv = (Float)1 / f; return Vec(x * inv, y * inv, z * inv); }
    Float LengthSquared() const { return x * x + y * y + z * z; }
    Float Length() const { return std::sqrt(LengthSquared()); }
    bool IsZero() const { return x == 0 && y == 0 && z == 0; }
};
inline Vec operator*(Float s, const Vec &v) { return v * s; }
inline Float Dot(const Vec &a, const Vec &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Float AbsDot(const Vec &a, const Vec &b) { return std::abs(Dot(a, b)); }
inline Vec Abs(const Vec &v) { return Vec(std::abs(v.x), std::abs(v.y), std::abs(v.z)); }
inline Vec Cross(const Vec &v1, const Vec &v2) {  // geometry.h:957-963
    double v1x = v1.x, v1y = v1.y, v1z = v1.z, v2x = v2.x, v2y = v2.y, v2z = v2.z;
    return Vec((Float)((v1y * v2z) - (v1z * v2y)), (Float)((v1z * v2x) - (v1x * v2z)), (Float)((v1x * v2y) - (v1y * v2x)));
}
inline Vec Normalize(const Vec &v) { return v / v.Length(); }
inline Float MaxComponent(const Vec &v) { return std::max(v.x, std::max(v.y, v.z)); }
inline int MaxDimension(const Vec &v) { return (v.x > v.y) ? ((v.x > v.z) ? 0 : 2) : ((v.y > v.z) ? 1 : 2); }
inline Vec Permute(const Vec &v, int x, int y, int z) { return Vec(v[x], v[y], v[z]); }
inline Vec Faceforward(const Vec &n, const Vec &v) { return (Dot(n, v) < 0.f) ? -n : n; }
inline Float DistanceSquared(const Vec &a, const Vec &b) { return (a - b).LengthSquared(); }
inline Float Distance(const Vec &a, const Vec &b) { return (a - b).Length(); }
inline void CoordinateSystem(const Vec &v1, Vec *v2, Vec *v3) {  // geometry.h:1020-1027
    if (std::abs(v1.x) > std::abs(v1.y)) *v2 = Vec(-v1.z, 0, v1.x) / std::sqrt(v1.x * v1.x + v1.z * v1.z);
    else *v2 = Vec(0, v1.z, -v1.y) / std::sqrt(v1.y * v1.y + v1.z * v1.z);
    *v3 = Cross(v1, *v2);
}
struct Vec2 {
    Float x, y;
    Vec2() : x(0), y(0) {}
    Vec2(Float x, Float y) : x(x), y(y) {}
    Float operator[](int i) const { return i == 0 ? x : y; }
};
// geometry.h:1440-1454
inline Vec OffsetRayOrigin(const Vec &p, const Vec &pError, const Vec &n, const Vec &w) {
    Float d = Dot(Abs(n), pError);
    Vec offset = d * n;
    if (Dot(w, n) < 0) offset = -offset;
    Vec po = p + offset;
    for (int i = 0; i < 3; ++i) {
        if (offset[i] > 0) po[i] = NextFloatUp(po[i]);
        else if (offset[i] < 0) po[i] = NextFloatDown(po[i]);
    }
    return po;
}

struct Ray {
    Vec o, d;
    mutable Float tMax;
    Ray() : tMax(Infinity) {}
    Ray(const Vec &o, const Vec &d, Float tMax = Infinity) : o(o), d(d), tMax(tMax) {}
    Vec operator()(Float t) const { return o + d * t; }
};

struct Bounds {
    Vec pMin, pMax;
    Bounds() {
        Float lo = std::numeric_limits<Float>::lowest(), hi = std::numeric_limits<Float>::max();
        pMin = Vec(hi, hi, hi);
        pMax = Vec(lo, lo, lo);
    }
    explicit Bounds(const Vec &p) : pMin(p), pMax(p) {}
    Bounds(const Vec &a, const Vec &b)
        : pMin(std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)),
          pMax(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)) {}
    const Vec &operator[](int i) const { return i == 0 ? pMin : pMax; }
    Vec Diagonal() const { return pMax - pMin; }
    Float SurfaceArea() const { Vec d = Diagonal(); return 2 * (d.x * d.y + d.x * d.z + d.y * d.z); }
    int MaximumExtent() const {
        Vec d = Diagonal();
        if (d.x > d.y && d.x > d.z) return 0;
        else if (d.y > d.z) return 1;
        else return 2;
    }
    Vec Offset(const Vec &p) const {
        Vec o = p - pMin;
        if (pMax.x > pMin.x) o.x /= pMax.x - pMin.x;
        if (pMax.y > pMin.y) o.y /= pMax.y - pMin.y;
        if (pMax.z > pMin.z) o.z /= pMax.z - pMin.z;
        return o;
    }
    Vec LerpP(const Vec &t) const { return Vec(Lerp(t.x, pMin.x, pMax.x), Lerp(t.y, pMin.y, pMax.y), Lerp(t.z, pMin.z, pMax.z)); }
    // geometry.h:1412-1438
    bool IntersectP(const Ray &ray, const Vec &invDir, const int dirIsNeg[3]) const {
        const Bounds &bounds = *this;
        Float tMin = (bounds[dirIsNeg[0]].x - ray.o.x) * invDir.x;
        Float tMax = (bounds[1 - dirIsNeg[0]].x - ray.o.x) * invDir.x;
        Float tyMin = (bounds[dirIsNeg[1]].y - ray.o.y) * invDir.y;
        Float tyMax = (bounds[1 - dirIsNeg[1]].y - ray.o.y) * invDir.y;
        tMax *= 1 + 2 * gamma(3);
        tyMax *= 1 + 2 * gamma(3);
        if (tMin > tyMax || tyMin > tMax) return false;
        if (tyMin > tMin) tMin = tyMin;
        if (tyMax < tMax) tMax = tyMax;
        Float tzMin = (bounds[dirIsNeg[2]].z - ray.o.z) * invDir.z;
        Float tzMax = (bounds[1 - dirIsNeg[2]].z - ray.o.z) * invDir.z;
        tzMax *= 1 + 2 * gamma(3);
        if (tMin > tzMax || tzMin > tMax) return false;
        if (tzMin > tMin) tMin = tzMin;
        if (tzMax < tMax) tMax = tzMax;
        return (tMin < ray.tMax) && (tMax > 0);
    }
};
inline Bounds Union(const Bounds &b, const Vec &p) {
    Bounds r;
    r.pMin = Vec(std::min(b.pMin.x, p.x), std::min(b.pMin.y, p.y), std::min(b.pMin.z, p.z));
    r.pMax = Vec(std::max(b.pMax.x, p.x), std::max(b.pMax.y, p.y), std::max(b.pMax.z, p.z));
    return r;
}
inline Bounds Union(const Bounds &a, const Bounds &b) {
    Bounds r;
    r.pMin = Vec(std::min(a.pMin.x, b.pMin.x), std::min(a.pMin.y, b.pMin.y), std::min(a.pMin.z, b.pMin.z));
    r.pMax = Vec(std::max(a.pMax.x, b.pMax.x), std::max(a.pMax.y, b.pMax.y), std::max(a.pMax.z, b.pMax.z));
    return r;
}

// ------------------------------------------------------------------ transforms (src/core/transform.h)
struct Xform {
    Float m[4][4], mInv[4][4];
    Xform() {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) m[i][j] = mInv[i][j] = (i == j) ? 1.f : 0.f;
    }
    Xform(const float a[16], const float b[16]) {
        std::memcpy(m, a, sizeof(m));
        std::memcpy(mInv, b, sizeof(mInv));
    }
    Vec Point(const Vec &p) const {  // transform.h:219-231
        Float x = p.x, y = p.y, z = p.z;
        Float xp = m[0][0] * x + m[0][1] * y + m[0][2] * z + m[0][3];
        Float yp = m[1][0] * x + m[1][1] * y + m[1][2] * z + m[1][3];
        Float zp = m[2][0] * x + m[2][1] * y + m[2][2] * z + m[2][3];
        Float wp = m[3][0] * x + m[3][1] * y + m[3][2] * z + m[3][3];
        if (wp == 1) return Vec(xp, yp, zp);
        Float inv = (Float)1 / wp;
        return Vec(inv * xp, inv * yp, inv * zp);
    }
    Vec Vector(const Vec &v) const {  // transform.h:233-239
        Float x = v.x, y = v.y, z = v.z;
        return Vec(m[0][0] * x + m[0][1] * y + m[0][2] * z, m[1][0] * x + m[1][1] * y + m[1][2] * z,
                   m[2][0] * x + m[2][1] * y + m[2][2] * z);
    }
    Vec Normal(const Vec &n) const {  // transform.h:241-249
        Float x = n.x, y = n.y, z = n.z;
        return Vec(mInv[0][0] * x + mInv[1][0] * y + mInv[2][0] * z, mInv[0][1] * x + mInv[1][1] * y + mInv[2][1] * z,
                   mInv[0][2] * x + mInv[1][2] * y + mInv[2][2] * z);
    }
    Vec PointErr(const Vec &p, Vec *pError) const {  // transform.h:277-301
        Float x = p.x, y = p.y, z = p.z;
        Float xp = (m[0][0] * x + m[0][1] * y) + (m[0][2] * z + m[0][3]);
        Float yp = (m[1][0] * x + m[1][1] * y) + (m[1][2] * z + m[1][3]);
        Float zp = (m[2][0] * x + m[2][1] * y) + (m[2][2] * z + m[2][3]);
        Float wp = (m[3][0] * x + m[3][1] * y) + (m[3][2] * z + m[3][3]);
        Float xAbsSum = (std::abs(m[0][0] * x) + std::abs(m[0][1] * y) + std::abs(m[0][2] * z) + std::abs(m[0][3]));
        Float yAbsSum = (std::abs(m[1][0] * x) + std::abs(m[1][1] * y) + std::abs(m[1][2] * z) + std::abs(m[1][3]));
        Float zAbsSum = (std::abs(m[2][0] * x) + std::abs(m[2][1] * y) + std::abs(m[2][2] * z) + std::abs(m[2][3]));
        *pError = gamma(3) * Vec(xAbsSum, yAbsSum, zAbsSum);
        if (wp == 1) return Vec(xp, yp, zp);
        Float inv = (Float)1 / wp;
        return Vec(inv * xp, inv * yp, inv * zp);
    }
    Vec PointErrIn(const Vec &pt, const Vec &ptError, Vec *absError) const {  // transform.h:303-335
        Float x = pt.x, y = pt.y, z = pt.z;
        Float xp = (m[0][0] * x + m[0][1] * y) + (m[0][2] * z + m[0][3]);
        Float yp = (m[1][0] * x + m[1][1] * y) + (m[1][2] * z + m[1][3]);
        Float zp = (m[2][0] * x + m[2][1] * y) + (m[2][2] * z + m[2][3]);
        Float wp = (m[3][0] * x + m[3][1] * y) + (m[3][2] * z + m[3][3]);
        for (int r = 0; r < 3; ++r)
            (*absError)[r] = (gamma(3) + (Float)1) * (std::abs(m[r][0]) * ptError.x + std::abs(m[r][1]) * ptError.y + std::abs(m[r][2]) * ptError.z) +
                             gamma(3) * (std::abs(m[r][0] * x) + std::abs(m[r][1] * y) + std::abs(m[r][2] * z) + std::abs(m[r][3]));
        if (wp == 1.) return Vec(xp, yp, zp);
        Float inv = (Float)1 / wp;
        return Vec(inv * xp, inv * yp, inv * zp);
    }
    Vec VectorErr(const Vec &v, Vec *absError) const {  // transform.h:337-352
        Float x = v.x, y = v.y, z = v.z;
        for (int r = 0; r < 3; ++r)
            (*absError)[r] = gamma(3) * (std::abs(m[r][0] * v.x) + std::abs(m[r][1] * v.y) + std::abs(m[r][2] * v.z));
        return Vec(m[0][0] * x + m[0][1] * y + m[0][2] * z, m[1][0] * x + m[1][1] * y + m[1][2] * z,
                   m[2][0] * x + m[2][1] * y + m[2][2] * z);
    }
    Ray RayPlain(const Ray &r) const {  // transform.h:251-264
        Vec oError;
        Vec o = PointErr(r.o, &oError);
        Vec d = Vector(r.d);
        Float lengthSquared = d.LengthSquared();
        Float tMax = r.tMax;
        if (lengthSquared > 0) {
            Float dt = Dot(Abs(d), oError) / lengthSquared;
            o = o + d * dt;
            tMax -= dt;
        }
        return Ray(o, d, tMax);
    }
    Ray RayErr(const Ray &r, Vec *oError, Vec *dError) const {  // transform.h:382-394 (tMax not reduced)
        Vec o = PointErr(r.o, oError);
        Vec d = VectorErr(r.d, dError);
        Float tMax = r.tMax;
        Float lengthSquared = d.LengthSquared();
        if (lengthSquared > 0) {
            Float dt = Dot(Abs(d), *oError) / lengthSquared;
            o = o + d * dt;
        }
        return Ray(o, d, tMax);
    }
    Bounds BoundsOp(const Bounds &b) const {  // transform.cpp:238-249
        Bounds ret;
        ret.pMin = ret.pMax = Point(Vec(b.pMin.x, b.pMin.y, b.pMin.z));
        ret = Union(ret, Point(Vec(b.pMax.x, b.pMin.y, b.pMin.z)));
        ret = Union(ret, Point(Vec(b.pMin.x, b.pMax.y, b.pMin.z)));
        ret = Union(ret, Point(Vec(b.pMin.x, b.pMin.y, b.pMax.z)));
        ret = Union(ret, Point(Vec(b.pMin.x, b.pMax.y, b.pMax.z)));
        ret = Union(ret, Point(Vec(b.pMax.x, b.pMax.y, b.pMin.z)));
        ret = Union(ret, Point(Vec(b.pMax.x, b.pMin.y, b.pMax.z)));
        ret = Union(ret, Point(Vec(b.pMax.x, b.pMax.y, b.pMax.z)));
        return ret;
    }
    bool IsIdentity() const {  // transform.h:137-143
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                if (m[i][j] != ((i == j) ? 1.f : 0.f)) return false;
        return true;
    }
    Xform Inverse() const {
        Xform t;
        std::memcpy(t.m, mInv, sizeof(m));
        std::memcpy(t.mInv, m, sizeof(m));
        return t;
    }
};

// ------------------------------------------------------------------ spectrum (src/core/spectrum.h, RGB)
struct Spectrum {
    Float c[3];
    Spectrum(Float v = 0.f) { c[0] = c[1] = c[2] = v; }
    Spectrum(Float r, Float g, Float b) { c[0] = r; c[1] = g; c[2] = b; }
    Spectrum operator+(const Spectrum &s) const { return Spectrum(c[0] + s.c[0], c[1] + s.c[1], c[2] + s.c[2]); }
    Spectrum &operator+=(const Spectrum &s) { for (int i = 0; i < 3; ++i) c[i] += s.c[i]; return *this; }
    Spectrum operator-(const Spectrum &s) const { return Spectrum(c[0] - s.c[0], c[1] - s.c[1], c[2] - s.c[2]); }
    Spectrum operator*(const Spectrum &s) const { return Spectrum(c[0] * s.c[0], c[1] * s.c[1], c[2] * s.c[2]); }
    Spectrum &operator*=(const Spectrum &s) { for (int i = 0; i < 3; ++i) c[i] *= s.c[i]; return *this; }
    Spectrum operator*(Float a) const { return Spectrum(c[0] * a, c[1] * a, c[2] * a); }
    Spectrum operator/(Float a) const { return Spectrum(c[0] / a, c[1] / a, c[2] / a); }
    Spectrum operator/(const Spectrum &s) const { return Spectrum(c[0] / s.c[0], c[1] / s.c[1], c[2] / s.c[2]); }
    Spectrum operator-() const { return Spectrum(-c[0], -c[1], -c[2]); }
    Spectrum &operator/=(Float a) { for (int i = 0; i < 3; ++i) c[i] /= a; return *this; }
    bool IsBlack() const { return c[0] == 0. && c[1] == 0. && c[2] == 0.; }
    bool HasNaNs() const { return std::isnan(c[0]) || std::isnan(c[1]) || std::isnan(c[2]); }
    Float y() const { const Float w[3] = {0.212671f, 0.715160f, 0.072169f}; return w[0] * c[0] + w[1] * c[1] + w[2] * c[2]; }
    Float MaxComponentValue() const { Float m = c[0]; for (int i = 1; i < 3; ++i) m = std::max(m, c[i]); return m; }
    Spectrum Clamp(Float low = 0, Float high = Infinity) const {
        return Spectrum(orc::Clamp(c[0], low, high), orc::Clamp(c[1], low, high), orc::Clamp(c[2], low, high));
    }
};
inline Spectrum operator*(Float a, const Spectrum &s) { return s * a; }
inline Spectrum Sqrt(const Spectrum &s) { return Spectrum(std::sqrt(s.c[0]), std::sqrt(s.c[1]), std::sqrt(s.c[2])); }  // spectrum.h:206-211
inline void RGBToXYZ(const Float rgb[3], Float xyz[3]) {  // spectrum.h:62-66
    xyz[0] = 0.412453f * rgb[0] + 0.357580f * rgb[1] + 0.180423f * rgb[2];
    xyz[1] = 0.212671f * rgb[0] + 0.715160f * rgb[1] + 0.072169f * rgb[2];
    xyz[2] = 0.019334f * rgb[0] + 0.119193f * rgb[1] + 0.950227f * rgb[2];
}
inline void XYZToRGB(const Float xyz[3], Float rgb[3]) {  // spectrum.h:56-60
    rgb[0] = 3.240479f * xyz[0] - 1.537150f * xyz[1] - 0.498535f * xyz[2];
    rgb[1] = -0.969256f * xyz[0] + 1.875991f * xyz[1] + 0.041556f * xyz[2];
    rgb[2] = 0.055648f * xyz[0] - 0.204043f * xyz[1] + 1.057311f * xyz[2];
}

// ------------------------------------------------------------------ interactions (src/core/interaction.h)
class Primitive;
class BSDF;
struct Interaction {
    Vec p, pError, wo, n;
    Interaction() {}
    Interaction(const Vec &p, const Vec &n, const Vec &pError, const Vec &wo) : p(p), pError(pError), wo(Normalize(wo)), n(n) {}
    Ray SpawnRay(const Vec &d) const { return Ray(OffsetRayOrigin(p, pError, n, d), d, Infinity); }
    Ray SpawnRayTo(const Interaction &it) const {  // interaction.h:73-78
        Vec origin = OffsetRayOrigin(p, pError, n, it.p - p);
        Vec target = OffsetRayOrigin(it.p, it.pError, it.n, origin - it.p);
        Vec d = target - origin;
        return Ray(origin, d, 1 - ShadowEpsilon);
    }
};
struct SurfaceInteraction : public Interaction {
    Vec2 uv;
    Vec dpdu, dpdv;
    struct { Vec n, dpdu, dpdv; } shading;
    const Primitive *primitive = nullptr;
    bool reverseOrientation = false, transformSwapsHandedness = false;
    SurfaceInteraction() {}
    // interaction.cpp:44-71
    SurfaceInteraction(const Vec &p, const Vec &pError, const Vec2 &uv, const Vec &wo, const Vec &dpdu, const Vec &dpdv,
                       bool reverseOrientation, bool swapsHandedness)
        : Interaction(p, Normalize(Cross(dpdu, dpdv)), pError, wo), uv(uv), dpdu(dpdu), dpdv(dpdv),
          reverseOrientation(reverseOrientation), transformSwapsHandedness(swapsHandedness) {
        shading.n = n;
        shading.dpdu = dpdu;
        shading.dpdv = dpdv;
        if (reverseOrientation ^ swapsHandedness) {
            n = n * -1;
            shading.n = shading.n * -1;
        }
    }
    // interaction.cpp:73-90
    void SetShadingGeometry(const Vec &dpdus, const Vec &dpdvs, bool orientationIsAuthoritative) {
        shading.n = Normalize(Cross(dpdus, dpdvs));
        if (orientationIsAuthoritative) n = Faceforward(n, shading.n);
        else shading.n = Faceforward(shading.n, n);
        shading.dpdu = dpdus;
        shading.dpdv = dpdvs;
    }
};

}  // namespace orc

#include "pb2_oracle_shapes.inc"
#include "pb2_oracle_shading.inc"
#include "pb2_oracle_render.inc"
