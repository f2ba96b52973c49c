// Scalar/vector helpers for the device hot path.
//
// Parity rule for this file and everything that includes it: the translation unit is compiled with
// -fmad=false and IEEE div/sqrt, every expression keeps the reference's operation order, and the
// places where the reference computes in double (Cross, geometry.h:957-963; base-2 radical inverse)
// do so here too.  With that, ray/box, ray/triangle, sampler and geometry code produce the
// reference's bits; only libdevice transcendentals (sinf/cosf/logf/atan2f/acosf) can differ from
// glibc by an ulp or two.
//
// Functions are PB2_HD (__host__ __device__) so that tests can also compile them with g++ and step
// through the same per-lane code on the CPU when chasing a parity difference; the shipped library
// only ever launches them on the GPU.
#ifndef PB2_MATH_CUH
#define PB2_MATH_CUH

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PB2_HD __host__ __device__ __forceinline__
#define PB2_D __device__ __forceinline__
#define PB2_HDN __host__ __device__ __noinline__
#else
#define PB2_HD inline
#define PB2_D inline
#define PB2_HDN inline
#endif

namespace pb2 {

#define PB2_INFINITY (__builtin_huge_valf())
constexpr float kMachineEpsilon = 5.9604644775390625e-08f;  // 2^-24 (pbrt.h:201-202)
constexpr float kShadowEpsilon = 0.0001f;
constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kPiOver2 = 1.57079632679489661923f;
constexpr float kPiOver4 = 0.78539816339744830961f;
constexpr float kOneMinusEpsilon = 0x1.fffffep-1f;
constexpr float gammaf_c(int n) { return (n * kMachineEpsilon) / (1 - n * kMachineEpsilon); }  // pbrt.h:289-291
constexpr float kGamma2 = gammaf_c(2), kGamma3 = gammaf_c(3), kGamma5 = gammaf_c(5), kGamma6 = gammaf_c(6),
                kGamma7 = gammaf_c(7);
constexpr float kSlabScale = 1 + 2 * gammaf_c(3);  // geometry.h:1422

struct V3 {
    float x, y, z;
};
struct V2 {
    float x, y;
};

// std::max / std::min semantics (first argument wins on ties and NaNs), not fmaxf/fminf
PB2_HD float pmax(float a, float b) { return (a < b) ? b : a; }
PB2_HD float pmin(float a, float b) { return (b < a) ? b : a; }
PB2_HD V3 mk3(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
PB2_HD V2 mk2(float x, float y) { V2 v; v.x = x; v.y = y; return v; }
PB2_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PB2_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PB2_HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
PB2_HD V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
PB2_HD V3 operator*(float s, V3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
PB2_HD V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
PB2_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PB2_HD float absDot(V3 a, V3 b) { return fabsf(dot(a, b)); }
PB2_HD float lengthSquared(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
PB2_HD float length(V3 a) { return sqrtf(lengthSquared(a)); }
PB2_HD V3 vabs(V3 a) { return mk3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
// v / f multiplies by the reciprocal (geometry.h:244-248)
PB2_HD V3 divf(V3 a, float f) { float inv = 1.f / f; return mk3(a.x * inv, a.y * inv, a.z * inv); }
PB2_HD V3 normalize(V3 a) { return divf(a, length(a)); }
PB2_HD float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// geometry.h:957-963: double products, single rounding
PB2_HD V3 cross(V3 a, V3 b) {
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return mk3((float)((ay * bz) - (az * by)), (float)((az * bx) - (ax * bz)), (float)((ax * by) - (ay * bx)));
}
PB2_HD V3 faceforward(V3 n, V3 v) { return (dot(n, v) < 0.f) ? -n : n; }
PB2_HD float maxComponent(V3 a) { return pmax(a.x, pmax(a.y, a.z)); }
// geometry.h:1020-1027
PB2_HD void coordinateSystem(V3 v1, V3 *v2, V3 *v3) {
    if (fabsf(v1.x) > fabsf(v1.y))
        *v2 = divf(mk3(-v1.z, 0, v1.x), sqrtf(v1.x * v1.x + v1.z * v1.z));
    else
        *v2 = divf(mk3(0, v1.z, -v1.y), sqrtf(v1.y * v1.y + v1.z * v1.z));
    *v3 = cross(v1, *v2);
}
// Transcendentals.  The reference calls glibc's float functions, which are (almost always) the
// correctly rounded result; libdevice's sinf/cosf/logf/atan2f/acosf are allowed 1-2 ulp.  On the
// device the double-precision routine rounded to float gives the correctly rounded value in all but
// vanishingly rare cases, which removes nearly all last-bit differences from sampled directions (they
// matter: a 1-ulp change of a direction moves a roughness-0.025 microfacet lobe's value by percents).
// These run a few times per path vertex, never inside the traversal loop.
#if defined(__CUDA_ARCH__)
PB2_HD float psinf(float x) { return (float)sin((double)x); }
PB2_HD float pcosf(float x) { return (float)cos((double)x); }
// both at once: one argument reduction instead of two (sincos() evaluates the same kernels as sin() and cos())
PB2_HD void psincosf(float x, float *s, float *c) {
    double ds, dc;
    sincos((double)x, &ds, &dc);
    *s = (float)ds;
    *c = (float)dc;
}
PB2_HD float plogf(float x) { return (float)log((double)x); }
PB2_HD float patan2f(float y, float x) { return (float)atan2((double)y, (double)x); }
PB2_HD float pacosf(float x) { return (float)acos((double)x); }
#else
PB2_HD float psinf(float x) { return sinf(x); }
PB2_HD float pcosf(float x) { return cosf(x); }
PB2_HD void psincosf(float x, float *s, float *c) { *s = sinf(x); *c = cosf(x); }
PB2_HD float plogf(float x) { return logf(x); }
PB2_HD float patan2f(float y, float x) { return atan2f(y, x); }
PB2_HD float pacosf(float x) { return acosf(x); }
#endif
PB2_HD float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
PB2_HD float lerpf(float t, float a, float b) { return (1 - t) * a + t * b; }

PB2_HD uint32_t floatBits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PB2_HD float bitsFloat(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
// pbrt.h:241-265
PB2_HD float nextFloatUp(float v) {
    if (isinf(v) && v > 0.f) return v;
    if (v == -0.f) v = 0.f;
    uint32_t ui = floatBits(v);
    if (v >= 0) ++ui; else --ui;
    return bitsFloat(ui);
}
PB2_HD float nextFloatDown(float v) {
    if (isinf(v) && v < 0.f) return v;
    if (v == 0.f) v = -0.f;
    uint32_t ui = floatBits(v);
    if (v > 0) --ui; else ++ui;
    return bitsFloat(ui);
}
// geometry.h:1440-1454
PB2_HD V3 offsetRayOrigin(V3 p, V3 pError, V3 n, V3 w) {
    float d = dot(vabs(n), pError);
    V3 offset = d * n;
    if (dot(w, n) < 0) offset = -offset;
    V3 po = p + offset;
    if (offset.x > 0) po.x = nextFloatUp(po.x); else if (offset.x < 0) po.x = nextFloatDown(po.x);
    if (offset.y > 0) po.y = nextFloatUp(po.y); else if (offset.y < 0) po.y = nextFloatDown(po.y);
    if (offset.z > 0) po.z = nextFloatUp(po.z); else if (offset.z < 0) po.z = nextFloatDown(po.z);
    return po;
}

// RGBSpectrum (spectrum.h:430-470)
PB2_HD float luminance(V3 c) { return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z; }
PB2_HD bool isBlack(V3 c) { return c.x == 0.f && c.y == 0.f && c.z == 0.f; }
PB2_HD float maxComponentValue(V3 c) { return pmax(c.x, pmax(c.y, c.z)); }

// Row-major 4x4 applied as the reference's Transform does (transform.h:219-264, 277-334).
struct M44 { float m[4][4]; };
PB2_HD V3 xfPoint(const M44 &t, V3 p) {
    float x = p.x, y = p.y, z = p.z;
    float xp = t.m[0][0] * x + t.m[0][1] * y + t.m[0][2] * z + t.m[0][3];
    float yp = t.m[1][0] * x + t.m[1][1] * y + t.m[1][2] * z + t.m[1][3];
    float zp = t.m[2][0] * x + t.m[2][1] * y + t.m[2][2] * z + t.m[2][3];
    float wp = t.m[3][0] * x + t.m[3][1] * y + t.m[3][2] * z + t.m[3][3];
    if (wp == 1) return mk3(xp, yp, zp);
    float inv = 1.f / wp;  // Point3::operator/ (geometry.h:499-503)
    return mk3(inv * xp, inv * yp, inv * zp);
}
PB2_HD V3 xfVector(const M44 &t, V3 v) {
    float x = v.x, y = v.y, z = v.z;
    return mk3(t.m[0][0] * x + t.m[0][1] * y + t.m[0][2] * z, t.m[1][0] * x + t.m[1][1] * y + t.m[1][2] * z,
               t.m[2][0] * x + t.m[2][1] * y + t.m[2][2] * z);
}
// Normal through the INVERSE matrix, transposed (transform.h:241-249)
PB2_HD V3 xfNormalInv(const M44 &inv, V3 n) {
    float x = n.x, y = n.y, z = n.z;
    return mk3(inv.m[0][0] * x + inv.m[1][0] * y + inv.m[2][0] * z, inv.m[0][1] * x + inv.m[1][1] * y + inv.m[2][1] * z,
               inv.m[0][2] * x + inv.m[1][2] * y + inv.m[2][2] * z);
}
// transform.h:277-301: point with absolute error bound of the transform itself
PB2_HD V3 xfPointErr(const M44 &t, V3 p, V3 *pError) {
    float x = p.x, y = p.y, z = p.z;
    float xp = (t.m[0][0] * x + t.m[0][1] * y) + (t.m[0][2] * z + t.m[0][3]);
    float yp = (t.m[1][0] * x + t.m[1][1] * y) + (t.m[1][2] * z + t.m[1][3]);
    float zp = (t.m[2][0] * x + t.m[2][1] * y) + (t.m[2][2] * z + t.m[2][3]);
    float wp = (t.m[3][0] * x + t.m[3][1] * y) + (t.m[3][2] * z + t.m[3][3]);
    float xAbs = (fabsf(t.m[0][0] * x) + fabsf(t.m[0][1] * y) + fabsf(t.m[0][2] * z) + fabsf(t.m[0][3]));
    float yAbs = (fabsf(t.m[1][0] * x) + fabsf(t.m[1][1] * y) + fabsf(t.m[1][2] * z) + fabsf(t.m[1][3]));
    float zAbs = (fabsf(t.m[2][0] * x) + fabsf(t.m[2][1] * y) + fabsf(t.m[2][2] * z) + fabsf(t.m[2][3]));
    *pError = kGamma3 * mk3(xAbs, yAbs, zAbs);
    if (wp == 1) return mk3(xp, yp, zp);
    float inv = 1.f / wp;
    return mk3(inv * xp, inv * yp, inv * zp);
}

}  // namespace pb2
#endif
