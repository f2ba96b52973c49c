// Sphere as a second leaf-primitive type (only needed because scenes/killeroo-simple.pbrt lights
// the scene with a sphere): interval-arithmetic intersection, surface interaction, and
// area-light sampling.
//
//   EFloat, Quadratic                              src/core/efloat.h:48-286
//   Sphere::Intersect / IntersectP                 src/shapes/sphere.cpp:49-215
//   Sphere::Sample(u) / Sample(ref,u) / Pdf        src/shapes/sphere.cpp:219-315
//   Transform ops with error bounds                src/core/transform.h:303-394, transform.cpp:262-297
//
// Included from pb2_shade.cuh after DInteraction / DLightSample-free helpers are defined.
#ifndef PB2_SPHERE_CUH
#define PB2_SPHERE_CUH

namespace pb2 {

struct EF {  // EFloat: value with a conservative [low, high] interval
    float v, low, high;
};
PB2_HD EF efMake(float v, float err = 0.f) {
    EF r;
    r.v = v;
    if (err == 0.f) r.low = r.high = v;
    else {
        r.low = nextFloatDown(v - err);
        r.high = nextFloatUp(v + err);
    }
    return r;
}
PB2_HD EF efAdd(EF a, EF b) {
    EF r;
    r.v = a.v + b.v;
    r.low = nextFloatDown(a.low + b.low);
    r.high = nextFloatUp(a.high + b.high);
    return r;
}
PB2_HD EF efSub(EF a, EF b) {
    EF r;
    r.v = a.v - b.v;
    r.low = nextFloatDown(a.low - b.high);
    r.high = nextFloatUp(a.high - b.low);
    return r;
}
PB2_HD EF efMul(EF a, EF b) {
    EF r;
    r.v = a.v * b.v;
    float p0 = a.low * b.low, p1 = a.high * b.low, p2 = a.low * b.high, p3 = a.high * b.high;
    r.low = nextFloatDown(pmin(pmin(p0, p1), pmin(p2, p3)));
    r.high = nextFloatUp(pmax(pmax(p0, p1), pmax(p2, p3)));
    return r;
}
PB2_HD EF efDiv(EF a, EF b) {
    EF r;
    r.v = a.v / b.v;
    if (b.low < 0 && b.high > 0) {
        r.low = -PB2_INFINITY;
        r.high = PB2_INFINITY;
    } else {
        float d0 = a.low / b.low, d1 = a.high / b.low, d2 = a.low / b.high, d3 = a.high / b.high;
        r.low = nextFloatDown(pmin(pmin(d0, d1), pmin(d2, d3)));
        r.high = nextFloatUp(pmax(pmax(d0, d1), pmax(d2, d3)));
    }
    return r;
}
// Quadratic(EFloat A, B, C) (efloat.h:268-286)
PB2_HD bool efQuadratic(EF A, EF B, EF C, EF *t0, EF *t1) {
    double discrim = (double)B.v * (double)B.v - 4. * (double)A.v * (double)C.v;
    if (discrim < 0.) return false;
    double rootDiscrim = sqrt(discrim);
    EF floatRootDiscrim = efMake((float)rootDiscrim, (float)((double)kMachineEpsilon * rootDiscrim));
    EF q;
    // "-.5 * (B -/+ root)": the double literal converts to float first (operator*(float, EFloat))
    if (B.v < 0) q = efMul(efMake(-.5f), efSub(B, floatRootDiscrim));
    else q = efMul(efMake(-.5f), efAdd(B, floatRootDiscrim));
    *t0 = efDiv(q, A);
    *t1 = efDiv(C, q);
    if (t0->v > t1->v) {
        EF tmp = *t0;
        *t0 = *t1;
        *t1 = tmp;
    }
    return true;
}

PB2_HD M44 loadM44(const float m[16]) {
    M44 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = m[4 * i + j];
    return r;
}
// Transform::operator()(const Vector3&, Vector3 *absError) (transform.h:337-352)
PB2_HD V3 xfVectorErr(const M44 &t, V3 v, V3 *e) {
    float x = v.x, y = v.y, z = v.z;
    e->x = kGamma3 * (fabsf(t.m[0][0] * v.x) + fabsf(t.m[0][1] * v.y) + fabsf(t.m[0][2] * v.z));
    e->y = kGamma3 * (fabsf(t.m[1][0] * v.x) + fabsf(t.m[1][1] * v.y) + fabsf(t.m[1][2] * v.z));
    e->z = kGamma3 * (fabsf(t.m[2][0] * v.x) + fabsf(t.m[2][1] * v.y) + fabsf(t.m[2][2] * v.z));
    return mk3(t.m[0][0] * x + t.m[0][1] * y + t.m[0][2] * z, t.m[1][0] * x + t.m[1][1] * y + t.m[1][2] * z,
               t.m[2][0] * x + t.m[2][1] * y + t.m[2][2] * z);
}
// Transform::operator()(const Point3&, const Vector3 &ptError, Vector3 *absError) (transform.h:303-335)
PB2_HD V3 xfPointErrIn(const M44 &t, V3 p, V3 pe, V3 *e) {
    float x = p.x, y = p.y, z = p.z;
    float xp = (t.m[0][0] * x + t.m[0][1] * y) + (t.m[0][2] * z + t.m[0][3]);
    float yp = (t.m[1][0] * x + t.m[1][1] * y) + (t.m[1][2] * z + t.m[1][3]);
    float zp = (t.m[2][0] * x + t.m[2][1] * y) + (t.m[2][2] * z + t.m[2][3]);
    float wp = (t.m[3][0] * x + t.m[3][1] * y) + (t.m[3][2] * z + t.m[3][3]);
    e->x = (kGamma3 + 1.f) * (fabsf(t.m[0][0]) * pe.x + fabsf(t.m[0][1]) * pe.y + fabsf(t.m[0][2]) * pe.z) +
           kGamma3 * (fabsf(t.m[0][0] * x) + fabsf(t.m[0][1] * y) + fabsf(t.m[0][2] * z) + fabsf(t.m[0][3]));
    e->y = (kGamma3 + 1.f) * (fabsf(t.m[1][0]) * pe.x + fabsf(t.m[1][1]) * pe.y + fabsf(t.m[1][2]) * pe.z) +
           kGamma3 * (fabsf(t.m[1][0] * x) + fabsf(t.m[1][1] * y) + fabsf(t.m[1][2] * z) + fabsf(t.m[1][3]));
    e->z = (kGamma3 + 1.f) * (fabsf(t.m[2][0]) * pe.x + fabsf(t.m[2][1]) * pe.y + fabsf(t.m[2][2]) * pe.z) +
           kGamma3 * (fabsf(t.m[2][0] * x) + fabsf(t.m[2][1] * y) + fabsf(t.m[2][2] * z) + fabsf(t.m[2][3]));
    if (wp == 1.) return mk3(xp, yp, zp);
    float inv = 1.f / wp;
    return mk3(inv * xp, inv * yp, inv * zp);
}

struct SphereRayHit {
    float tHit, phi;
    V3 pHit;      // object space, refined
    V3 oRayD;     // object-space ray direction
};

// The geometric part shared by Sphere::Intersect and IntersectP (sphere.cpp:49-103 / 158-214).
PB2_HDN bool sphereTest(const pb2_sphere &s, const DRay &r, float rayTMax, SphereRayHit *out) {
    M44 w2o = loadM44(s.world_to_object);
    // Transform::operator()(Ray, &oErr, &dErr) (transform.h:382-394): tMax is NOT reduced
    V3 oErr, dErr;
    V3 o = xfPointErr(w2o, r.o, &oErr);
    V3 d = xfVectorErr(w2o, r.d, &dErr);
    float l2 = lengthSquared(d);
    if (l2 > 0) {
        float dt = dot(vabs(d), oErr) / l2;
        o = o + d * dt;
    }
    EF ox = efMake(o.x, oErr.x), oy = efMake(o.y, oErr.y), oz = efMake(o.z, oErr.z);
    EF dx = efMake(d.x, dErr.x), dy = efMake(d.y, dErr.y), dz = efMake(d.z, dErr.z);
    EF a = efAdd(efAdd(efMul(dx, dx), efMul(dy, dy)), efMul(dz, dz));
    EF b = efMul(efMake(2.f), efAdd(efAdd(efMul(dx, ox), efMul(dy, oy)), efMul(dz, oz)));
    EF c = efSub(efAdd(efAdd(efMul(ox, ox), efMul(oy, oy)), efMul(oz, oz)), efMul(efMake(s.radius), efMake(s.radius)));
    EF t0, t1;
    if (!efQuadratic(a, b, c, &t0, &t1)) return false;
    if (t0.high > rayTMax || t1.low <= 0) return false;
    EF tShapeHit = t0;
    if (tShapeHit.low <= 0) {
        tShapeHit = t1;
        if (tShapeHit.high > rayTMax) return false;
    }
    V3 pHit = o + d * tShapeHit.v;
    pHit = pHit * (s.radius / length(pHit));  // Distance(pHit, origin)
    if (pHit.x == 0 && pHit.y == 0) pHit.x = 1e-5f * s.radius;
    float phi = patan2f(pHit.y, pHit.x);
    if (phi < 0) phi += 2 * kPi;
    if ((s.z_min > -s.radius && pHit.z < s.z_min) || (s.z_max < s.radius && pHit.z > s.z_max) || phi > s.phi_max) {
        if (tShapeHit.v == t1.v) return false;
        if (t1.high > rayTMax) return false;
        tShapeHit = t1;
        pHit = o + d * tShapeHit.v;
        pHit = pHit * (s.radius / length(pHit));
        if (pHit.x == 0 && pHit.y == 0) pHit.x = 1e-5f * s.radius;
        phi = patan2f(pHit.y, pHit.x);
        if (phi < 0) phi += 2 * kPi;
        if ((s.z_min > -s.radius && pHit.z < s.z_min) || (s.z_max < s.radius && pHit.z > s.z_max) || phi > s.phi_max)
            return false;
    }
    out->tHit = tShapeHit.v;
    out->phi = phi;
    out->pHit = pHit;
    out->oRayD = d;
    return true;
}

PB2_HDN bool sphereLeafTest(const DScene &sc, int sphereIndex, const DRay &ray, float rayTMax, float *tHit, float *phi) {
    SphereRayHit h;
    if (!sphereTest(sc.spheres[sphereIndex], ray, rayTMax, &h)) return false;
    *tHit = h.tHit;
    *phi = h.phi;
    return true;
}

// SurfaceInteraction of a sphere hit (sphere.cpp:105-155) mapped to world space
// (transform.cpp:262-297).  The hit is recomputed from the ray with the tMax the traversal saw just
// before accepting it: any tMax >= tHit accepts the same root, so +inf is used.
PB2_HDN DInteraction sphereInteraction(const DScene &sc, int prim, const DRay &ray, float tHit, float, DTexGeom *tg = nullptr) {
    (void)tHit;
    DInteraction it;
    const pb2_sphere s = sc.spheres[sc.primIndex[prim]];
    SphereRayHit h;
    h.tHit = 0; h.phi = 0; h.pHit = mk3(0, 0, 0); h.oRayD = mk3(0, 0, 0);
    sphereTest(s, ray, PB2_INFINITY, &h);
    V3 pHit = h.pHit;
    float u = h.phi / s.phi_max;
    float theta = pacosf(clampf(pHit.z / s.radius, -1.f, 1.f));
    float v = (theta - s.theta_min) / (s.theta_max - s.theta_min);
    float zRadius = sqrtf(pHit.x * pHit.x + pHit.y * pHit.y);
    float invZRadius = 1 / zRadius;
    float cosPhiV = pHit.x * invZRadius;
    float sinPhiV = pHit.y * invZRadius;
    V3 dpdu = mk3(-s.phi_max * pHit.y, s.phi_max * pHit.x, 0);
    V3 dpdv = (s.theta_max - s.theta_min) * mk3(pHit.z * cosPhiV, pHit.z * sinPhiV, -s.radius * psinf(theta));
    V3 pError = kGamma5 * vabs(pHit);
    // object-space SurfaceInteraction ctor (interaction.cpp:44-71)
    V3 n = normalize(cross(dpdu, dpdv));
    if ((s.reverse_orientation != 0) ^ (s.transform_swaps_handedness != 0)) n = n * -1.f;
    V3 nsObj = n;
    // to world space
    M44 o2w = loadM44(s.object_to_world), w2o = loadM44(s.world_to_object);
    it.p = xfPointErrIn(o2w, pHit, pError, &it.pError);
    it.n = normalize(xfNormalInv(w2o, n));
    it.wo = normalize(xfVector(o2w, normalize(-h.oRayD)));
    it.uv = mk2(u, v);
    it.dpdus = xfVector(o2w, dpdu);
    if (tg) {
        tg->dpdu = it.dpdus;
        tg->dpdv = tg->dpdvs = xfVector(o2w, dpdv);
        // dndu, dndv from the fundamental forms (sphere.cpp:122-143), taken to world space as normals
        const V3 d2Pduu = (-s.phi_max * s.phi_max) * mk3(pHit.x, pHit.y, 0);
        const V3 d2Pduv = (((s.theta_max - s.theta_min) * pHit.z) * s.phi_max) * mk3(-sinPhiV, cosPhiV, 0.f);
        const V3 d2Pdvv = (-(s.theta_max - s.theta_min) * (s.theta_max - s.theta_min)) * mk3(pHit.x, pHit.y, pHit.z);
        const float E = dot(dpdu, dpdu), F = dot(dpdu, dpdv), G = dot(dpdv, dpdv);
        const V3 Nn = normalize(cross(dpdu, dpdv));
        const float e = dot(Nn, d2Pduu), f = dot(Nn, d2Pduv), g = dot(Nn, d2Pdvv);
        const float invEGF2 = 1 / (E * G - F * F);
        const V3 dndu = ((f * F - e * G) * invEGF2) * dpdu + ((e * F - f * E) * invEGF2) * dpdv;
        const V3 dndv = ((g * F - f * G) * invEGF2) * dpdu + ((f * F - g * E) * invEGF2) * dpdv;
        tg->dndus = xfNormalInv(w2o, dndu);
        tg->dndvs = xfNormalInv(w2o, dndv);
    }
    it.ns = normalize(xfNormalInv(w2o, nsObj));
    it.ns = faceforward(it.ns, it.n);
    it.prim = prim;
    return it;
}

}  // namespace pb2

// ---- light sampling needs DLightSample, declared in pb2_shade.cuh after this include; forward
// declare the struct here and define the functions as templates-free inline code below it.
namespace pb2 {
struct DLightSample;
PB2_HD V3 lightL(const pb2_light &l, V3 n, V3 w);
PB2_HDN DLightSample sampleSphereLight(const DScene &sc, const pb2_light &l, const DInteraction &ref, V2 u);
PB2_HDN float sphereLightPdf(const DScene &sc, const pb2_light &l, const DInteraction &ref, V3 wi);
}  // namespace pb2
#endif
