// Host-side mirror of the reference's plugin interfaces for the path-tracing hot path.
// Class names, virtual signatures and factory names are the reference's, so calling code and
// scene files are unchanged; the bodies are new.  None of these classes traces a ray or shades a
// path on the CPU: Shape/Primitive/Aggregate/Scene::Intersect[P] and Integrator::Render hand the
// work to the CUDA library through the C ABI in include/pb2.h, and return an error (no fallback)
// when no device is available.
//
//   Shape                 src/core/shape.h:51-89          TriangleMesh/Triangle  src/shapes/triangle.h:46-114
//   Sphere                src/shapes/sphere.h:47-77       Material               src/core/material.h:51-61
//   AreaLight             src/core/light.h:100-112        DiffuseAreaLight       src/lights/diffuse.h:49-79
//   Primitive/Aggregate   src/core/primitive.h:51-127     BVHAccel               src/accelerators/bvh.h:51-98
//   Scene                 src/core/scene.h:50-80          Film                   src/core/film.h:58-105
//   Camera                src/core/camera.h:53-115        HaltonSampler          src/samplers/halton.h:49-82
//   Integrator            src/core/integrator.h:53-106    PathIntegrator         src/integrators/path.h:49-72
#ifndef PB2_HOST_SCENE_H
#define PB2_HOST_SCENE_H

#include "core.h"
#include "paramset.h"
#include "pb2.h"

namespace pbrt {

struct Ray {
    Point3f o;
    Vector3f d;
    mutable Float tMax = Infinity;  // in/out channel of the closest hit (geometry.h:887)
    Float time = 0;
    Ray() {}
    Ray(const Point3f &o, const Vector3f &d, Float tMax = Infinity, Float time = 0) : o(o), d(d), tMax(tMax), time(time) {}
};
typedef Ray RayDifferential;  // differentials only feed filtered textures (out of scope)

class Primitive;
class Shape;
class Scene;
struct DeviceScene;  // owns the pb2_scene handle of a flattened aggregate

struct SurfaceInteraction {
    Point3f p;
    Vector3f pError;
    Vector3f wo;
    Normal3f n;
    Point2f uv;
    struct { Normal3f n; Vector3f dpdu; } shading;
    Float b[3] = {0, 0, 0};
    const Primitive *primitive = nullptr;
};

// ---------------------------------------------------------------- shapes
class Shape {
  public:
    Shape(const Transform *ObjectToWorld, const Transform *WorldToObject, bool reverseOrientation)
        : ObjectToWorld(ObjectToWorld), WorldToObject(WorldToObject), reverseOrientation(reverseOrientation),
          transformSwapsHandedness(ObjectToWorld->SwapsHandedness()) {}
    virtual ~Shape() {}
    virtual Bounds3f ObjectBound() const = 0;
    virtual Bounds3f WorldBound() const { return (*ObjectToWorld)(ObjectBound()); }
    // Routed to the device through a one-primitive aggregate (see host_device.cpp).
    virtual bool Intersect(const Ray &ray, Float *tHit, SurfaceInteraction *isect, bool testAlphaTexture = true) const;
    virtual bool IntersectP(const Ray &ray, bool testAlphaTexture = true) const;
    virtual Float Area() const = 0;
    const Transform *ObjectToWorld, *WorldToObject;
    const bool reverseOrientation;
    const bool transformSwapsHandedness;
};

struct TriangleMesh {
    TriangleMesh(const Transform &ObjectToWorld, int nTriangles, const int *vertexIndices, int nVertices,
                 const Point3f *P, const Vector3f *S, const Normal3f *N, const Point2f *UV);
    const int nTriangles, nVertices;
    std::vector<int> vertexIndices;
    std::vector<Point3f> p;   // world space (triangle.cpp:73-74)
    std::vector<Normal3f> n;  // empty if absent
    std::vector<Vector3f> s;
    std::vector<Point2f> uv;
    std::shared_ptr<ImageTexture> alphaMask, shadowAlphaMask;   // triangle.h:61: one-channel textures, or null
};

class Triangle : public Shape {
  public:
    Triangle(const Transform *o2w, const Transform *w2o, bool reverseOrientation,
             const std::shared_ptr<TriangleMesh> &mesh, int triNumber)
        : Shape(o2w, w2o, reverseOrientation), mesh(mesh), triNumber(triNumber), v(&mesh->vertexIndices[3 * triNumber]) {}
    Bounds3f ObjectBound() const override;
    Bounds3f WorldBound() const override;
    Float Area() const override;
    std::shared_ptr<TriangleMesh> mesh;
    const int triNumber;
    const int *v;
};

std::vector<std::shared_ptr<Shape>> CreateTriangleMesh(const Transform *o2w, const Transform *w2o, bool reverseOrientation,
                                                       int nTriangles, const int *vertexIndices, int nVertices,
                                                       const Point3f *p, const Vector3f *s, const Normal3f *n,
                                                       const Point2f *uv);
std::vector<std::shared_ptr<Shape>> CreateTriangleMeshShape(const Transform *o2w, const Transform *w2o,
                                                            bool reverseOrientation, const ParamSet &params);
std::vector<std::shared_ptr<Shape>> CreateLoopSubdiv(const Transform *o2w, const Transform *w2o,
                                                     bool reverseOrientation, const ParamSet &params);
std::vector<std::shared_ptr<Shape>> CreatePLYMesh(const Transform *o2w, const Transform *w2o,
                                                  bool reverseOrientation, const ParamSet &params);
// Loop subdivision core, exposed for tests: returns positions/normals/indices of the limit mesh.
void LoopSubdivide(int nLevels, int nIndices, const int *vertexIndices, int nVertices, const Point3f *p,
                   std::vector<Point3f> *pLimit, std::vector<Normal3f> *Ns, std::vector<int> *indices);
bool ReadPLY(const std::string &filename, std::vector<Point3f> *P, std::vector<Normal3f> *N,
             std::vector<Point2f> *UV, std::vector<int> *indices);
bool WritePLY(const std::string &filename, const std::vector<Point3f> &P, const std::vector<int> &indices);

class Sphere : public Shape {
  public:
    Sphere(const Transform *o2w, const Transform *w2o, bool reverseOrientation, Float radius, Float zMin,
           Float zMax, Float phiMax);
    Bounds3f ObjectBound() const override;
    Float Area() const override { return phiMax * radius * (zMax - zMin); }
    const Float radius, zMin, zMax, thetaMin, thetaMax, phiMax;
};
std::shared_ptr<Shape> CreateSphereShape(const Transform *o2w, const Transform *w2o, bool reverseOrientation,
                                         const ParamSet &params);

// ---------------------------------------------------------------- materials / lights
class Material {
  public:
    virtual ~Material() {}
    virtual pb2_material Record() const = 0;  // constant-texture parameters for the device BSDF
    // image textures that replace constants of the record (PB2_TEX_* slots); null = the constant
    std::shared_ptr<ImageTexture> tex[PB2_TEX_SLOTS];
};
bool ReadImage(const std::string &name, std::vector<float> *rgb, int *w, int *h);   // RGB per pixel, row 0 at the top
std::shared_ptr<ImageTexture> CreateImageTexture(const TextureParams &tp, bool spectrum);
std::shared_ptr<ImageTexture> ConstantFloatImage(Float v);
class MatteMaterial : public Material {
  public:
    MatteMaterial(const Spectrum &Kd, Float sigma) : Kd(Kd), sigma(sigma) {}
    pb2_material Record() const override;
    Spectrum Kd;
    Float sigma;
};
class PlasticMaterial : public Material {
  public:
    PlasticMaterial(const Spectrum &Kd, const Spectrum &Ks, Float roughness, bool remapRoughness)
        : Kd(Kd), Ks(Ks), roughness(roughness), remapRoughness(remapRoughness) {}
    pb2_material Record() const override;
    Spectrum Kd, Ks;
    Float roughness;
    bool remapRoughness;
};
// mirror.h:49-66 / glass.h:49-77 with constant textures
class MirrorMaterial : public Material {
  public:
    explicit MirrorMaterial(const Spectrum &Kr) : Kr(Kr) {}
    pb2_material Record() const override;
    Spectrum Kr;
};
class GlassMaterial : public Material {
  public:
    GlassMaterial(const Spectrum &Kr, const Spectrum &Kt, Float uRoughness, Float vRoughness, Float index, bool remapRoughness)
        : Kr(Kr), Kt(Kt), uRoughness(uRoughness), vRoughness(vRoughness), index(index), remapRoughness(remapRoughness) {}
    pb2_material Record() const override;
    Spectrum Kr, Kt;
    Float uRoughness, vRoughness, index;
    bool remapRoughness;
};
// substrate.h:49-72 with constant textures
class SubstrateMaterial : public Material {
  public:
    SubstrateMaterial(const Spectrum &Kd, const Spectrum &Ks, Float nu, Float nv, bool remapRoughness)
        : Kd(Kd), Ks(Ks), nu(nu), nv(nv), remapRoughness(remapRoughness) {}
    pb2_material Record() const override;
    Spectrum Kd, Ks;
    Float nu, nv;
    bool remapRoughness;
};
// uber.h:49-85 / metal.h:49-76 with constant textures.  The "uroughness" / "vroughness" fall-backs of
// ComputeScatteringFunctions (uber.cpp:71-80, metal.cpp:67-70) are resolved at creation.
class UberMaterial : public Material {
  public:
    UberMaterial(const Spectrum &Kd, const Spectrum &Ks, const Spectrum &Kr, const Spectrum &Kt, Float roughnessu, Float roughnessv,
                 const Spectrum &opacity, Float eta, bool remapRoughness)
        : Kd(Kd), Ks(Ks), Kr(Kr), Kt(Kt), opacity(opacity), roughnessu(roughnessu), roughnessv(roughnessv), eta(eta), remapRoughness(remapRoughness) {}
    pb2_material Record() const override;
    Spectrum Kd, Ks, Kr, Kt, opacity;
    Float roughnessu, roughnessv, eta;
    bool remapRoughness;
};
class MetalMaterial : public Material {
  public:
    MetalMaterial(const Spectrum &eta, const Spectrum &k, Float uRoughness, Float vRoughness, bool remapRoughness)
        : eta(eta), k(k), uRoughness(uRoughness), vRoughness(vRoughness), remapRoughness(remapRoughness) {}
    pb2_material Record() const override;
    Spectrum eta, k;
    Float uRoughness, vRoughness;
    bool remapRoughness;
};
UberMaterial *CreateUberMaterial(const TextureParams &mp);
MetalMaterial *CreateMetalMaterial(const TextureParams &mp);
SubstrateMaterial *CreateSubstrateMaterial(const TextureParams &mp);
MirrorMaterial *CreateMirrorMaterial(const TextureParams &mp);
GlassMaterial *CreateGlassMaterial(const TextureParams &mp);
MatteMaterial *CreateMatteMaterial(const TextureParams &mp);
PlasticMaterial *CreatePlasticMaterial(const TextureParams &mp);

class Light {
  public:
    virtual ~Light() {}
};
class AreaLight : public Light {};
// point.h:49-71, spot.h:49-76, distant.h:49-72: the delta lights, as the parameters their Sample_Li reads
class PointLight : public Light {
  public:
    PointLight(const Transform &LightToWorld, const Spectrum &I) : pLight(LightToWorld(Point3f(0, 0, 0))), I(I) {}
    const Point3f pLight;
    const Spectrum I;
};
class SpotLight : public Light {
  public:
    SpotLight(const Transform &LightToWorld, const Spectrum &I, Float totalWidth, Float falloffStart)
        : pLight(LightToWorld(Point3f(0, 0, 0))), I(I), WorldToLight(Inverse(LightToWorld)), totalWidth(totalWidth),
          falloffStart(falloffStart), cosTotalWidth(std::cos(Radians(totalWidth))), cosFalloffStart(std::cos(Radians(falloffStart))) {}
    const Point3f pLight;
    const Spectrum I;
    const Transform WorldToLight;
    const Float totalWidth, falloffStart;   // degrees, as given to the constructor
    const Float cosTotalWidth, cosFalloffStart;
};
class DistantLight : public Light {
  public:
    DistantLight(const Transform &LightToWorld, const Spectrum &L, const Vector3f &w)
        : L(L), wWorld(LightToWorld.ApplyVector(w)), wLight(Normalize(wWorld)) {}
    const Spectrum L;
    const Vector3f wWorld;   // LightToWorld(w) before normalisation
    const Vector3f wLight;
};
// infinite.h:49-83 without a texture map: constant radiance from every direction
class InfiniteAreaLight : public Light {
  public:
    InfiniteAreaLight(const Transform &LightToWorld, const Spectrum &L, std::shared_ptr<ImageTexture> envMap = nullptr)
        : LightToWorld(LightToWorld), WorldToLight(Inverse(LightToWorld)), L(L), envMap(std::move(envMap)) {}
    const Transform LightToWorld, WorldToLight;
    const Spectrum L;
    // "mapname": the texels ReadImage returned, times L (infinite.cpp:50-57), as the constructor hands them to Lmap; null = constant
    const std::shared_ptr<ImageTexture> envMap;
};
std::shared_ptr<InfiniteAreaLight> CreateInfiniteLight(const Transform &light2world, const ParamSet &paramSet);
std::shared_ptr<PointLight> CreatePointLight(const Transform &light2world, const ParamSet &paramSet);
std::shared_ptr<SpotLight> CreateSpotLight(const Transform &light2world, const ParamSet &paramSet);
std::shared_ptr<DistantLight> CreateDistantLight(const Transform &light2world, const ParamSet &paramSet);
class DiffuseAreaLight : public AreaLight {
  public:
    DiffuseAreaLight(const Transform &LightToWorld, const Spectrum &Lemit, int nSamples,
                     const std::shared_ptr<Shape> &shape, bool twoSided);
    Spectrum Lemit;
    std::shared_ptr<Shape> shape;
    bool twoSided;
    Float area;
};
std::shared_ptr<AreaLight> CreateDiffuseAreaLight(const Transform &light2world, const ParamSet &paramSet,
                                                  const std::shared_ptr<Shape> &shape);

// ---------------------------------------------------------------- primitives / aggregates
class Primitive {
  public:
    virtual ~Primitive() {}
    virtual Bounds3f WorldBound() const = 0;
    virtual bool Intersect(const Ray &r, SurfaceInteraction *) const = 0;
    virtual bool IntersectP(const Ray &r) const = 0;
    virtual const AreaLight *GetAreaLight() const = 0;
    virtual const Material *GetMaterial() const = 0;
};

class GeometricPrimitive : public Primitive {
  public:
    GeometricPrimitive(const std::shared_ptr<Shape> &shape, const std::shared_ptr<Material> &material,
                       const std::shared_ptr<AreaLight> &areaLight)
        : shape(shape), material(material), areaLight(areaLight) {}
    Bounds3f WorldBound() const override { return shape->WorldBound(); }
    bool Intersect(const Ray &r, SurfaceInteraction *isect) const override;
    bool IntersectP(const Ray &r) const override;
    const AreaLight *GetAreaLight() const override { return areaLight.get(); }
    const Material *GetMaterial() const override { return material.get(); }
    std::shared_ptr<Shape> shape;
    std::shared_ptr<Material> material;
    std::shared_ptr<AreaLight> areaLight;
};

// primitive.h:93-116 with a static transform (the reference holds an AnimatedTransform; animated
// instance transforms are outside this path and refused by pbrtObjectInstance).
class BVHAccel;
class TransformedPrimitive : public Primitive {
  public:
    TransformedPrimitive(std::shared_ptr<Primitive> primitive, const Transform &PrimitiveToWorld)
        : primitive(std::move(primitive)), PrimitiveToWorld(PrimitiveToWorld) {}
    Bounds3f WorldBound() const override { return PrimitiveToWorld(primitive->WorldBound()); }  // MotionBounds, not animated
    bool Intersect(const Ray &r, SurfaceInteraction *isect) const override;
    bool IntersectP(const Ray &r) const override;
    const AreaLight *GetAreaLight() const override { return nullptr; }
    const Material *GetMaterial() const override { return nullptr; }
    std::shared_ptr<Primitive> primitive;
    const Transform PrimitiveToWorld;
  private:
    mutable std::shared_ptr<BVHAccel> single;  // one-primitive aggregate behind the direct Intersect calls
};

class Aggregate : public Primitive {
  public:
    const AreaLight *GetAreaLight() const override;
    const Material *GetMaterial() const override;
};

class BVHAccel : public Aggregate {
  public:
    enum class SplitMethod { SAH, HLBVH, Middle, EqualCounts };
    BVHAccel(std::vector<std::shared_ptr<Primitive>> p, int maxPrimsInNode = 1,
             SplitMethod splitMethod = SplitMethod::SAH, bool deviceBuild = false);
    ~BVHAccel();
    Bounds3f WorldBound() const override;
    bool Intersect(const Ray &ray, SurfaceInteraction *isect) const override;
    bool IntersectP(const Ray &ray) const override;

    // Host-built tree, uploaded verbatim.
    std::vector<pb2_bvh_node> nodes;
    std::vector<std::shared_ptr<Primitive>> primitives;         // BVH order (BVHAccel::primitives)
    std::vector<std::shared_ptr<Primitive>> sceneOrderPrims;    // as handed to the constructor
    std::vector<int32_t> orderedPrimNumbers;                    // primitives[j] == sceneOrderPrims[orderedPrimNumbers[j]]
    const int maxPrimsInNode;
    const SplitMethod splitMethod;
    const bool deviceBuild;           // HLBVH treelets built by pb2_hlbvh_treelets
    double lastBuildDeviceMs = 0;     // its device time
    mutable std::shared_ptr<DeviceScene> device;  // lazily created by Intersect/IntersectP/Render
};
std::shared_ptr<BVHAccel> CreateBVHAccelerator(std::vector<std::shared_ptr<Primitive>> prims, const ParamSet &ps);

class Scene {
  public:
    Scene(std::shared_ptr<Primitive> aggregate, const std::vector<std::shared_ptr<Light>> &lights)
        : lights(lights), aggregate(aggregate) { worldBound = aggregate->WorldBound(); }
    const Bounds3f &WorldBound() const { return worldBound; }
    bool Intersect(const Ray &ray, SurfaceInteraction *isect) const { return aggregate->Intersect(ray, isect); }
    bool IntersectP(const Ray &ray) const { return aggregate->IntersectP(ray); }
    std::vector<std::shared_ptr<Light>> lights;
    std::shared_ptr<Primitive> aggregate;
  private:
    Bounds3f worldBound;
};

// Flattens scene.aggregate (must be a BVHAccel of GeometricPrimitives) + scene.lights into a
// pb2_scene_desc.  The returned object owns every array the desc points into.
struct FlatScene {
    pb2_scene_desc desc;
    std::vector<float> P, N, UV, S;
    std::vector<int32_t> triIndex, triMesh;
    std::vector<pb2_mesh> meshes;
    std::vector<pb2_sphere> spheres;
    std::vector<uint8_t> primType;
    std::vector<int32_t> primIndex, primMaterial, primLight;
    std::vector<pb2_material> materials;
    std::vector<pb2_light> lights;
    std::vector<pb2_delta_light> deltaLights;   // empty when every light is an area light
    // object instancing: every BVH's nodes / ordered primitive numbers concatenated (scene BVH first)
    std::vector<pb2_bvh_node> nodes;
    std::vector<int32_t> bvhPrims;
    std::vector<pb2_bvh> bvhs;
    std::vector<pb2_instance> instances;
    std::vector<const Primitive *> primObjects;   // primitive number -> object (GeometricPrimitive or TransformedPrimitive)
    std::vector<pb2_texture> textures;            // image textures named by materials / meshes (1-based there)
    std::vector<std::shared_ptr<ImageTexture>> textureObjects;   // keeps the texel arrays alive
};
std::unique_ptr<FlatScene> FlattenScene(const BVHAccel &bvh, const std::vector<std::shared_ptr<Light>> &lights,
                                        const std::string &lightStrategy);
// Creates (once) the device copy of a BVHAccel-rooted scene; returns nullptr and reports through
// Error() when the CUDA library refuses (no device, unsupported feature).
bool EnsureDevice();   // pb2_init once per process (PB2_DEVICE / LOCAL_RANK pick the GPU); false + Error() without one
std::shared_ptr<DeviceScene> GetDeviceScene(const BVHAccel &bvh, const std::vector<std::shared_ptr<Light>> &lights,
                                            const std::string &lightStrategy);
pb2_scene *DeviceSceneHandle(const DeviceScene &);

// ---------------------------------------------------------------- film / camera / sampler
// filter.h:50-62 and src/filters/*.h: the host keeps radius and parameters; the 16x16 weight table of the
// Film (film.cpp:68-77) is computed where it is used, inside the CUDA library.
class Filter {
  public:
    Filter(Float xr, Float yr, int type = PB2_FILTER_BOX, Float p0 = 0, Float p1 = 0) : radius{xr, yr}, type(type), param{p0, p1} {}
    virtual ~Filter() {}
    Float radius[2];
    const int type;        // PB2_FILTER_*
    const Float param[2];  // gaussian: alpha; mitchell: B, C; sinc: tau
};
class BoxFilter : public Filter {
  public:
    BoxFilter(Float xr, Float yr) : Filter(xr, yr) {}
};
class GaussianFilter : public Filter {
  public:
    GaussianFilter(Float xr, Float yr, Float alpha) : Filter(xr, yr, PB2_FILTER_GAUSSIAN, alpha) {}
};
class MitchellFilter : public Filter {
  public:
    MitchellFilter(Float xr, Float yr, Float B, Float C) : Filter(xr, yr, PB2_FILTER_MITCHELL, B, C) {}
};
class LanczosSincFilter : public Filter {
  public:
    LanczosSincFilter(Float xr, Float yr, Float tau) : Filter(xr, yr, PB2_FILTER_SINC, tau) {}
};
class TriangleFilter : public Filter {
  public:
    TriangleFilter(Float xr, Float yr) : Filter(xr, yr, PB2_FILTER_TRIANGLE) {}
};
BoxFilter *CreateBoxFilter(const ParamSet &ps);
GaussianFilter *CreateGaussianFilter(const ParamSet &ps);
MitchellFilter *CreateMitchellFilter(const ParamSet &ps);
LanczosSincFilter *CreateSincFilter(const ParamSet &ps);
TriangleFilter *CreateTriangleFilter(const ParamSet &ps);

class Film {
  public:
    Film(const Point2i &resolution, const Bounds2f &cropWindow, std::unique_ptr<Filter> filter, Float diagonal,
         const std::string &filename, Float scale, Float maxSampleLuminance);
    Bounds2i GetSampleBounds() const;
    // MergeFilmTile for the whole film at once: rgbw = per-pixel (contribSum RGB, filterWeightSum)
    // as produced by pb2_render_path; converts to XYZ like film.cpp:125-128.
    void MergeDeviceFilm(const float *rgbw);
    void WriteImage(Float splatScale = 1);
    // Final RGB (film.cpp:178-203) without writing a file; 3 floats per cropped pixel.
    std::vector<Float> ResolveRGB() const;
    pb2_film_desc Desc() const;
    const Point2i fullResolution;
    const Float diagonal;
    std::unique_ptr<Filter> filter;
    const std::string filename;
    Bounds2i croppedPixelBounds;
    struct Pixel { Float xyz[3] = {0, 0, 0}; Float filterWeightSum = 0; };
    std::vector<Pixel> pixels;
    const Float scale;
    const Float maxSampleLuminance;
};
Film *CreateFilm(const ParamSet &params, std::unique_ptr<Filter> filter);
bool WriteImagePFM(const std::string &filename, const Float *rgb, int width, int height);
bool WriteImage(const std::string &name, const Float *rgb, int xRes, int yRes, int totalXRes, int totalYRes, int xOffset, int yOffset);
bool ReadImagePFM(const std::string &filename, std::vector<Float> *rgb, int *width, int *height);

class Camera {
  public:
    Camera(const Transform &CameraToWorld, Float shutterOpen, Float shutterClose, Film *film)
        : CameraToWorld(CameraToWorld), shutterOpen(shutterOpen), shutterClose(shutterClose), film(film) {}
    virtual ~Camera() {}
    virtual pb2_camera Desc() const = 0;
    Transform CameraToWorld;
    const Float shutterOpen, shutterClose;
    Film *film;
};
class PerspectiveCamera : public Camera {
  public:
    PerspectiveCamera(const Transform &CameraToWorld, const Bounds2f &screenWindow, Float shutterOpen,
                      Float shutterClose, Float lensRadius, Float focalDistance, Float fov, Film *film);
    pb2_camera Desc() const override;
    Transform CameraToScreen, RasterToCamera, ScreenToRaster, RasterToScreen;
    Bounds2f screenWindow;
    Float lensRadius, focalDistance, fov;
    Vector3f dxCamera, dyCamera;
};
PerspectiveCamera *CreatePerspectiveCamera(const ParamSet &params, const Transform &cam2world, Film *film);

class Sampler {
  public:
    explicit Sampler(int64_t spp) : samplesPerPixel(spp) {}
    virtual ~Sampler() {}
    const int64_t samplesPerPixel;
};
class HaltonSampler : public Sampler {
  public:
    HaltonSampler(int nsamp, const Bounds2i &sampleBounds, bool sampleAtCenter = false)
        : Sampler(nsamp), sampleBounds(sampleBounds), sampleAtPixelCenter(sampleAtCenter) {}
    Bounds2i sampleBounds;
    bool sampleAtPixelCenter;
};
HaltonSampler *CreateHaltonSampler(const ParamSet &params, const Bounds2i &sampleBounds);
// sobol.h:47-77: the sample count is rounded up to a power of two
class SobolSampler : public Sampler {
  public:
    SobolSampler(int64_t nsamp, const Bounds2i &sampleBounds);
    Bounds2i sampleBounds;
};
SobolSampler *CreateSobolSampler(const ParamSet &params, const Bounds2i &sampleBounds);

// ---------------------------------------------------------------- integrators
class Integrator {
  public:
    virtual ~Integrator() {}
    virtual void Render(const Scene &scene) = 0;
};

class SamplerIntegrator : public Integrator {
  public:
    SamplerIntegrator(std::shared_ptr<const Camera> camera, std::shared_ptr<Sampler> sampler, const Bounds2i &pixelBounds)
        : camera(camera), sampler(sampler), pixelBounds(pixelBounds) {}
    virtual void Preprocess(const Scene &scene, Sampler &sampler) {}
    void Render(const Scene &scene) override = 0;
  protected:
    std::shared_ptr<const Camera> camera;
    std::shared_ptr<Sampler> sampler;
    const Bounds2i pixelBounds;
};

class PathIntegrator : public SamplerIntegrator {
  public:
    PathIntegrator(int maxDepth, std::shared_ptr<const Camera> camera, std::shared_ptr<Sampler> sampler,
                   const Bounds2i &pixelBounds, Float rrThreshold = 1, const std::string &lightSampleStrategy = "spatial")
        : SamplerIntegrator(camera, sampler, pixelBounds), maxDepth(maxDepth), rrThreshold(rrThreshold),
          lightSampleStrategy(lightSampleStrategy) {}
    void Preprocess(const Scene &scene, Sampler &sampler) override;
    // The whole of SamplerIntegrator::Render + PathIntegrator::Li runs as one blocking device call.
    void Render(const Scene &scene) override;
    pb2_path_params Params() const;
    const std::string &LightSampleStrategy() const { return lightSampleStrategy; }
    pb2_stats lastStats{};
    bool writeImage = true;
    // 0 = the library's own partition: all of the film on one device, or the film's tiles dealt to the bound device group /
    // the ranks of the communicator (include/pb2.h: pb2_init_devices, pb2_dist_init) with the films merged on the first
    int tileRank = 0, tileCount = 0;
  private:
    const int maxDepth;
    const Float rrThreshold;
    const std::string lightSampleStrategy;
};
PathIntegrator *CreatePathIntegrator(const ParamSet &params, std::shared_ptr<Sampler> sampler,
                                     std::shared_ptr<const Camera> camera);

}  // namespace pbrt
#endif
