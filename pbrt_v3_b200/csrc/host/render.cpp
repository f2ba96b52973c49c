// Host-side render plumbing around the GPU hot path: materials/lights as parameter records,
// film (merge + XYZ round trip + PFM), perspective camera matrices, sampler/integrator factories,
// and the adapters that turn Shape/Primitive/Aggregate/Integrator calls into C-ABI calls.
//
//   Film ctor / GetSampleBounds / MergeFilmTile / WriteImage   src/core/film.cpp:45-211
//   PerspectiveCamera / ProjectiveCamera                       src/cameras/perspective.cpp:45-67,227-273, src/core/camera.h:84-115
//   CreateHaltonSampler / CreatePathIntegrator                 src/samplers/halton.cpp:133-139, src/integrators/path.cpp:190-213
//   Matte / Plastic / DiffuseAreaLight factories               src/materials/matte.cpp:64-71, plastic.cpp:72-83, src/lights/diffuse.cpp:135-148
#include <mutex>

#include "scene.h"

namespace pbrt {

// ---------------------------------------------------------------- materials / lights
static void clampSpectrum(const Spectrum &s, float out[3]) {
    for (int i = 0; i < 3; ++i) out[i] = s.c[i];
}
pb2_material MatteMaterial::Record() const {
    pb2_material m;
    std::memset(&m, 0, sizeof(m));
    m.type = PB2_MAT_MATTE;
    clampSpectrum(Kd, m.kd);
    m.sigma = sigma;
    return m;
}
pb2_material PlasticMaterial::Record() const {
    pb2_material m;
    std::memset(&m, 0, sizeof(m));
    m.type = PB2_MAT_PLASTIC;
    clampSpectrum(Kd, m.kd);
    clampSpectrum(Ks, m.ks);
    m.roughness = roughness;
    m.remap_roughness = remapRoughness ? 1 : 0;
    return m;
}
pb2_material MirrorMaterial::Record() const {
    pb2_material m;
    std::memset(&m, 0, sizeof(m));
    m.type = PB2_MAT_MIRROR;
    clampSpectrum(Kr, m.kr);
    return m;
}
pb2_material GlassMaterial::Record() const {
    pb2_material m;
    std::memset(&m, 0, sizeof(m));
    m.type = PB2_MAT_GLASS;
    clampSpectrum(Kr, m.kr);
    clampSpectrum(Kt, m.kt);
    m.eta = index;
    m.uroughness = uRoughness;
    m.vroughness = vRoughness;
    m.remap_roughness = remapRoughness ? 1 : 0;
    return m;
}
pb2_material SubstrateMaterial::Record() const {
    pb2_material m;
    std::memset(&m, 0, sizeof(m));
    m.type = PB2_MAT_SUBSTRATE;
    clampSpectrum(Kd, m.kd);
    clampSpectrum(Ks, m.ks);
    m.uroughness = nu;
    m.vroughness = nv;
    m.remap_roughness = remapRoughness ? 1 : 0;
    return m;
}
pb2_material UberMaterial::Record() const {
    pb2_material m;
    std::memset(&m, 0, sizeof(m));
    m.type = PB2_MAT_UBER;
    // the device clamps like uber.cpp:54-55, 64, 70, 92, 98 do; the record carries the parameters as given
    for (int c = 0; c < 3; ++c) {
        m.kd[c] = Kd.c[c];
        m.ks[c] = Ks.c[c];
        m.kr[c] = Kr.c[c];
        m.kt[c] = Kt.c[c];
        m.opacity[c] = opacity.c[c];
    }
    m.eta = eta;
    m.uroughness = roughnessu;
    m.vroughness = roughnessv;
    m.remap_roughness = remapRoughness ? 1 : 0;
    return m;
}
pb2_material MetalMaterial::Record() const {
    pb2_material m;
    std::memset(&m, 0, sizeof(m));
    m.type = PB2_MAT_METAL;
    for (int c = 0; c < 3; ++c) {
        m.metal_eta[c] = eta.c[c];
        m.metal_k[c] = k.c[c];
    }
    m.uroughness = uRoughness;
    m.vroughness = vRoughness;
    m.remap_roughness = remapRoughness ? 1 : 0;
    return m;
}
// point.cpp:84-92, spot.cpp:103-125, distant.cpp:92-100
std::shared_ptr<PointLight> CreatePointLight(const Transform &light2world, const ParamSet &paramSet) {
    Spectrum I = paramSet.FindOneSpectrum("I", Spectrum(1.0));
    Spectrum sc = paramSet.FindOneSpectrum("scale", Spectrum(1.0));
    Point3f P = paramSet.FindOnePoint3f("from", Point3f(0, 0, 0));
    Transform l2w = Translate(Vector3f(P.x, P.y, P.z)) * light2world;
    return std::make_shared<PointLight>(l2w, I * sc);
}
std::shared_ptr<SpotLight> CreateSpotLight(const Transform &l2w, const ParamSet &paramSet) {
    Spectrum I = paramSet.FindOneSpectrum("I", Spectrum(1.0));
    Spectrum sc = paramSet.FindOneSpectrum("scale", Spectrum(1.0));
    Float coneangle = paramSet.FindOneFloat("coneangle", 30.);
    Float conedelta = paramSet.FindOneFloat("conedeltaangle", 5.);
    Point3f from = paramSet.FindOnePoint3f("from", Point3f(0, 0, 0));
    Point3f to = paramSet.FindOnePoint3f("to", Point3f(0, 0, 1));
    Vector3f dir = Normalize(to - from);
    Vector3f du, dv;
    CoordinateSystem(dir, &du, &dv);
    Matrix4x4 rows;
    const Vector3f axes[3] = {du, dv, dir};
    for (int r = 0; r < 3; ++r) {
        rows.m[r][0] = axes[r].x;
        rows.m[r][1] = axes[r].y;
        rows.m[r][2] = axes[r].z;
        rows.m[r][3] = 0;
    }
    Transform dirToZ(rows);
    Transform light2world = l2w * Translate(Vector3f(from.x, from.y, from.z)) * Inverse(dirToZ);
    return std::make_shared<SpotLight>(light2world, I * sc, coneangle, coneangle - conedelta);
}
std::shared_ptr<DistantLight> CreateDistantLight(const Transform &light2world, const ParamSet &paramSet) {
    Spectrum L = paramSet.FindOneSpectrum("L", Spectrum(1.0));
    Spectrum sc = paramSet.FindOneSpectrum("scale", Spectrum(1.0));
    Point3f from = paramSet.FindOnePoint3f("from", Point3f(0, 0, 0));
    Point3f to = paramSet.FindOnePoint3f("to", Point3f(0, 0, 1));
    Vector3f dir = from - to;
    return std::make_shared<DistantLight>(light2world, L * sc, dir);
}
// CreateInfiniteLight (infinite.cpp:177-188).  "mapname": the image (PFM, PNG, TGA; texture.cpp ReadImage) times L becomes
// the light's radiance map; a file that cannot be read leaves the constant light, as in the reference (infinite.cpp:58-62).
extern std::string g_sceneDirectory;
std::shared_ptr<InfiniteAreaLight> CreateInfiniteLight(const Transform &light2world, const ParamSet &paramSet) {
    Spectrum L = paramSet.FindOneSpectrum("L", Spectrum(1.0));
    Spectrum sc = paramSet.FindOneSpectrum("scale", Spectrum(1.0));
    std::string texmap = paramSet.FindOneString("mapname", "");
    paramSet.FindOneInt("samples", paramSet.FindOneInt("nsamples", 1));
    L = L * sc;
    std::shared_ptr<ImageTexture> envMap;
    if (texmap != "") {
        std::string path = texmap;
        if (path[0] != '/' && !g_sceneDirectory.empty()) path = g_sceneDirectory + "/" + path;
        std::vector<float> rgb;
        int w = 0, h = 0;
        if (ReadImage(path, &rgb, &w, &h)) {
            envMap = std::make_shared<ImageTexture>();
            envMap->channels = 3;
            envMap->width = w;
            envMap->height = h;
            envMap->texels.resize(rgb.size());
            for (size_t i = 0; i < rgb.size(); ++i) envMap->texels[i] = rgb[i] * L.c[i % 3];   // texels[i] *= L (infinite.cpp:56)
        }
    }
    return std::make_shared<InfiniteAreaLight>(light2world, L, envMap);
}
// Textured parameters.  A parameter that names an image texture ("imagemap") of the right kind is attached to the
// material's slot (Material::tex -> pb2_material::tex); one that names anything else that varies (procedural textures,
// unknown names, a float texture where a spectrum is expected) is reported and keeps its default.
struct TexParam {
    const char *name;
    int slot;          // PB2_TEX_*
    bool spectrum;
};
static void attachTextures(Material *m, const TextureParams &mp, const char *what, std::initializer_list<TexParam> params) {
    for (const TexParam &tp : params) {
        if (auto t = mp.GetImageTexture(tp.name, tp.spectrum)) m->tex[tp.slot] = t;
        else if (mp.IsVaryingTexture(tp.name))
            Error("%s: the texture named by \"%s\" is outside the GPU path's scope (constant and \"imagemap\" textures with a "
                  "\"uv\" mapping, SURVEY.md §8 f.2); using the default value", what, tp.name);
    }
    // "bumpmap" (GetFloatTextureOrNull("bumpmap"), e.g. matte.cpp:70-71): an image texture, or a constant one as a 1 x 1 image
    Float bumpValue;
    if (auto t = mp.GetImageTexture("bumpmap", false)) m->tex[PB2_TEX_BUMP] = t;
    else if (mp.GetFloatOrNull("bumpmap", &bumpValue)) m->tex[PB2_TEX_BUMP] = ConstantFloatImage(bumpValue);
    else if (mp.NamedTexture("bumpmap") != "")
        Error("%s: the texture named by \"bumpmap\" is outside the GPU path's scope (constant and \"imagemap\" textures); ignored", what);
}
// A float parameter that may be absent, a constant or an image texture (GetFloatTextureOrNull, paramset.cpp:703-732)
struct FloatParam {
    bool present = false;
    std::shared_ptr<ImageTexture> tex;
    Float value = 0;
};
static FloatParam floatParam(const TextureParams &mp, const char *name) {
    FloatParam r;
    r.tex = mp.GetImageTexture(name, false);
    if (r.tex) r.present = true;
    else r.present = mp.GetFloatOrNull(name, &r.value);
    return r;
}
MatteMaterial *CreateMatteMaterial(const TextureParams &mp) {
    Spectrum Kd = mp.GetSpectrumTexture("Kd", Spectrum(0.5f));
    Float sigma = mp.GetFloatTexture("sigma", 0.f);
    auto *m = new MatteMaterial(Kd, sigma);
    attachTextures(m, mp, "matte", {{"Kd", PB2_TEX_KD, true}, {"sigma", PB2_TEX_SIGMA, false}});
    return m;
}
// substrate.cpp:67-81
SubstrateMaterial *CreateSubstrateMaterial(const TextureParams &mp) {
    Spectrum Kd = mp.GetSpectrumTexture("Kd", Spectrum(.5f));
    Spectrum Ks = mp.GetSpectrumTexture("Ks", Spectrum(.5f));
    Float uroughness = mp.GetFloatTexture("uroughness", .1f);
    Float vroughness = mp.GetFloatTexture("vroughness", .1f);
    bool remap = mp.FindBool("remaproughness", true);
    auto *m = new SubstrateMaterial(Kd, Ks, uroughness, vroughness, remap);
    attachTextures(m, mp, "substrate", {{"Kd", PB2_TEX_KD, true}, {"Ks", PB2_TEX_KS, true}, {"uroughness", PB2_TEX_UROUGHNESS, false},
                                        {"vroughness", PB2_TEX_VROUGHNESS, false}});
    return m;
}
// uber.cpp:106-131
UberMaterial *CreateUberMaterial(const TextureParams &mp) {
    Spectrum Kd = mp.GetSpectrumTexture("Kd", Spectrum(0.25f));
    Spectrum Ks = mp.GetSpectrumTexture("Ks", Spectrum(0.25f));
    Spectrum Kr = mp.GetSpectrumTexture("Kr", Spectrum(0.f));
    Spectrum Kt = mp.GetSpectrumTexture("Kt", Spectrum(0.f));
    Float roughness = mp.GetFloatTexture("roughness", .1f);
    // uber.cpp:71-80: roughu = "uroughness" if given, else "roughness"; roughv = "vroughness" if given, else roughu
    const FloatParam pu = floatParam(mp, "uroughness"), pv = floatParam(mp, "vroughness");
    Float roughu = pu.present ? pu.value : roughness;
    std::shared_ptr<ImageTexture> texU = pu.present ? pu.tex : mp.GetImageTexture("roughness", false);
    Float roughv = pv.present ? pv.value : roughu;
    std::shared_ptr<ImageTexture> texV = pv.present ? pv.tex : texU;
    const FloatParam pe = floatParam(mp, "eta");
    Float eta = pe.present ? pe.value : mp.GetFloatTexture("index", 1.5f);
    Spectrum opacity = mp.GetSpectrumTexture("opacity", Spectrum(1.f));
    bool remap = mp.FindBool("remaproughness", true);
    auto *m = new UberMaterial(Kd, Ks, Kr, Kt, roughu, roughv, opacity, eta, remap);
    attachTextures(m, mp, "uber", {{"Kd", PB2_TEX_KD, true}, {"Ks", PB2_TEX_KS, true}, {"Kr", PB2_TEX_KR, true}, {"Kt", PB2_TEX_KT, true},
                                   {"opacity", PB2_TEX_OPACITY, true}, {"roughness", PB2_TEX_ROUGHNESS, false},
                                   {"uroughness", PB2_TEX_UROUGHNESS, false}, {"vroughness", PB2_TEX_VROUGHNESS, false},
                                   {pe.present ? "eta" : "index", PB2_TEX_ETA, false}});
    m->tex[PB2_TEX_ROUGHNESS] = nullptr;   // resolved into the u / v slots
    m->tex[PB2_TEX_UROUGHNESS] = texU;
    m->tex[PB2_TEX_VROUGHNESS] = texV;
    return m;
}
// metal.cpp:120-140.  The defaults are copper's measured index and absorption converted to RGB by
// Spectrum::FromSampled (metal.cpp:82-118, spectrum.h:438-466); the six numbers below are that conversion's result,
// recorded from the compiled reference (tests/golden/metal_defaults.npz, tests/make_golden.py).
MetalMaterial *CreateMetalMaterial(const TextureParams &mp) {
    static const Spectrum copperN(0.19999069f, 0.92208463f, 1.09987593f), copperK(3.90463543f, 2.44763327f, 2.13765264f);
    Spectrum eta = mp.GetSpectrumTexture("eta", copperN);
    Spectrum k = mp.GetSpectrumTexture("k", copperK);
    Float roughness = mp.GetFloatTexture("roughness", .01f);
    // metal.cpp:67-70: each of the two falls back to "roughness"
    const FloatParam pu = floatParam(mp, "uroughness"), pv = floatParam(mp, "vroughness");
    const std::shared_ptr<ImageTexture> texR = mp.GetImageTexture("roughness", false);
    Float uRough = pu.present ? pu.value : roughness, vRough = pv.present ? pv.value : roughness;
    bool remap = mp.FindBool("remaproughness", true);
    auto *m = new MetalMaterial(eta, k, uRough, vRough, remap);
    attachTextures(m, mp, "metal", {{"eta", PB2_TEX_METAL_ETA, true}, {"k", PB2_TEX_METAL_K, true}, {"roughness", PB2_TEX_ROUGHNESS, false},
                                    {"uroughness", PB2_TEX_UROUGHNESS, false}, {"vroughness", PB2_TEX_VROUGHNESS, false}});
    m->tex[PB2_TEX_ROUGHNESS] = nullptr;
    m->tex[PB2_TEX_UROUGHNESS] = pu.present ? pu.tex : texR;
    m->tex[PB2_TEX_VROUGHNESS] = pv.present ? pv.tex : texR;
    return m;
}
// mirror.cpp:60-66
MirrorMaterial *CreateMirrorMaterial(const TextureParams &mp) {
    auto *m = new MirrorMaterial(mp.GetSpectrumTexture("Kr", Spectrum(0.9f)));
    attachTextures(m, mp, "mirror", {{"Kr", PB2_TEX_KR, true}});
    return m;
}
// glass.cpp:95-112
GlassMaterial *CreateGlassMaterial(const TextureParams &mp) {
    Spectrum Kr = mp.GetSpectrumTexture("Kr", Spectrum(1.f));
    Spectrum Kt = mp.GetSpectrumTexture("Kt", Spectrum(1.f));
    // "eta" wins over "index" when both are given (GetFloatTextureOrNull("eta") first)
    const FloatParam pe = floatParam(mp, "eta");
    Float eta = pe.present ? pe.value : mp.GetFloatTexture("index", 1.5f);
    Float roughu = mp.GetFloatTexture("uroughness", 0.f);
    Float roughv = mp.GetFloatTexture("vroughness", 0.f);
    bool remap = mp.FindBool("remaproughness", true);
    auto *m = new GlassMaterial(Kr, Kt, roughu, roughv, eta, remap);
    attachTextures(m, mp, "glass", {{"Kr", PB2_TEX_KR, true}, {"Kt", PB2_TEX_KT, true}, {"uroughness", PB2_TEX_UROUGHNESS, false},
                                    {"vroughness", PB2_TEX_VROUGHNESS, false}, {pe.present ? "eta" : "index", PB2_TEX_ETA, false}});
    return m;
}
PlasticMaterial *CreatePlasticMaterial(const TextureParams &mp) {
    Spectrum Kd = mp.GetSpectrumTexture("Kd", Spectrum(0.25f));
    Spectrum Ks = mp.GetSpectrumTexture("Ks", Spectrum(0.25f));
    Float roughness = mp.GetFloatTexture("roughness", .1f);
    bool remap = mp.FindBool("remaproughness", true);
    auto *m = new PlasticMaterial(Kd, Ks, roughness, remap);
    attachTextures(m, mp, "plastic", {{"Kd", PB2_TEX_KD, true}, {"Ks", PB2_TEX_KS, true}, {"roughness", PB2_TEX_ROUGHNESS, false}});
    return m;
}

DiffuseAreaLight::DiffuseAreaLight(const Transform &LightToWorld, const Spectrum &Lemit, int nSamples,
                                   const std::shared_ptr<Shape> &shape, bool twoSided)
    : Lemit(Lemit), shape(shape), twoSided(twoSided), area(shape->Area()) {
    if (Inverse(LightToWorld).HasScale() && dynamic_cast<const Triangle *>(shape.get()) == nullptr)
        Warning("Scaling detected in world to light transformation! The system has numerous assumptions, implicit and "
                "explicit, that this transform will have no scale factors in it. Proceed at your own risk; your image "
                "may have errors.");
}
std::shared_ptr<AreaLight> CreateDiffuseAreaLight(const Transform &light2world, const ParamSet &paramSet,
                                                  const std::shared_ptr<Shape> &shape) {
    Spectrum L = paramSet.FindOneSpectrum("L", Spectrum(1.0));
    Spectrum sc = paramSet.FindOneSpectrum("scale", Spectrum(1.0));
    int nSamples = paramSet.FindOneInt("samples", paramSet.FindOneInt("nsamples", 1));
    bool twoSided = paramSet.FindOneBool("twosided", false);
    return std::make_shared<DiffuseAreaLight>(light2world, L * sc, nSamples, shape, twoSided);
}

// ---------------------------------------------------------------- film
BoxFilter *CreateBoxFilter(const ParamSet &ps) {
    Float xw = ps.FindOneFloat("xwidth", 0.5f);
    Float yw = ps.FindOneFloat("ywidth", 0.5f);
    return new BoxFilter(xw, yw);
}
// gaussian.cpp:45-51, mitchell.cpp:45-52, sinc.cpp:45-50, triangle.cpp:46-51
GaussianFilter *CreateGaussianFilter(const ParamSet &ps) {
    Float xw = ps.FindOneFloat("xwidth", 2.f), yw = ps.FindOneFloat("ywidth", 2.f);
    return new GaussianFilter(xw, yw, ps.FindOneFloat("alpha", 2.f));
}
MitchellFilter *CreateMitchellFilter(const ParamSet &ps) {
    Float xw = ps.FindOneFloat("xwidth", 2.f), yw = ps.FindOneFloat("ywidth", 2.f);
    Float B = ps.FindOneFloat("B", 1.f / 3.f), C = ps.FindOneFloat("C", 1.f / 3.f);
    return new MitchellFilter(xw, yw, B, C);
}
LanczosSincFilter *CreateSincFilter(const ParamSet &ps) {
    Float xw = ps.FindOneFloat("xwidth", 4.f), yw = ps.FindOneFloat("ywidth", 4.f);
    return new LanczosSincFilter(xw, yw, ps.FindOneFloat("tau", 3.f));
}
TriangleFilter *CreateTriangleFilter(const ParamSet &ps) {
    Float xw = ps.FindOneFloat("xwidth", 2.f), yw = ps.FindOneFloat("ywidth", 2.f);
    return new TriangleFilter(xw, yw);
}

Film::Film(const Point2i &resolution, const Bounds2f &cropWindow, std::unique_ptr<Filter> filt, Float diagonal,
           const std::string &filename, Float scale, Float maxSampleLuminance)
    : fullResolution(resolution), diagonal(diagonal * .001), filter(std::move(filt)), filename(filename), scale(scale),
      maxSampleLuminance(maxSampleLuminance) {
    croppedPixelBounds = Bounds2i(Point2i((int)std::ceil(fullResolution.x * cropWindow.pMin.x),
                                          (int)std::ceil(fullResolution.y * cropWindow.pMin.y)),
                                  Point2i((int)std::ceil(fullResolution.x * cropWindow.pMax.x),
                                          (int)std::ceil(fullResolution.y * cropWindow.pMax.y)));
    pixels.resize(std::max(0, croppedPixelBounds.Area()));
}

Bounds2i Film::GetSampleBounds() const {
    Float x0 = std::floor(Float(croppedPixelBounds.pMin.x) + 0.5f - filter->radius[0]);
    Float y0 = std::floor(Float(croppedPixelBounds.pMin.y) + 0.5f - filter->radius[1]);
    Float x1 = std::ceil(Float(croppedPixelBounds.pMax.x) - 0.5f + filter->radius[0]);
    Float y1 = std::ceil(Float(croppedPixelBounds.pMax.y) - 0.5f + filter->radius[1]);
    return Bounds2i(Point2i((int)x0, (int)y0), Point2i((int)x1, (int)y1));
}

static inline void RGBToXYZ(const Float rgb[3], Float xyz[3]) {
    xyz[0] = 0.412453f * rgb[0] + 0.357580f * rgb[1] + 0.180423f * rgb[2];
    xyz[1] = 0.212671f * rgb[0] + 0.715160f * rgb[1] + 0.072169f * rgb[2];
    xyz[2] = 0.019334f * rgb[0] + 0.119193f * rgb[1] + 0.950227f * rgb[2];
}
static inline void XYZToRGB(const Float xyz[3], Float rgb[3]) {
    rgb[0] = 3.240479f * xyz[0] - 1.537150f * xyz[1] - 0.498535f * xyz[2];
    rgb[1] = -0.969256f * xyz[0] + 1.875991f * xyz[1] + 0.041556f * xyz[2];
    rgb[2] = 0.055648f * xyz[0] - 0.204043f * xyz[1] + 1.057311f * xyz[2];
}

void Film::MergeDeviceFilm(const float *rgbw) {
    size_t n = pixels.size();
    for (size_t i = 0; i < n; ++i) {
        Float xyz[3];
        RGBToXYZ(&rgbw[4 * i], xyz);
        for (int c = 0; c < 3; ++c) pixels[i].xyz[c] += xyz[c];
        pixels[i].filterWeightSum += rgbw[4 * i + 3];
    }
}

std::vector<Float> Film::ResolveRGB() const {
    std::vector<Float> rgb(3 * pixels.size());
    for (size_t i = 0; i < pixels.size(); ++i) {
        const Pixel &px = pixels[i];
        XYZToRGB(px.xyz, &rgb[3 * i]);
        Float w = px.filterWeightSum;
        if (w != 0) {
            Float invWt = (Float)1 / w;
            for (int c = 0; c < 3; ++c) rgb[3 * i + c] = std::max((Float)0, rgb[3 * i + c] * invWt);
        }
        // splats are BDPT/MLT only (film.cpp:142-167): always zero on this path, but the
        // reference still adds XYZToRGB(0) = 0 here, which cannot change a value.
        for (int c = 0; c < 3; ++c) rgb[3 * i + c] *= scale;
    }
    return rgb;
}

bool WriteImagePFM(const std::string &filename, const Float *rgb, int width, int height) {
    FILE *fp = std::fopen(filename.c_str(), "wb");
    if (!fp) {
        Error("Unable to open output PFM file \"%s\"", filename.c_str());
        return false;
    }
    std::fprintf(fp, "PF\n%d %d\n%f\n", width, height, -1.f);
    for (int y = height - 1; y >= 0; --y) std::fwrite(&rgb[(size_t)y * width * 3], sizeof(float), (size_t)width * 3, fp);
    std::fclose(fp);
    return true;
}

bool ReadImagePFM(const std::string &filename, std::vector<Float> *rgb, int *width, int *height) {
    FILE *fp = std::fopen(filename.c_str(), "rb");
    if (!fp) return false;
    char magic[8];
    float scale;
    if (std::fscanf(fp, "%7s %d %d %f", magic, width, height, &scale) != 4 || std::string(magic) != "PF" || scale > 0) {
        std::fclose(fp);
        return false;
    }
    std::fgetc(fp);
    rgb->resize((size_t)3 * *width * *height);
    bool ok = true;
    for (int y = *height - 1; y >= 0 && ok; --y)
        ok = std::fread(&(*rgb)[(size_t)y * *width * 3], sizeof(float), (size_t)*width * 3, fp) == (size_t)*width * 3;
    std::fclose(fp);
    return ok;
}

// ---------------------------------------------------------------- image writers (src/core/imageio.cpp:81-122)
// The reference hands EXR to OpenEXR (RgbaOutputFile, half RGB), PNG to lodepng and TGA to libtarga.  These
// writers produce the same pixel values in the same formats with the simplest legal encoding of each container:
// scan-line EXR without compression, PNG with stored deflate blocks, type-2 TGA.
namespace {
// float -> IEEE half, round to nearest even (what OpenEXR's half(float) does), overflow to infinity
uint16_t floatToHalf(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = (int32_t)((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));   // inf / nan
    if (exp >= 31) return (uint16_t)(sign | 0x7c00u);
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;              // rounds to zero
        man |= 0x800000u;                                  // denormal half: shift the implicit one in
        int shift = 14 - exp;
        uint32_t h = man >> shift, rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1))) ++h;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((uint32_t)exp << 10) | (man >> 13), rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;   // may carry into the exponent, up to infinity: still right
    return (uint16_t)(sign | h);
}
struct ByteSink {
    std::vector<uint8_t> b;
    void u8(uint8_t v) { b.push_back(v); }
    void le16(uint16_t v) { u8(v & 0xff); u8(v >> 8); }
    void le32(uint32_t v) { for (int i = 0; i < 4; ++i) u8((v >> (8 * i)) & 0xff); }
    void le64(uint64_t v) { for (int i = 0; i < 8; ++i) u8((v >> (8 * i)) & 0xff); }
    void be32(uint32_t v) { for (int i = 3; i >= 0; --i) u8((v >> (8 * i)) & 0xff); }
    void str(const char *s) { while (*s) u8((uint8_t)*s++); u8(0); }
    void raw(const void *p, size_t n) { const uint8_t *q = (const uint8_t *)p; b.insert(b.end(), q, q + n); }
};
bool writeAll(const std::string &filename, const std::vector<uint8_t> &bytes) {
    FILE *fp = std::fopen(filename.c_str(), "wb");
    if (!fp) {
        Error("Unable to open output file \"%s\"", filename.c_str());
        return false;
    }
    bool ok = std::fwrite(bytes.data(), 1, bytes.size(), fp) == bytes.size();
    std::fclose(fp);
    return ok;
}
// 8-bit formats: gamma (pbrt.h:293-296), then TO_BYTE of imageio.cpp:100
inline Float GammaCorrect(Float value) {
    if (value <= 0.0031308f) return 12.92f * value;
    return 1.055f * std::pow(value, (Float)(1.f / 2.4f)) - 0.055f;
}
std::vector<uint8_t> toRGB8(const Float *rgb, int w, int h) {
    std::vector<uint8_t> out((size_t)3 * w * h);
    for (size_t i = 0; i < out.size(); ++i) {
        Float v = 255.f * GammaCorrect(rgb[i]) + 0.5f;
        out[i] = (uint8_t)(v < 0.f ? 0.f : (v > 255.f ? 255.f : v));
    }
    return out;
}
uint32_t crc32(const uint8_t *p, size_t n, uint32_t crc = 0) {
    struct Table {
        uint32_t v[256];
        Table() {
            for (uint32_t i = 0; i < 256; ++i) {
                uint32_t c = i;
                for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
                v[i] = c;
            }
        }
    };
    static const Table table;   // initialised once, thread-safe
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table.v[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}
}  // namespace

// imageio.cpp:164-190: RGB half channels, display window = the full resolution, data window = the cropped pixels
bool WriteImageEXR(const std::string &filename, const Float *rgb, int xRes, int yRes, int totalXRes, int totalYRes, int xOffset, int yOffset) {
    ByteSink o;
    o.le32(20000630);   // magic
    o.le32(2);          // version 2, single-part scan-line file
    auto attr = [&](const char *name, const char *type, uint32_t size) { o.str(name); o.str(type); o.le32(size); };
    attr("channels", "chlist", 3 * 18 + 1);
    for (const char *ch : {"B", "G", "R"}) {   // alphabetical, as the format requires
        o.str(ch);
        o.le32(1);      // HALF
        o.u8(0); o.u8(0); o.u8(0); o.u8(0);   // pLinear + reserved
        o.le32(1); o.le32(1);                 // sampling
    }
    o.u8(0);
    attr("compression", "compression", 1); o.u8(0);   // NO_COMPRESSION
    attr("dataWindow", "box2i", 16);
    o.le32(xOffset); o.le32(yOffset); o.le32(xOffset + xRes - 1); o.le32(yOffset + yRes - 1);
    attr("displayWindow", "box2i", 16);
    o.le32(0); o.le32(0); o.le32(totalXRes - 1); o.le32(totalYRes - 1);
    attr("lineOrder", "lineOrder", 1); o.u8(0);       // INCREASING_Y
    float one = 1.f, zero = 0.f;
    attr("pixelAspectRatio", "float", 4); o.raw(&one, 4);
    attr("screenWindowCenter", "v2f", 8); o.raw(&zero, 4); o.raw(&zero, 4);
    attr("screenWindowWidth", "float", 4); o.raw(&one, 4);
    o.u8(0);            // end of header
    const uint64_t lineBytes = (uint64_t)xRes * 3 * 2, tableStart = o.b.size();
    for (int y = 0; y < yRes; ++y) o.le64(tableStart + 8ull * yRes + (uint64_t)y * (8 + lineBytes));
    std::vector<uint16_t> line((size_t)xRes * 3);
    for (int y = 0; y < yRes; ++y) {
        o.le32(yOffset + y);
        o.le32((uint32_t)lineBytes);
        for (int c = 0; c < 3; ++c)            // B, G, R planes of this scan line
            for (int x = 0; x < xRes; ++x) line[(size_t)c * xRes + x] = floatToHalf(rgb[3 * ((size_t)y * xRes + x) + (2 - c)]);
        for (uint16_t v : line) o.le16(v);
    }
    return writeAll(filename, o.b);
}
// imageio.cpp:93-117 (lodepng_encode24_file): 8-bit RGB, no interlace; zlib stream of stored blocks
bool WriteImagePNG(const std::string &filename, const Float *rgb, int w, int h) {
    std::vector<uint8_t> px = toRGB8(rgb, w, h), rawData;
    rawData.reserve((size_t)h * (3 * w + 1));
    for (int y = 0; y < h; ++y) {
        rawData.push_back(0);   // filter type None
        rawData.insert(rawData.end(), px.begin() + (size_t)y * 3 * w, px.begin() + (size_t)(y + 1) * 3 * w);
    }
    ByteSink z;
    z.u8(0x78); z.u8(0x01);
    size_t pos = 0;
    do {
        size_t n = std::min<size_t>(65535, rawData.size() - pos);
        z.u8(pos + n == rawData.size() ? 1 : 0);
        z.le16((uint16_t)n); z.le16((uint16_t)~n);
        z.raw(rawData.data() + pos, n);
        pos += n;
    } while (pos < rawData.size());
    uint32_t a = 1, b = 0;   // Adler-32
    for (uint8_t v : rawData) { a = (a + v) % 65521; b = (b + a) % 65521; }
    z.be32((b << 16) | a);
    ByteSink o;
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    o.raw(sig, 8);
    auto chunk = [&](const char *type, const std::vector<uint8_t> &data) {
        o.be32((uint32_t)data.size());
        size_t start = o.b.size();
        o.raw(type, 4);
        o.raw(data.data(), data.size());
        o.be32(crc32(&o.b[start], o.b.size() - start));
    };
    ByteSink ihdr;
    ihdr.be32(w); ihdr.be32(h); ihdr.u8(8); ihdr.u8(2); ihdr.u8(0); ihdr.u8(0); ihdr.u8(0);
    chunk("IHDR", ihdr.b);
    chunk("IDAT", z.b);
    chunk("IEND", {});
    return writeAll(filename, o.b);
}
// imageio.cpp:193-214 + targa.cpp:494-541: uncompressed 24-bit BGR, rows top to bottom
bool WriteImageTGA(const std::string &filename, const Float *rgb, int w, int h) {
    std::vector<uint8_t> px = toRGB8(rgb, w, h);
    ByteSink o;
    o.u8(0); o.u8(0); o.u8(2);                  // no id, no colour map, true-colour
    for (int i = 0; i < 5; ++i) o.u8(0);        // colour map specification
    o.le16(0); o.le16(0); o.le16((uint16_t)w); o.le16((uint16_t)h);
    o.u8(24); o.u8(0x20);                       // TGA_T_TO_B_BIT
    for (size_t i = 0; i < (size_t)w * h; ++i) { o.u8(px[3 * i + 2]); o.u8(px[3 * i + 1]); o.u8(px[3 * i]); }
    return writeAll(filename, o.b);
}

// WriteImage (imageio.cpp:81-122): the container follows the file name's extension
bool WriteImage(const std::string &name, const Float *rgb, int xRes, int yRes, int totalXRes, int totalYRes, int xOffset, int yOffset) {
    size_t dot = name.rfind('.');
    std::string ext = dot == std::string::npos ? "" : name.substr(dot);
    for (char &c : ext) c = (char)std::tolower((unsigned char)c);
    if (ext == ".exr") return WriteImageEXR(name, rgb, xRes, yRes, totalXRes, totalYRes, xOffset, yOffset);
    if (ext == ".pfm") return WriteImagePFM(name, rgb, xRes, yRes);
    if (ext == ".png") return WriteImagePNG(name, rgb, xRes, yRes);
    if (ext == ".tga") return WriteImageTGA(name, rgb, xRes, yRes);
    Error("Can't determine image file type from suffix of filename \"%s\"", name.c_str());
    return false;
}

void Film::WriteImage(Float splatScale) {
    std::vector<Float> rgb = ResolveRGB();
    int w = croppedPixelBounds.pMax.x - croppedPixelBounds.pMin.x;
    int h = croppedPixelBounds.pMax.y - croppedPixelBounds.pMin.y;
    pbrt::WriteImage(filename, rgb.data(), w, h, fullResolution.x, fullResolution.y, croppedPixelBounds.pMin.x, croppedPixelBounds.pMin.y);
}

pb2_film_desc Film::Desc() const {
    pb2_film_desc d;
    std::memset(&d, 0, sizeof(d));
    d.full_resolution[0] = fullResolution.x;
    d.full_resolution[1] = fullResolution.y;
    d.cropped_pixel_bounds[0] = croppedPixelBounds.pMin.x;
    d.cropped_pixel_bounds[1] = croppedPixelBounds.pMin.y;
    d.cropped_pixel_bounds[2] = croppedPixelBounds.pMax.x;
    d.cropped_pixel_bounds[3] = croppedPixelBounds.pMax.y;
    d.filter_radius[0] = filter->radius[0];
    d.filter_radius[1] = filter->radius[1];
    d.filter_type = filter->type;
    d.filter_param[0] = filter->param[0];
    d.filter_param[1] = filter->param[1];
    d.max_sample_luminance = maxSampleLuminance;
    d.scale = scale;
    return d;
}

extern std::string g_imageFileOverride;  // api.cpp (--outfile)
extern Float g_cropWindow[2][2];
extern bool g_quickRender;             // --quick
Film *CreateFilm(const ParamSet &params, std::unique_ptr<Filter> filter) {
    std::string filename;
    if (g_imageFileOverride != "") {
        filename = g_imageFileOverride;
        std::string paramsFilename = params.FindOneString("filename", "");
        if (paramsFilename != "")
            Warning("Output filename supplied on command line, \"%s\" is overriding filename provided in scene "
                    "description file, \"%s\".", g_imageFileOverride.c_str(), paramsFilename.c_str());
    } else
        filename = params.FindOneString("filename", "pbrt.exr");
    int xres = params.FindOneInt("xresolution", 1280);
    int yres = params.FindOneInt("yresolution", 720);
    if (g_quickRender) xres = std::max(1, xres / 4);   // film.cpp:228-229
    if (g_quickRender) yres = std::max(1, yres / 4);
    Bounds2f crop;
    bool haveCrop = false;
    std::vector<Float> cr = params.FindFloats(ParamSet::Type::Float, "cropwindow", &haveCrop);
    if (haveCrop && cr.size() == 4) {
        crop.pMin.x = Clamp(std::min(cr[0], cr[1]), 0.f, 1.f);
        crop.pMax.x = Clamp(std::max(cr[0], cr[1]), 0.f, 1.f);
        crop.pMin.y = Clamp(std::min(cr[2], cr[3]), 0.f, 1.f);
        crop.pMax.y = Clamp(std::max(cr[2], cr[3]), 0.f, 1.f);
    } else {
        if (haveCrop) Error("%d values supplied for \"cropwindow\". Expected 4.", (int)cr.size());
        crop.pMin = Point2f(Clamp(g_cropWindow[0][0], 0, 1), Clamp(g_cropWindow[1][0], 0, 1));
        crop.pMax = Point2f(Clamp(g_cropWindow[0][1], 0, 1), Clamp(g_cropWindow[1][1], 0, 1));
    }
    Float scale = params.FindOneFloat("scale", 1.);
    Float diagonal = params.FindOneFloat("diagonal", 35.);
    Float maxSampleLuminance = params.FindOneFloat("maxsampleluminance", Infinity);
    return new Film(Point2i(xres, yres), crop, std::move(filter), diagonal, filename, scale, maxSampleLuminance);
}

// ---------------------------------------------------------------- camera
PerspectiveCamera::PerspectiveCamera(const Transform &c2w, const Bounds2f &screenWindow, Float shutterOpen,
                                     Float shutterClose, Float lensr, Float focald, Float fov, Film *film)
    : Camera(c2w, shutterOpen, shutterClose, film), CameraToScreen(Perspective(fov, 1e-2f, 1000.f)),
      screenWindow(screenWindow), lensRadius(lensr), focalDistance(focald), fov(fov) {
    ScreenToRaster = Scale(film->fullResolution.x, film->fullResolution.y, 1) *
                     Scale(1 / (screenWindow.pMax.x - screenWindow.pMin.x), 1 / (screenWindow.pMin.y - screenWindow.pMax.y), 1) *
                     Translate(Vector3f(-screenWindow.pMin.x, -screenWindow.pMax.y, 0));
    RasterToScreen = Inverse(ScreenToRaster);
    RasterToCamera = Inverse(CameraToScreen) * RasterToScreen;
    dxCamera = RasterToCamera(Point3f(1, 0, 0)) - RasterToCamera(Point3f(0, 0, 0));
    dyCamera = RasterToCamera(Point3f(0, 1, 0)) - RasterToCamera(Point3f(0, 0, 0));
}

pb2_camera PerspectiveCamera::Desc() const {
    pb2_camera c;
    std::memset(&c, 0, sizeof(c));
    std::memcpy(c.camera_to_world, CameraToWorld.GetMatrix().m, sizeof(c.camera_to_world));
    std::memcpy(c.world_to_camera, CameraToWorld.GetInverseMatrix().m, sizeof(c.world_to_camera));
    c.screen_window[0] = screenWindow.pMin.x; c.screen_window[1] = screenWindow.pMax.x;
    c.screen_window[2] = screenWindow.pMin.y; c.screen_window[3] = screenWindow.pMax.y;
    c.fov = fov;
    c.lens_radius = lensRadius;
    c.focal_distance = focalDistance;
    c.shutter_open = shutterOpen;
    c.shutter_close = shutterClose;
    std::memcpy(c.raster_to_camera, RasterToCamera.GetMatrix().m, sizeof(c.raster_to_camera));
    for (int k = 0; k < 3; ++k) { c.dx_camera[k] = dxCamera[k]; c.dy_camera[k] = dyCamera[k]; }
    return c;
}

PerspectiveCamera *CreatePerspectiveCamera(const ParamSet &params, const Transform &cam2world, Film *film) {
    Float shutteropen = params.FindOneFloat("shutteropen", 0.f);
    Float shutterclose = params.FindOneFloat("shutterclose", 1.f);
    if (shutterclose < shutteropen) {
        Warning("Shutter close time [%f] < shutter open [%f].  Swapping them.", shutterclose, shutteropen);
        std::swap(shutterclose, shutteropen);
    }
    Float lensradius = params.FindOneFloat("lensradius", 0.f);
    Float focaldistance = params.FindOneFloat("focaldistance", 1e6);
    Float frame = params.FindOneFloat("frameaspectratio", Float(film->fullResolution.x) / Float(film->fullResolution.y));
    Bounds2f screen;
    if (frame > 1.f) {
        screen.pMin.x = -frame; screen.pMax.x = frame; screen.pMin.y = -1.f; screen.pMax.y = 1.f;
    } else {
        screen.pMin.x = -1.f; screen.pMax.x = 1.f; screen.pMin.y = -1.f / frame; screen.pMax.y = 1.f / frame;
    }
    bool haveSw = false;
    std::vector<Float> sw = params.FindFloats(ParamSet::Type::Float, "screenwindow", &haveSw);
    if (haveSw) {
        if (sw.size() == 4) {
            screen.pMin.x = sw[0]; screen.pMax.x = sw[1]; screen.pMin.y = sw[2]; screen.pMax.y = sw[3];
        } else
            Error("\"screenwindow\" should have four values");
    }
    Float fov = params.FindOneFloat("fov", 90.);
    Float halffov = params.FindOneFloat("halffov", -1.f);
    if (halffov > 0.f) fov = 2.f * halffov;
    return new PerspectiveCamera(cam2world, screen, shutteropen, shutterclose, lensradius, focaldistance, fov, film);
}

// ---------------------------------------------------------------- sampler / integrator factories
HaltonSampler *CreateHaltonSampler(const ParamSet &params, const Bounds2i &sampleBounds) {
    int nsamp = params.FindOneInt("pixelsamples", 16);
    if (g_quickRender) nsamp = 1;                      // halton.cpp:136
    bool sampleAtCenter = params.FindOneBool("samplepixelcenter", false);
    return new HaltonSampler(nsamp, sampleBounds, sampleAtCenter);
}

static int64_t roundUpPow2(int64_t v) {
    int64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}
SobolSampler::SobolSampler(int64_t nsamp, const Bounds2i &sampleBounds) : Sampler(roundUpPow2(nsamp)), sampleBounds(sampleBounds) {
    if (samplesPerPixel != nsamp)
        Warning("Non power-of-two sample count rounded up to %lld for SobolSampler.", (long long)samplesPerPixel);
}
// sobol.cpp:65-70
SobolSampler *CreateSobolSampler(const ParamSet &params, const Bounds2i &sampleBounds) {
    int nsamp = params.FindOneInt("pixelsamples", 16);
    if (g_quickRender) nsamp = 1;
    return new SobolSampler(nsamp, sampleBounds);
}

PathIntegrator *CreatePathIntegrator(const ParamSet &params, std::shared_ptr<Sampler> sampler,
                                     std::shared_ptr<const Camera> camera) {
    int maxDepth = params.FindOneInt("maxdepth", 5);
    Bounds2i pixelBounds = camera->film->GetSampleBounds();
    bool havePb = false;
    std::vector<int> pb = params.FindInts("pixelbounds", &havePb);
    if (havePb) {
        if (pb.size() != 4)
            Error("Expected four values for \"pixelbounds\" parameter. Got %d.", (int)pb.size());
        else {
            Bounds2i b(Point2i(std::max(pixelBounds.pMin.x, pb[0]), std::max(pixelBounds.pMin.y, pb[2])),
                       Point2i(std::min(pixelBounds.pMax.x, pb[1]), std::min(pixelBounds.pMax.y, pb[3])));
            pixelBounds = b;
            if (pixelBounds.Area() == 0) Error("Degenerate \"pixelbounds\" specified.");
        }
    }
    Float rrThreshold = params.FindOneFloat("rrthreshold", 1.);
    std::string lightStrategy = params.FindOneString("lightsamplestrategy", "spatial");
    return new PathIntegrator(maxDepth, camera, sampler, pixelBounds, rrThreshold, lightStrategy);
}

// ---------------------------------------------------------------- device adapters
struct DeviceScene {
    pb2_scene *handle = nullptr;
    std::unique_ptr<FlatScene> flat;
    std::vector<const Light *> lightKey;   // what the scene was flattened with (cache key of GetDeviceScene)
    std::string strategyKey;
    ~DeviceScene() {
        if (handle) pb2_scene_destroy(handle);
    }
};
pb2_scene *DeviceSceneHandle(const DeviceScene &d) { return d.handle; }

bool EnsureDevice() {
    static std::once_flag once;
    static int status = PB2_OK;
    std::call_once(once, [] {
        // a process that has bound its device(s) already keeps them (tests, bench.py: one process per GPU); PB2_DEVICE /
        // LOCAL_RANK name one device; otherwise - the pb2_pbrt command line - every visible GPU renders (film tiles dealt to
        // the devices, films merged on the first one)
        if (pb2_device_count() > 0) return;
        const char *dev = std::getenv("PB2_DEVICE");
        if (!dev) dev = std::getenv("LOCAL_RANK");
        status = dev ? pb2_init(std::atoi(dev)) : pb2_init_devices(0, nullptr);
    });
    if (status != PB2_OK) Error("pb2_init failed: %s (there is no CPU fallback for the path-tracing hot path)", pb2_last_error());
    return status == PB2_OK;
}

// The device copy is cached on the aggregate under the lights and the light-sampling strategy it was flattened with:
// a Render after a stand-alone BVHAccel::Intersect (which flattens without lights), or a second Render with another
// lightsamplestrategy, gets a scene of its own instead of the stale tables.
std::shared_ptr<DeviceScene> GetDeviceScene(const BVHAccel &bvh, const std::vector<std::shared_ptr<Light>> &lights,
                                            const std::string &lightStrategy) {
    std::vector<const Light *> key;
    for (const auto &l : lights) key.push_back(l.get());
    if (bvh.device && bvh.device->lightKey == key && bvh.device->strategyKey == lightStrategy) return bvh.device;
    if (!EnsureDevice()) return nullptr;
    auto ds = std::make_shared<DeviceScene>();
    ds->lightKey = key;
    ds->strategyKey = lightStrategy;
    ds->flat = FlattenScene(bvh, lights, lightStrategy);
    if (!ds->flat) return nullptr;
    if (pb2_scene_create(&ds->flat->desc, &ds->handle) != PB2_OK) {
        Error("pb2_scene_create failed: %s", pb2_last_error());
        return nullptr;
    }
    bvh.device = ds;
    return ds;
}

static void fillInteraction(const pb2_hit &h, const FlatScene &flat, const Ray &ray, SurfaceInteraction *isect) {
    isect->p = Point3f(h.p[0], h.p[1], h.p[2]);
    isect->pError = Vector3f(h.p_error[0], h.p_error[1], h.p_error[2]);
    isect->n = Normal3f(h.n[0], h.n[1], h.n[2]);
    isect->shading.n = Normal3f(h.ns[0], h.ns[1], h.ns[2]);
    isect->shading.dpdu = Vector3f(h.dpdu[0], h.dpdu[1], h.dpdu[2]);
    isect->uv = Point2f(h.uv[0], h.uv[1]);
    isect->wo = Normalize(-ray.d);
    for (int k = 0; k < 3; ++k) isect->b[k] = h.b[k];
    isect->primitive = flat.primObjects[h.prim];   // the GeometricPrimitive, also for a hit inside an instance
}

bool BVHAccel::Intersect(const Ray &ray, SurfaceInteraction *isect) const {
    if (nodes.empty()) return false;
    std::vector<std::shared_ptr<Light>> noLights;
    std::shared_ptr<DeviceScene> ds = device ? device : GetDeviceScene(*this, noLights, "uniform");   // any cached copy answers a geometric query
    if (!ds) return false;
    pb2_ray r = {{ray.o.x, ray.o.y, ray.o.z}, {ray.d.x, ray.d.y, ray.d.z}, ray.tMax};
    pb2_hit h;
    if (pb2_intersect(ds->handle, &r, 1, &h) != PB2_OK) {
        Error("pb2_intersect failed: %s", pb2_last_error());
        return false;
    }
    if (h.prim < 0) return false;
    ray.tMax = h.t;
    fillInteraction(h, *ds->flat, ray, isect);
    return true;
}

bool BVHAccel::IntersectP(const Ray &ray) const {
    if (nodes.empty()) return false;
    std::vector<std::shared_ptr<Light>> noLights;
    std::shared_ptr<DeviceScene> ds = device ? device : GetDeviceScene(*this, noLights, "uniform");   // any cached copy answers a geometric query
    if (!ds) return false;
    pb2_ray r = {{ray.o.x, ray.o.y, ray.o.z}, {ray.d.x, ray.d.y, ray.d.z}, ray.tMax};
    uint8_t occ = 0;
    if (pb2_intersect_p(ds->handle, &r, 1, &occ) != PB2_OK) {
        Error("pb2_intersect_p failed: %s", pb2_last_error());
        return false;
    }
    return occ != 0;
}

// A Shape (or GeometricPrimitive) asked to intersect on its own is wrapped in a one-primitive
// aggregate so that the same device kernels answer (no host intersection code exists).
namespace {
struct ShapeProxy : public Shape {
    // Non-owning view used only to build a shared_ptr<Shape> aliasing an existing shape.
    using Shape::Shape;
};
std::mutex g_singleMutex;
std::map<const Shape *, std::shared_ptr<BVHAccel>> g_singleShapeAccels;
std::shared_ptr<BVHAccel> singleShapeAccel(const Shape *shape) {
    std::lock_guard<std::mutex> lock(g_singleMutex);
    auto it = g_singleShapeAccels.find(shape);
    if (it != g_singleShapeAccels.end()) return it->second;
    std::shared_ptr<Shape> alias(std::shared_ptr<Shape>(), const_cast<Shape *>(shape));
    std::vector<std::shared_ptr<Primitive>> prims{std::make_shared<GeometricPrimitive>(alias, nullptr, nullptr)};
    auto accel = std::make_shared<BVHAccel>(std::move(prims), 1);
    g_singleShapeAccels[shape] = accel;
    return accel;
}
}  // namespace

bool Shape::Intersect(const Ray &ray, Float *tHit, SurfaceInteraction *isect, bool) const {
    Ray r = ray;
    if (!singleShapeAccel(this)->Intersect(r, isect)) return false;
    *tHit = r.tMax;
    return true;
}
bool Shape::IntersectP(const Ray &ray, bool) const { return singleShapeAccel(this)->IntersectP(ray); }

bool GeometricPrimitive::Intersect(const Ray &r, SurfaceInteraction *isect) const {
    Float tHit;
    if (!shape->Intersect(r, &tHit, isect)) return false;
    r.tMax = tHit;
    isect->primitive = this;
    return true;
}
bool GeometricPrimitive::IntersectP(const Ray &r) const { return shape->IntersectP(r); }

// TransformedPrimitive::Intersect[P] (primitive.cpp:76-96) run on the device like everything else:
// a one-primitive aggregate holding a copy of this instance.
bool TransformedPrimitive::Intersect(const Ray &r, SurfaceInteraction *isect) const {
    if (!single)
        single = std::make_shared<BVHAccel>(std::vector<std::shared_ptr<Primitive>>{std::make_shared<TransformedPrimitive>(primitive, PrimitiveToWorld)});
    return single->Intersect(r, isect);
}
bool TransformedPrimitive::IntersectP(const Ray &r) const {
    if (!single)
        single = std::make_shared<BVHAccel>(std::vector<std::shared_ptr<Primitive>>{std::make_shared<TransformedPrimitive>(primitive, PrimitiveToWorld)});
    return single->IntersectP(r);
}

// ---------------------------------------------------------------- PathIntegrator
void PathIntegrator::Preprocess(const Scene &, Sampler &) {}

pb2_path_params PathIntegrator::Params() const {
    pb2_path_params p;
    std::memset(&p, 0, sizeof(p));
    const HaltonSampler *hs = dynamic_cast<const HaltonSampler *>(sampler.get());
    p.samples_per_pixel = (int32_t)sampler->samplesPerPixel;
    p.sample_at_pixel_center = hs && hs->sampleAtPixelCenter;
    p.max_depth = maxDepth;
    p.rr_threshold = rrThreshold;
    p.pixel_bounds[0] = pixelBounds.pMin.x; p.pixel_bounds[1] = pixelBounds.pMin.y;
    p.pixel_bounds[2] = pixelBounds.pMax.x; p.pixel_bounds[3] = pixelBounds.pMax.y;
    p.tile_rank = tileRank;
    p.tile_count = tileCount;
    p.sampler = dynamic_cast<const SobolSampler *>(sampler.get()) ? PB2_SAMPLER_SOBOL : PB2_SAMPLER_HALTON;
    return p;
}

void PathIntegrator::Render(const Scene &scene) {
    Preprocess(scene, *sampler);
    const BVHAccel *bvh = dynamic_cast<const BVHAccel *>(scene.aggregate.get());
    if (!bvh) {
        Error("PathIntegrator::Render: the GPU path needs Accelerator \"bvh\" (kd-tree is out of scope, SURVEY.md §2 row 41)");
        return;
    }
    if (!dynamic_cast<const HaltonSampler *>(sampler.get()) && !dynamic_cast<const SobolSampler *>(sampler.get())) {
        Error("PathIntegrator::Render: only the GlobalSamplers \"halton\" and \"sobol\" are inside the GPU path's scope (SURVEY.md §2 rows 19-20)");
        return;
    }
    if (bvh->nodes.empty()) {
        if (writeImage) camera->film->WriteImage();
        return;
    }
    std::shared_ptr<DeviceScene> ds = GetDeviceScene(*bvh, scene.lights, lightSampleStrategy);
    if (!ds) return;
    Film *film = camera->film;
    pb2_camera cam = camera->Desc();
    pb2_film_desc fd = film->Desc();
    pb2_path_params pp = Params();
    std::vector<float> rgbw((size_t)4 * film->pixels.size());
    if (pb2_render_path(ds->handle, &cam, &fd, &pp, rgbw.data(), &lastStats) != PB2_OK) {
        Error("pb2_render_path failed: %s", pb2_last_error());
        return;
    }
    film->MergeDeviceFilm(rgbw.data());
    if (writeImage) film->WriteImage();
}

}  // namespace pbrt
