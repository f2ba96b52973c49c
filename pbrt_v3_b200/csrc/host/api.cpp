// Scene-description state machine + tokenizer/parser for the .pbrt surface the path-tracing
// configurations use.  Behaviour follows the reference:
//   graphics state, CTM stack, pbrtXxx calls, pbrtWorldEnd     src/core/api.cpp:871-1649
//   name -> factory tables                                      src/core/api.cpp:426-868
//   tokenizer, parameter lists, directive dispatch              src/core/parser.cpp:98-320, 413-485, 712-1091
// Only static transforms (no ActiveTransform / animation), constant textures and the plugins in
// scope are accepted; anything else is reported through Error() and skipped, like an unknown
// plugin name in the reference.
#include "api.h"

#include <fstream>
#include <chrono>
#include <list>
#include <sstream>

namespace pbrt {

std::string g_sceneDirectory;
std::string g_imageFileOverride;
Float g_cropWindow[2][2] = {{0, 1}, {0, 1}};
bool g_quickRender = false;

namespace {

enum class APIState { Uninitialized, OptionsBlock, WorldBlock };
APIState apiState = APIState::Uninitialized;

struct MaterialInstance {
    std::string name;
    ParamSet params;
    std::shared_ptr<Material> material;
};

struct GraphicsState {
    std::shared_ptr<MaterialInstance> currentMaterial;
    std::shared_ptr<std::map<std::string, std::shared_ptr<MaterialInstance>>> namedMaterials =
        std::make_shared<std::map<std::string, std::shared_ptr<MaterialInstance>>>();
    ParamSet areaLightParams;
    std::string areaLight;
    bool reverseOrientation = false;
    // floatTextures / spectrumTextures (api.cpp:160-165), constant-valued ones only; copied with the graphics state
    ConstantTextures textures;
};

struct RenderOptions {
    std::string FilterName = "box";
    ParamSet FilterParams;
    std::string FilmName = "image";
    ParamSet FilmParams;
    std::string SamplerName = "halton";
    ParamSet SamplerParams;
    std::string AcceleratorName = "bvh";
    ParamSet AcceleratorParams;
    std::string IntegratorName = "path";
    ParamSet IntegratorParams;
    std::string CameraName = "perspective";
    ParamSet CameraParams;
    Transform CameraToWorld;
    std::vector<std::shared_ptr<Light>> lights;
    std::vector<std::shared_ptr<Primitive>> primitives;
    // object instancing (api.cpp:181-183)
    std::map<std::string, std::vector<std::shared_ptr<Primitive>>> instances;
    std::vector<std::shared_ptr<Primitive>> *currentInstance = nullptr;
};

Transform curTransform;
std::map<std::string, Transform> namedCoordinateSystems;
std::unique_ptr<RenderOptions> renderOptions;
GraphicsState graphicsState;
std::vector<GraphicsState> pushedGraphicsStates;
std::vector<Transform> pushedTransforms;
std::unique_ptr<RenderSetup> lastSetup;
bool renderAtWorldEnd = true;

// TransformCache (api.cpp:259-360): interned transforms outlive the shapes that point at them.
std::list<Transform> transformCache;
const Transform *internTransform(const Transform &t) {
    for (const Transform &c : transformCache)
        if (c == t) return &c;
    transformCache.push_back(t);
    return &transformCache.back();
}

bool verifyInitialized(const char *func) {
    if (apiState == APIState::Uninitialized) {
        Error("pbrtInit() must be before calling \"%s()\". Ignoring.", func);
        return false;
    }
    return true;
}
bool verifyOptions(const char *func) {
    if (!verifyInitialized(func)) return false;
    if (apiState == APIState::WorldBlock) {
        Error("Options cannot be set inside world block; \"%s\" not allowed.  Ignoring.", func);
        return false;
    }
    return true;
}
bool verifyWorld(const char *func) {
    if (!verifyInitialized(func)) return false;
    if (apiState == APIState::OptionsBlock) {
        Error("Scene description must be inside world block; \"%s\" not allowed. Ignoring.", func);
        return false;
    }
    return true;
}

std::shared_ptr<Material> MakeMaterial(const std::string &name, const TextureParams &mp) {
    Material *material = nullptr;
    if (name == "" || name == "none")
        return nullptr;
    else if (name == "matte")
        material = CreateMatteMaterial(mp);
    else if (name == "plastic")
        material = CreatePlasticMaterial(mp);
    else if (name == "substrate")
        material = CreateSubstrateMaterial(mp);
    else if (name == "uber")
        material = CreateUberMaterial(mp);
    else if (name == "metal")
        material = CreateMetalMaterial(mp);
    else if (name == "mirror")
        material = CreateMirrorMaterial(mp);
    else if (name == "glass")
        material = CreateGlassMaterial(mp);
    else {
        // api.cpp:587-590 falls back to matte for unknown names; materials that exist in the
        // reference but not here are an error so the difference is visible.
        Error("Material \"%s\" is outside the GPU path's scope (matte, plastic, substrate, metal, uber, mirror, glass). Using \"matte\".", name.c_str());
        material = CreateMatteMaterial(mp);
    }
    mp.ReportUnused();
    return std::shared_ptr<Material>(material);
}

bool shapeMaySetMaterialParameters(const ParamSet &ps) {
    for (const auto &it : ps.items) {
        if (it.type == ParamSet::Type::Texture) return true;
        if (it.type == ParamSet::Type::Rgb) return true;
        // api.cpp:214-238: any float parameter other than the mesh attributes may override a material value
        if (it.type == ParamSet::Type::Float && it.name != "alpha" && it.name != "shadowalpha" && it.name != "uv" &&
            it.name != "st" && it.name != "radius" && it.name != "zmin" && it.name != "zmax" && it.name != "phimax")
            return true;
        if (it.type == ParamSet::Type::String && it.name != "filename" && it.name != "type" && it.name != "scheme")
            return true;
        if (it.type == ParamSet::Type::Bool && it.name != "discarddegenerateUVs") return true;
    }
    return false;
}

std::shared_ptr<Material> materialForShape(const ParamSet &shapeParams) {
    if (!graphicsState.currentMaterial) {
        // default material: matte Kd 0.5 (api.cpp:207-210)
        static ParamSet empty;
        TextureParams mp(shapeParams, empty, &graphicsState.textures);
        return MakeMaterial("matte", mp);
    }
    if (shapeMaySetMaterialParameters(shapeParams)) {
        TextureParams mp(shapeParams, graphicsState.currentMaterial->params, &graphicsState.textures);
        return MakeMaterial(graphicsState.currentMaterial->name, mp);
    }
    return graphicsState.currentMaterial->material;
}

std::vector<std::shared_ptr<Shape>> MakeShapes(const std::string &name, const Transform *o2w, const Transform *w2o,
                                               bool reverseOrientation, const ParamSet &ps) {
    std::vector<std::shared_ptr<Shape>> shapes;
    if (name == "sphere") {
        shapes.push_back(CreateSphereShape(o2w, w2o, reverseOrientation, ps));
    } else if (name == "trianglemesh")
        shapes = CreateTriangleMeshShape(o2w, w2o, reverseOrientation, ps);
    else if (name == "plymesh")
        shapes = CreatePLYMesh(o2w, w2o, reverseOrientation, ps);
    else if (name == "loopsubdiv")
        shapes = CreateLoopSubdiv(o2w, w2o, reverseOrientation, ps);
    else
        Error("Shape \"%s\" is outside the GPU path's scope (sphere, trianglemesh, plymesh, loopsubdiv).", name.c_str());
    if ((name == "trianglemesh" || name == "plymesh") && !shapes.empty()) {
        // the "alpha" / "shadowalpha" masks of a mesh (triangle.cpp:717-740, plymesh.cpp:349-372): a float texture by name,
        // or the constant 0 (a mesh that is never hit)
        auto mask = [&](const char *pn) -> std::shared_ptr<ImageTexture> {
            const std::string texName = ps.FindTexture(pn);
            if (texName != "") {
                const ConstantTextures &tx = graphicsState.textures;
                auto it = tx.floatImages.find(texName);
                if (it != tx.floatImages.end()) return it->second;
                auto ic = tx.floats.find(texName);
                if (ic != tx.floats.end()) return ConstantFloatImage(ic->second);
                Error("Couldn't find float texture \"%s\" for \"%s\" parameter", texName.c_str(), pn);
                return nullptr;
            }
            if (ps.FindOneFloat(pn, 1.f) == 0.f) return ConstantFloatImage(0.f);
            return nullptr;
        };
        TriangleMesh *mesh = static_cast<Triangle *>(shapes[0].get())->mesh.get();
        mesh->alphaMask = mask("alpha");
        mesh->shadowAlphaMask = mask("shadowalpha");
    }
    return shapes;
}

}  // namespace

// ---------------------------------------------------------------- API
void pbrtInit(const Options &opt) {
    g_imageFileOverride = opt.imageFile;
    std::memcpy(g_cropWindow, opt.cropWindow, sizeof(g_cropWindow));
    g_quickRender = opt.quickRender;
    if (apiState != APIState::Uninitialized) Error("pbrtInit() has already been called.");
    apiState = APIState::OptionsBlock;
    renderOptions.reset(new RenderOptions);
    graphicsState = GraphicsState();
    curTransform = Transform();
}

bool pbrtIsInitialized() { return apiState != APIState::Uninitialized; }

void pbrtCleanup() {
    if (apiState == APIState::Uninitialized)
        Error("pbrtCleanup() called without pbrtInit().");
    else if (apiState == APIState::WorldBlock)
        Error("pbrtCleanup() called while inside world block.");
    apiState = APIState::Uninitialized;
    lastSetup.reset();
    renderOptions.reset();
    transformCache.clear();
}

void pbrtIdentity() { if (verifyInitialized("Identity")) curTransform = Transform(); }
void pbrtTranslate(Float dx, Float dy, Float dz) {
    if (verifyInitialized("Translate")) curTransform = curTransform * Translate(Vector3f(dx, dy, dz));
}
void pbrtRotate(Float angle, Float dx, Float dy, Float dz) {
    if (verifyInitialized("Rotate")) curTransform = curTransform * Rotate(angle, Vector3f(dx, dy, dz));
}
void pbrtScale(Float sx, Float sy, Float sz) {
    if (verifyInitialized("Scale")) curTransform = curTransform * Scale(sx, sy, sz);
}
void pbrtLookAt(Float ex, Float ey, Float ez, Float lx, Float ly, Float lz, Float ux, Float uy, Float uz) {
    if (!verifyInitialized("LookAt")) return;
    curTransform = curTransform * LookAt(Point3f(ex, ey, ez), Point3f(lx, ly, lz), Vector3f(ux, uy, uz));
}
static Matrix4x4 columnMajor(const Float tr[16]) {
    return Matrix4x4(tr[0], tr[4], tr[8], tr[12], tr[1], tr[5], tr[9], tr[13], tr[2], tr[6], tr[10], tr[14], tr[3],
                     tr[7], tr[11], tr[15]);
}
void pbrtTransform(Float tr[16]) { if (verifyInitialized("Transform")) curTransform = Transform(columnMajor(tr)); }
void pbrtConcatTransform(Float tr[16]) {
    if (verifyInitialized("ConcatTransform")) curTransform = curTransform * Transform(columnMajor(tr));
}
void pbrtCoordinateSystem(const std::string &name) {
    if (verifyInitialized("CoordinateSystem")) namedCoordinateSystems[name] = curTransform;
}
void pbrtCoordSysTransform(const std::string &name) {
    if (!verifyInitialized("CoordSysTransform")) return;
    auto it = namedCoordinateSystems.find(name);
    if (it != namedCoordinateSystems.end()) curTransform = it->second;
    else Warning("Couldn't find named coordinate system \"%s\"", name.c_str());
}
void pbrtPixelFilter(const std::string &name, const ParamSet &params) {
    if (!verifyOptions("PixelFilter")) return;
    renderOptions->FilterName = name;
    renderOptions->FilterParams = params;
}
void pbrtFilm(const std::string &type, const ParamSet &params) {
    if (!verifyOptions("Film")) return;
    renderOptions->FilmName = type;
    renderOptions->FilmParams = params;
}
void pbrtSampler(const std::string &name, const ParamSet &params) {
    if (!verifyOptions("Sampler")) return;
    renderOptions->SamplerName = name;
    renderOptions->SamplerParams = params;
}
void pbrtAccelerator(const std::string &name, const ParamSet &params) {
    if (!verifyOptions("Accelerator")) return;
    renderOptions->AcceleratorName = name;
    renderOptions->AcceleratorParams = params;
}
void pbrtIntegrator(const std::string &name, const ParamSet &params) {
    if (!verifyOptions("Integrator")) return;
    renderOptions->IntegratorName = name;
    renderOptions->IntegratorParams = params;
}
void pbrtCamera(const std::string &name, const ParamSet &params) {
    if (!verifyOptions("Camera")) return;
    renderOptions->CameraName = name;
    renderOptions->CameraParams = params;
    renderOptions->CameraToWorld = Inverse(curTransform);
    namedCoordinateSystems["camera"] = renderOptions->CameraToWorld;
}
void pbrtWorldBegin() {
    if (!verifyOptions("WorldBegin")) return;
    apiState = APIState::WorldBlock;
    curTransform = Transform();
    namedCoordinateSystems["world"] = curTransform;
}
void pbrtAttributeBegin() {
    if (!verifyWorld("AttributeBegin")) return;
    pushedGraphicsStates.push_back(graphicsState);
    pushedTransforms.push_back(curTransform);
}
void pbrtAttributeEnd() {
    if (!verifyWorld("AttributeEnd")) return;
    if (pushedGraphicsStates.empty()) {
        Error("Unmatched pbrtAttributeEnd() encountered. Ignoring it.");
        return;
    }
    graphicsState = pushedGraphicsStates.back();
    pushedGraphicsStates.pop_back();
    curTransform = pushedTransforms.back();
    pushedTransforms.pop_back();
}
void pbrtTransformBegin() {
    if (verifyWorld("TransformBegin")) pushedTransforms.push_back(curTransform);
}
void pbrtTransformEnd() {
    if (!verifyWorld("TransformEnd")) return;
    if (pushedTransforms.empty()) {
        Error("Unmatched pbrtTransformEnd() encountered. Ignoring it.");
        return;
    }
    curTransform = pushedTransforms.back();
    pushedTransforms.pop_back();
}
void pbrtMaterial(const std::string &name, const ParamSet &params) {
    if (!verifyWorld("Material")) return;
    static ParamSet empty;
    TextureParams mp(empty, params, &graphicsState.textures);
    auto mi = std::make_shared<MaterialInstance>();
    mi->name = name;
    mi->params = params;
    mi->material = MakeMaterial(name, mp);
    graphicsState.currentMaterial = mi;
}
void pbrtMakeNamedMaterial(const std::string &name, const ParamSet &params) {
    if (!verifyWorld("MakeNamedMaterial")) return;
    static ParamSet empty;
    TextureParams mp(empty, params, &graphicsState.textures);
    std::string matName = mp.FindString("type");
    if (matName == "") {
        Error("No parameter string \"type\" found in MakeNamedMaterial");
        return;
    }
    auto mi = std::make_shared<MaterialInstance>();
    mi->name = matName;
    mi->params = params;
    mi->material = MakeMaterial(matName, mp);
    if (graphicsState.namedMaterials->count(name)) Warning("Named material \"%s\" redefined.", name.c_str());
    // copy-on-write so that an enclosing AttributeBegin's map is not changed (api.cpp:1281-1286)
    graphicsState.namedMaterials =
        std::make_shared<std::map<std::string, std::shared_ptr<MaterialInstance>>>(*graphicsState.namedMaterials);
    (*graphicsState.namedMaterials)[name] = mi;
}
void pbrtNamedMaterial(const std::string &name) {
    if (!verifyWorld("NamedMaterial")) return;
    auto it = graphicsState.namedMaterials->find(name);
    if (it == graphicsState.namedMaterials->end()) {
        Error("NamedMaterial \"%s\" unknown.", name.c_str());
        return;
    }
    graphicsState.currentMaterial = it->second;
}
// pbrtTexture (api.cpp:1189-1245) for the texture classes whose value is the same everywhere: "constant" (constant.cpp),
// "scale" (scale.cpp:40-52: tex1 * tex2) and "mix" (mix.cpp:40-54: (1 - amount) * tex1 + amount * tex2) of constants.
void pbrtTexture(const std::string &name, const std::string &type, const std::string &texname, const ParamSet &params) {
    if (!verifyWorld("Texture")) return;
    TextureParams tp(params, params, &graphicsState.textures);
    const bool isFloat = type == "float", isSpectrum = type == "color" || type == "spectrum";
    if (!isFloat && !isSpectrum) {
        Error("Texture type \"%s\" unknown.", type.c_str());
        return;
    }
    auto varying = [&](std::initializer_list<const char *> names) {
        for (const char *n : names)
            if (tp.IsVaryingTexture(n)) return true;
        return false;
    };
    auto &floats = graphicsState.textures.floats;
    auto &spectra = graphicsState.textures.spectra;
    auto &floatImages = graphicsState.textures.floatImages;
    auto &spectrumImages = graphicsState.textures.spectrumImages;
    if (texname == "imagemap") {
        // CreateImageFloatTexture / CreateImageSpectrumTexture (imagemap.cpp:113-197); a later definition of the name
        // replaces an earlier one of either kind (api.cpp:1216-1243)
        std::shared_ptr<ImageTexture> tex = CreateImageTexture(tp, isSpectrum);
        params.ReportUnused();
        if ((isFloat ? floats.count(name) + floatImages.count(name) : spectra.count(name) + spectrumImages.count(name)) != 0)
            Warning("Texture \"%s\" being redefined", name.c_str());
        if (isFloat) {
            floats.erase(name);
            floatImages[name] = tex;
        } else {
            spectra.erase(name);
            spectrumImages[name] = tex;
        }
        return;
    }
    // "checkerboard" (2D, "uv" mapping; checkerboard.cpp:41-87) and "uv" (uv.cpp): always nodes
    if (texname == "checkerboard" || texname == "uv") {
        auto node = std::make_shared<ImageTexture>();
        node->channels = isSpectrum ? 3 : 1;
        bool ok = true;
        std::string mapping = tp.FindString("mapping", "uv");
        if (mapping != "uv") {
            Error("2D texture mapping \"%s\" is outside the GPU path's scope (\"uv\" only); using \"uv\"", mapping.c_str());
        }
        node->su = tp.FindFloat("uscale", 1.);
        node->sv = tp.FindFloat("vscale", 1.);
        node->du = tp.FindFloat("udelta", 0.);
        node->dv = tp.FindFloat("vdelta", 0.);
        if (texname == "uv") {
            node->kind = PB2_TEXKIND_UV;
            if (isFloat) {   // CreateUVFloatTexture returns nullptr (uv.cpp:40-43)
                Error("Texture \"%s\": \"uv\" is a spectrum texture", name.c_str());
                ok = false;
            }
        } else {
            const int dim = params.FindOneInt("dimension", 2);
            if (dim != 2) {
                Error("Texture \"%s\": the %d-dimensional checkerboard is outside the GPU path's scope (2D only)", name.c_str(), dim);
                ok = false;
            }
            auto operand = [&](const char *pn, Float def) -> std::shared_ptr<ImageTexture> {
                if (auto t = tp.GetImageTexture(pn, isSpectrum)) return t;
                auto c = std::make_shared<ImageTexture>();
                c->kind = PB2_TEXKIND_CONSTANT;
                c->channels = isSpectrum ? 3 : 1;
                if (isSpectrum) {
                    Spectrum v = tp.GetSpectrumTexture(pn, Spectrum(def));
                    for (int k = 0; k < 3; ++k) c->value[k] = v.c[k];
                } else
                    c->value[0] = tp.GetFloatTexture(pn, def);
                return c;
            };
            node->kind = PB2_TEXKIND_CHECKERBOARD;
            node->child[0] = operand("tex1", 1.f);
            node->child[1] = operand("tex2", 0.f);
            std::string aa = tp.FindString("aamode", "closedform");
            if (aa != "none" && aa != "closedform")
                Warning("Antialiasing mode \"%s\" not understood by Checkerboard2DTexture; using \"closedform\"", aa.c_str());
            node->value[0] = aa == "none" ? 0.f : 1.f;
        }
        params.ReportUnused();
        if (!ok) return;
        if (isFloat) {
            floats.erase(name);
            floatImages[name] = node;
        } else {
            spectra.erase(name);
            spectrumImages[name] = node;
        }
        return;
    }
    // "scale" / "mix" with an operand that varies (an image map, or a combinator of one): a node over the operands
    // (CreateScale*Texture scale.cpp:40-52, CreateMix*Texture mix.cpp:40-54); constants among them become constant nodes
    if ((texname == "scale" || texname == "mix") &&
        (tp.GetImageTexture("tex1", isSpectrum) || tp.GetImageTexture("tex2", isSpectrum) || (texname == "mix" && tp.GetImageTexture("amount", false)))) {
        auto operand = [&](const char *pn, bool spectrum, Float def) -> std::shared_ptr<ImageTexture> {
            if (auto t = tp.GetImageTexture(pn, spectrum)) return t;
            auto c = std::make_shared<ImageTexture>();
            c->kind = PB2_TEXKIND_CONSTANT;
            c->channels = spectrum ? 3 : 1;
            if (spectrum) {
                Spectrum v = tp.GetSpectrumTexture(pn, Spectrum(def));
                for (int k = 0; k < 3; ++k) c->value[k] = v.c[k];
            } else
                c->value[0] = tp.GetFloatTexture(pn, def);
            return c;
        };
        auto node = std::make_shared<ImageTexture>();
        node->channels = isSpectrum ? 3 : 1;
        if (texname == "scale") {
            node->kind = PB2_TEXKIND_SCALE;
            node->child[0] = operand("tex1", isSpectrum, 1.f);
            node->child[1] = operand("tex2", isSpectrum, 1.f);
        } else {
            node->kind = PB2_TEXKIND_MIX;
            node->child[0] = operand("tex1", isSpectrum, 0.f);
            node->child[1] = operand("tex2", isSpectrum, 1.f);
            node->child[2] = operand("amount", false, 0.5f);
        }
        params.ReportUnused();
        if (isFloat) {
            floats.erase(name);
            floatImages[name] = node;
        } else {
            spectra.erase(name);
            spectrumImages[name] = node;
        }
        return;
    }
    bool ok = false;
    Float fv = 0;
    Spectrum sv;
    if (texname == "constant") {
        ok = true;
        if (isFloat) fv = params.FindOneFloat("value", 1.f);     // constant.cpp:40-48: plain values, not texture names
        else sv = params.FindOneSpectrum("value", Spectrum(1.f));
    } else if (texname == "scale" && !varying({"tex1", "tex2"})) {
        ok = true;
        if (isFloat) fv = tp.GetFloatTexture("tex1", 1.f) * tp.GetFloatTexture("tex2", 1.f);
        else sv = tp.GetSpectrumTexture("tex1", Spectrum(1.f)) * tp.GetSpectrumTexture("tex2", Spectrum(1.f));
    } else if (texname == "mix" && !varying({"tex1", "tex2", "amount"})) {
        ok = true;
        Float amt = tp.GetFloatTexture("amount", 0.5f);
        if (isFloat) fv = (1 - amt) * tp.GetFloatTexture("tex1", 0.f) + amt * tp.GetFloatTexture("tex2", 1.f);
        else {
            Spectrum t1 = tp.GetSpectrumTexture("tex1", Spectrum(0.f)), t2 = tp.GetSpectrumTexture("tex2", Spectrum(1.f));
            for (int c = 0; c < 3; ++c) sv.c[c] = (1 - amt) * t1.c[c] + amt * t2.c[c];
        }
    }
    if (!ok) {
        Error("Texture \"%s\" of class \"%s\" is outside the GPU path's scope (constant textures and \"imagemap\", SURVEY.md §8 f.2); "
              "parameters that name it keep their defaults", name.c_str(), texname.c_str());
        return;
    }
    params.ReportUnused();
    if (isFloat) {
        if (floats.count(name) || floatImages.count(name)) Warning("Texture \"%s\" being redefined", name.c_str());
        floatImages.erase(name);
        floats[name] = fv;
    } else {
        if (spectra.count(name) || spectrumImages.count(name)) Warning("Texture \"%s\" being redefined", name.c_str());
        spectrumImages.erase(name);
        spectra[name] = sv;
    }
}
// api.cpp:1302-1316 with MakeLight (api.cpp:727-754) for the delta lights in scope
void pbrtLightSource(const std::string &name, const ParamSet &params) {
    if (!verifyWorld("LightSource")) return;
    std::shared_ptr<Light> lt;
    if (name == "point")
        lt = CreatePointLight(curTransform, params);
    else if (name == "spot")
        lt = CreateSpotLight(curTransform, params);
    else if (name == "distant")
        lt = CreateDistantLight(curTransform, params);
    else if (name == "infinite" || name == "exinfinite")
        lt = CreateInfiniteLight(curTransform, params);
    else if (name == "goniometric" || name == "projection")
        Error("LightSource \"%s\" is outside the GPU path's scope (point, spot, distant, infinite and diffuse area lights); skipped", name.c_str());
    else
        Error("LightSource: light type \"%s\" unknown.", name.c_str());
    params.ReportUnused();
    if (lt) renderOptions->lights.push_back(lt);
}
void pbrtAreaLightSource(const std::string &name, const ParamSet &params) {
    if (!verifyWorld("AreaLightSource")) return;
    graphicsState.areaLight = name;
    graphicsState.areaLightParams = params;
}
void pbrtReverseOrientation() {
    if (verifyWorld("ReverseOrientation")) graphicsState.reverseOrientation = !graphicsState.reverseOrientation;
}

void pbrtShape(const std::string &name, const ParamSet &params) {
    if (!verifyWorld("Shape")) return;
    const Transform *ObjToWorld = internTransform(curTransform);
    const Transform *WorldToObj = internTransform(Inverse(curTransform));
    std::vector<std::shared_ptr<Shape>> shapes = MakeShapes(name, ObjToWorld, WorldToObj, graphicsState.reverseOrientation, params);
    if (shapes.empty()) return;
    std::shared_ptr<Material> mtl = materialForShape(params);
    params.ReportUnused();
    std::vector<std::shared_ptr<Primitive>> prims;
    std::vector<std::shared_ptr<Light>> areaLights;
    prims.reserve(shapes.size());
    for (auto &s : shapes) {
        std::shared_ptr<AreaLight> area;
        if (graphicsState.areaLight != "") {
            if (graphicsState.areaLight == "area" || graphicsState.areaLight == "diffuse")
                area = CreateDiffuseAreaLight(curTransform, graphicsState.areaLightParams, s);
            else
                Warning("Area light \"%s\" unknown.", graphicsState.areaLight.c_str());
            if (area) areaLights.push_back(area);
        }
        prims.push_back(std::make_shared<GeometricPrimitive>(s, mtl, area));
    }
    graphicsState.areaLightParams.ReportUnused();
    // Add prims and areaLights to scene or current instance (api.cpp:1406-1419)
    if (renderOptions->currentInstance) {
        if (areaLights.size()) Warning("Area lights not supported with object instancing");
        renderOptions->currentInstance->insert(renderOptions->currentInstance->end(), prims.begin(), prims.end());
    } else {
        renderOptions->primitives.insert(renderOptions->primitives.end(), prims.begin(), prims.end());
        renderOptions->lights.insert(renderOptions->lights.end(), areaLights.begin(), areaLights.end());
    }
}

// api.cpp:1520-1543
void pbrtObjectBegin(const std::string &name) {
    if (!verifyWorld("ObjectBegin")) return;
    pbrtAttributeBegin();
    if (renderOptions->currentInstance) Error("ObjectBegin called inside of instance definition");
    renderOptions->instances[name] = std::vector<std::shared_ptr<Primitive>>();
    renderOptions->currentInstance = &renderOptions->instances[name];
}
void pbrtObjectEnd() {
    if (!verifyWorld("ObjectEnd")) return;
    if (!renderOptions->currentInstance) Error("ObjectEnd called outside of instance definition");
    renderOptions->currentInstance = nullptr;
    pbrtAttributeEnd();
}
// api.cpp:1547-1588.  The instance transform is static here (no ActiveTransform / TransformTimes on this path).
void pbrtObjectInstance(const std::string &name) {
    if (!verifyWorld("ObjectInstance")) return;
    if (renderOptions->currentInstance) {
        Error("ObjectInstance can't be called inside instance definition");
        return;
    }
    if (renderOptions->instances.find(name) == renderOptions->instances.end()) {
        Error("Unable to find instance named \"%s\"", name.c_str());
        return;
    }
    std::vector<std::shared_ptr<Primitive>> &in = renderOptions->instances[name];
    if (in.empty()) return;
    if (in.size() > 1) {
        // Create aggregate for instance Primitives
        std::shared_ptr<Primitive> accel;
        if (renderOptions->AcceleratorName == "bvh")
            accel = CreateBVHAccelerator(std::move(in), renderOptions->AcceleratorParams);
        else {
            Error("Accelerator \"%s\" is outside the GPU path's scope (bvh). Using \"bvh\".", renderOptions->AcceleratorName.c_str());
            accel = std::make_shared<BVHAccel>(std::move(in));
        }
        in.clear();
        in.push_back(accel);
    }
    renderOptions->primitives.push_back(std::make_shared<TransformedPrimitive>(in[0], *internTransform(curTransform)));
}

RenderSetup *pbrtLastSetup() { return lastSetup.get(); }
void pbrtSetRenderAtWorldEnd(bool render) { renderAtWorldEnd = render; }

void pbrtWorldEnd() {
    if (!verifyWorld("WorldEnd")) return;
    while (pushedGraphicsStates.size()) {
        Warning("Missing end to pbrtAttributeBegin()");
        pushedGraphicsStates.pop_back();
        pushedTransforms.pop_back();
    }
    while (pushedTransforms.size()) {
        Warning("Missing end to pbrtTransformBegin()");
        pushedTransforms.pop_back();
    }
    std::unique_ptr<RenderSetup> setup(new RenderSetup);
    RenderOptions &ro = *renderOptions;
    // MakeCamera (api.cpp:1716-1727)
    std::unique_ptr<Filter> filter;
    // MakeFilter (api.cpp:862-878)
    if (ro.FilterName == "box")
        filter.reset(CreateBoxFilter(ro.FilterParams));
    else if (ro.FilterName == "gaussian")
        filter.reset(CreateGaussianFilter(ro.FilterParams));
    else if (ro.FilterName == "mitchell")
        filter.reset(CreateMitchellFilter(ro.FilterParams));
    else if (ro.FilterName == "sinc")
        filter.reset(CreateSincFilter(ro.FilterParams));
    else if (ro.FilterName == "triangle")
        filter.reset(CreateTriangleFilter(ro.FilterParams));
    else {
        Error("Filter \"%s\" unknown. Using \"box\" with default radius.", ro.FilterName.c_str());
        filter.reset(new BoxFilter(0.5f, 0.5f));
    }
    ro.FilterParams.ReportUnused();
    if (ro.FilmName != "image") Error("Film \"%s\" unknown.", ro.FilmName.c_str());
    setup->film.reset(CreateFilm(ro.FilmParams, std::move(filter)));
    ro.FilmParams.ReportUnused();
    if (ro.CameraName != "perspective")
        Error("Camera \"%s\" is outside the GPU path's scope (perspective).", ro.CameraName.c_str());
    else
        setup->camera.reset(CreatePerspectiveCamera(ro.CameraParams, ro.CameraToWorld, setup->film.get()));
    ro.CameraParams.ReportUnused();
    // MakeIntegrator (api.cpp:1662-1714)
    if (setup->camera) {
        if (ro.SamplerName == "halton")
            setup->sampler.reset(CreateHaltonSampler(ro.SamplerParams, setup->film->GetSampleBounds()));
        else if (ro.SamplerName == "sobol")
            setup->sampler.reset(CreateSobolSampler(ro.SamplerParams, setup->film->GetSampleBounds()));
        else
            Error("Sampler \"%s\" is outside the GPU path's scope (halton, sobol).", ro.SamplerName.c_str());
        ro.SamplerParams.ReportUnused();
        if (setup->sampler) {
            if (ro.IntegratorName != "path")
                Error("Integrator \"%s\" is outside the GPU path's scope (path).", ro.IntegratorName.c_str());
            else
                setup->integrator.reset(CreatePathIntegrator(ro.IntegratorParams, setup->sampler, setup->camera));
            ro.IntegratorParams.ReportUnused();
        }
        if (ro.lights.empty()) Warning("No light sources defined in scene; rendering a black image.");
    }
    // MakeScene (api.cpp:1651-1660)
    std::shared_ptr<Primitive> accel;
    const auto tBuild0 = std::chrono::steady_clock::now();
    const size_t nScenePrims = ro.primitives.size();
    if (ro.AcceleratorName == "bvh")
        accel = CreateBVHAccelerator(std::move(ro.primitives), ro.AcceleratorParams);
    else {
        Error("Accelerator \"%s\" is outside the GPU path's scope (bvh). Using \"bvh\".", ro.AcceleratorName.c_str());
        accel = std::make_shared<BVHAccel>(std::move(ro.primitives));
    }
    ro.AcceleratorParams.ReportUnused();
    if (std::getenv("PB2_VERBOSE"))
        std::fprintf(stderr, "pb2: BVH over %zu primitives built in %.2f s\n", nScenePrims,
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - tBuild0).count());
    setup->scene.reset(new Scene(accel, ro.lights));
    ro.primitives.clear();
    ro.lights.clear();

    if (renderAtWorldEnd && setup->scene && setup->integrator) setup->integrator->Render(*setup->scene);

    lastSetup = std::move(setup);
    graphicsState = GraphicsState();
    apiState = APIState::OptionsBlock;
    renderOptions.reset(new RenderOptions);
    curTransform = Transform();
    namedCoordinateSystems.clear();
}

// ---------------------------------------------------------------- tokenizer + parser
namespace {

struct Tokenizer {
    std::string text, filename;
    size_t pos = 0;
    int line = 1;
    // Returns false at end of input.  Tokens: quoted strings (quotes kept), '[' , ']', bare words/numbers.
    bool next(std::string *tok) {
        while (pos < text.size()) {
            char c = text[pos];
            if (c == '\n') { ++line; ++pos; }
            else if (c == ' ' || c == '\t' || c == '\r') ++pos;
            else if (c == '#') { while (pos < text.size() && text[pos] != '\n') ++pos; }
            else break;
        }
        if (pos >= text.size()) return false;
        char c = text[pos];
        if (c == '"') {
            size_t start = pos++;
            std::string out = "\"";
            while (pos < text.size() && text[pos] != '"') {
                if (text[pos] == '\n') {
                    Error("%s:%d: Unterminated string", filename.c_str(), line);
                    ++line;
                }
                if (text[pos] == '\\' && pos + 1 < text.size()) {
                    ++pos;
                    switch (text[pos]) {
                    case 'b': out.push_back('\b'); break;
                    case 'f': out.push_back('\f'); break;
                    case 'n': out.push_back('\n'); break;
                    case 'r': out.push_back('\r'); break;
                    case 't': out.push_back('\t'); break;
                    case '\\': out.push_back('\\'); break;
                    case '\'': out.push_back('\''); break;
                    case '"': out.push_back('"'); break;
                    default: Error("%s:%d: unexpected escaped character \"%c\"", filename.c_str(), line, text[pos]);
                    }
                    ++pos;
                } else
                    out.push_back(text[pos++]);
            }
            if (pos >= text.size()) {
                Error("%s:%d: premature EOF", filename.c_str(), line);
                return false;
            }
            ++pos;
            out.push_back('"');
            (void)start;
            *tok = out;
            return true;
        }
        if (c == '[' || c == ']') {
            *tok = std::string(1, c);
            ++pos;
            return true;
        }
        size_t start = pos;
        while (pos < text.size()) {
            char d = text[pos];
            if (d == ' ' || d == '\n' || d == '\t' || d == '\r' || d == '"' || d == '[' || d == ']') break;
            ++pos;
        }
        *tok = text.substr(start, pos - start);
        return true;
    }
};

bool isQuoted(const std::string &s) { return s.size() >= 2 && s.front() == '"' && s.back() == '"'; }
std::string dequote(const std::string &s) { return s.substr(1, s.size() - 2); }

// parser.cpp:322-372: all-digit tokens go through strtol, everything else through strtof.
double parseNumber(const std::string &s, const Tokenizer &t) {
    if (s.size() == 1) {
        if (!(s[0] >= '0' && s[0] <= '9')) {
            Error("%s:%d: \"%c\": expected a number", t.filename.c_str(), t.line, s[0]);
            return 0;
        }
        return s[0] - '0';
    }
    bool isInt = true;
    for (char ch : s)
        if (!(ch >= '0' && ch <= '9')) isInt = false;
    char *end = nullptr;
    double val = isInt ? double(std::strtol(s.c_str(), &end, 10)) : (double)std::strtof(s.c_str(), &end);
    if (val == 0 && end == s.c_str()) Error("%s:%d: %s: expected a number", t.filename.c_str(), t.line, s.c_str());
    return val;
}

struct Parser {
    std::vector<std::unique_ptr<Tokenizer>> stack;
    bool haveUnget = false;
    std::string ungetTok;

    bool nextToken(std::string *tok) {
        if (haveUnget) {
            haveUnget = false;
            *tok = ungetTok;
            return true;
        }
        while (!stack.empty()) {
            if (stack.back()->next(tok)) return true;
            stack.pop_back();
            if (!stack.empty()) {
                size_t slash = stack.back()->filename.rfind('/');
                g_sceneDirectory = slash == std::string::npos ? "" : stack.back()->filename.substr(0, slash);
            }
        }
        return false;
    }
    void unget(const std::string &tok) {
        haveUnget = true;
        ungetTok = tok;
    }
    Tokenizer &cur() { return *stack.back(); }

    bool pushFile(const std::string &filename) {
        std::ifstream f(filename.c_str(), std::ios::binary);
        if (!f) {
            Error("Couldn't open scene file \"%s\"", filename.c_str());
            return false;
        }
        std::stringstream ss;
        ss << f.rdbuf();
        std::unique_ptr<Tokenizer> t(new Tokenizer);
        t->text = ss.str();
        t->filename = filename;
        stack.push_back(std::move(t));
        size_t slash = filename.rfind('/');
        g_sceneDirectory = slash == std::string::npos ? "" : filename.substr(0, slash);
        return true;
    }

    std::string requireString() {
        std::string tok;
        if (!nextToken(&tok) || !isQuoted(tok)) {
            Error("expected a quoted string, got \"%s\"", tok.c_str());
            return "";
        }
        return dequote(tok);
    }
    double requireNumber() {
        std::string tok;
        if (!nextToken(&tok)) {
            Error("premature EOF");
            return 0;
        }
        return parseNumber(tok, cur());
    }

    // parser.cpp:712-784 + AddParam 486-710
    ParamSet parseParams() {
        ParamSet ps;
        while (true) {
            std::string decl;
            if (!nextToken(&decl)) return ps;
            if (!isQuoted(decl)) {
                unget(decl);
                return ps;
            }
            decl = dequote(decl);
            std::istringstream ds(decl);
            std::string type, name;
            ds >> type >> name;
            if (type.empty() || name.empty()) {
                Error("Parameter \"%s\" doesn't have a type declaration?!", decl.c_str());
                return ps;
            }
            std::vector<double> nums;
            std::vector<std::string> strs;
            auto addVal = [&](const std::string &v) {
                if (isQuoted(v)) strs.push_back(dequote(v));
                else if (v == "true" || v == "false") strs.push_back(v);  // legacy bare bools
                else nums.push_back(parseNumber(v, cur()));
            };
            std::string val;
            if (!nextToken(&val)) {
                Error("premature EOF");
                return ps;
            }
            if (val == "[") {
                while (true) {
                    if (!nextToken(&val)) {
                        Error("premature EOF");
                        return ps;
                    }
                    if (val == "]") break;
                    addVal(val);
                }
            } else
                addVal(val);

            typedef ParamSet::Type T;
            auto needNums = [&](size_t multiple) {
                if (!strs.empty() || nums.empty() || nums.size() % multiple != 0) {
                    Error("Parameter \"%s\": unexpected number of values for type \"%s\"", name.c_str(), type.c_str());
                    return false;
                }
                return true;
            };
            if (type == "integer") { if (needNums(1)) ps.Add(T::Int, name, nums); }
            else if (type == "float") { if (needNums(1)) ps.Add(T::Float, name, nums); }
            else if (type == "point2") { if (needNums(2)) ps.Add(T::Point2, name, nums); }
            else if (type == "vector2") { if (needNums(2)) ps.Add(T::Vector2, name, nums); }
            else if (type == "point3" || type == "point") { if (needNums(3)) ps.Add(T::Point3, name, nums); }
            else if (type == "vector3" || type == "vector") { if (needNums(3)) ps.Add(T::Vector3, name, nums); }
            else if (type == "normal3" || type == "normal") { if (needNums(3)) ps.Add(T::Normal, name, nums); }
            else if (type == "rgb" || type == "color") { if (needNums(3)) ps.Add(T::Rgb, name, nums); }
            else if (type == "bool") {
                if (strs.size() == 1 && nums.empty()) ps.Add(T::Bool, name, {strs[0] == "true" ? 1.0 : 0.0});
                else Error("Parameter \"%s\": expected \"true\" or \"false\"", name.c_str());
            }
            else if (type == "string") { ps.Add(T::String, name, {}, strs); }
            else if (type == "texture") { ps.Add(T::Texture, name, {}, strs); }
            else if (type == "xyz" || type == "blackbody" || type == "spectrum")
                Error("Parameter \"%s\": spectrum type \"%s\" is outside the GPU path's scope (rgb/color only)", name.c_str(), type.c_str());
            else
                Error("Unable to decode type for name \"%s\"", decl.c_str());
        }
    }

    void run() {
        std::string tok;
        auto basic = [&](void (*fn)(const std::string &, const ParamSet &)) {
            std::string n = requireString();
            ParamSet ps = parseParams();
            fn(n, ps);
        };
        while (nextToken(&tok)) {
            if (tok == "AttributeBegin") pbrtAttributeBegin();
            else if (tok == "AttributeEnd") pbrtAttributeEnd();
            else if (tok == "TransformBegin") pbrtTransformBegin();
            else if (tok == "TransformEnd") pbrtTransformEnd();
            else if (tok == "WorldBegin") pbrtWorldBegin();
            else if (tok == "WorldEnd") pbrtWorldEnd();
            else if (tok == "Identity") pbrtIdentity();
            else if (tok == "ReverseOrientation") pbrtReverseOrientation();
            else if (tok == "Translate") { Float v[3]; for (Float &x : v) x = (Float)requireNumber(); pbrtTranslate(v[0], v[1], v[2]); }
            else if (tok == "Scale") { Float v[3]; for (Float &x : v) x = (Float)requireNumber(); pbrtScale(v[0], v[1], v[2]); }
            else if (tok == "Rotate") { Float v[4]; for (Float &x : v) x = (Float)requireNumber(); pbrtRotate(v[0], v[1], v[2], v[3]); }
            else if (tok == "LookAt") { Float v[9]; for (Float &x : v) x = (Float)requireNumber(); pbrtLookAt(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]); }
            else if (tok == "Transform" || tok == "ConcatTransform") {
                std::string b;
                nextToken(&b);
                if (b != "[") Error("%s: expected \"[\" after %s", cur().filename.c_str(), tok.c_str());
                Float m[16];
                for (Float &x : m) x = (Float)requireNumber();
                nextToken(&b);
                if (b != "]") Error("%s: expected \"]\" after %s values", cur().filename.c_str(), tok.c_str());
                if (tok == "Transform") pbrtTransform(m); else pbrtConcatTransform(m);
            }
            else if (tok == "CoordinateSystem") pbrtCoordinateSystem(requireString());
            else if (tok == "CoordSysTransform") pbrtCoordSysTransform(requireString());
            else if (tok == "Camera") basic(pbrtCamera);
            else if (tok == "Film") basic(pbrtFilm);
            else if (tok == "Sampler") basic(pbrtSampler);
            else if (tok == "Integrator") basic(pbrtIntegrator);
            else if (tok == "Accelerator") basic(pbrtAccelerator);
            else if (tok == "PixelFilter") basic(pbrtPixelFilter);
            else if (tok == "Material") basic(pbrtMaterial);
            else if (tok == "MakeNamedMaterial") basic(pbrtMakeNamedMaterial);
            else if (tok == "NamedMaterial") pbrtNamedMaterial(requireString());
            else if (tok == "AreaLightSource") basic(pbrtAreaLightSource);
            else if (tok == "Shape") basic(pbrtShape);
            else if (tok == "Include") {
                std::string fn = requireString();
                if (!fn.empty() && fn[0] != '/' && !g_sceneDirectory.empty()) fn = g_sceneDirectory + "/" + fn;
                pushFile(fn);
            }
            else if (tok == "ObjectBegin") pbrtObjectBegin(requireString());
            else if (tok == "ObjectEnd") pbrtObjectEnd();
            else if (tok == "ObjectInstance") pbrtObjectInstance(requireString());
            else if (tok == "LightSource") basic(pbrtLightSource);
            else if (tok == "Texture") {
                std::string name = requireString(), type = requireString(), texname = requireString();
                pbrtTexture(name, type, texname, parseParams());
            }
            else if (tok == "MakeNamedMedium" || tok == "MediumInterface") {
                // directives of the reference that lead outside this path (SURVEY.md §2 rows 15,33,42; §8 a19)
                Error("%s:%d: directive \"%s\" is outside the GPU path's scope; skipped", cur().filename.c_str(), cur().line, tok.c_str());
                std::string n;
                if (tok != "MediumInterface") {
                    requireString();
                    parseParams();
                } else {
                    requireString();
                    if (tok == "MediumInterface") {
                        if (nextToken(&n) && !isQuoted(n)) unget(n);
                    }
                }
            }
            else if (tok == "ActiveTransform" || tok == "TransformTimes") {
                Error("%s:%d: directive \"%s\" is outside the GPU path's scope; skipped", cur().filename.c_str(), cur().line, tok.c_str());
                if (tok == "ActiveTransform") nextToken(&tok);
                if (tok == "TransformTimes") { requireNumber(); requireNumber(); }
            }
            else
                Error("%s:%d: Unknown directive: %s", stack.empty() ? "" : cur().filename.c_str(), stack.empty() ? 0 : cur().line, tok.c_str());
        }
    }
};

}  // namespace

void pbrtParseFile(std::string filename) {
    Parser p;
    if (!p.pushFile(filename)) return;
    p.run();
}

void pbrtParseString(std::string str) {
    Parser p;
    std::unique_ptr<Tokenizer> t(new Tokenizer);
    t->text = std::move(str);
    t->filename = "<string>";
    p.stack.push_back(std::move(t));
    p.run();
}

}  // namespace pbrt
