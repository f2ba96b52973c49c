// Typed name/value lists attached to every scene-file directive.
// Interface mirrors the reference's ParamSet (src/core/paramset.h:57-140: AddXxx / FindXxx /
// FindOneXxx / ReportUnused); the storage is a single tagged vector instead of one vector per type.
#ifndef PB2_HOST_PARAMSET_H
#define PB2_HOST_PARAMSET_H

#include <map>
#include <memory>
#include "core.h"

namespace pbrt {

struct Spectrum {  // RGBSpectrum (src/core/spectrum.h:430-470); the only build configuration in scope
    Float c[3];
    Spectrum(Float v = 0.f) { c[0] = c[1] = c[2] = v; }
    Spectrum(Float r, Float g, Float b) { c[0] = r; c[1] = g; c[2] = b; }
    Spectrum operator*(const Spectrum &s) const { return Spectrum(c[0] * s.c[0], c[1] * s.c[1], c[2] * s.c[2]); }
    bool IsBlack() const { return c[0] == 0 && c[1] == 0 && c[2] == 0; }
};

class ParamSet {
  public:
    enum class Type { Bool, Int, Float, Point2, Vector2, Point3, Vector3, Normal, Rgb, String, Texture };
    struct Item {
        Type type;
        std::string name;
        std::vector<double> nums;  // Bool/Int/Float/Point*/Vector*/Normal/Rgb, flattened
        std::vector<std::string> strs;
        mutable bool lookedUp = false;
    };

    void Add(Type t, const std::string &name, std::vector<double> nums, std::vector<std::string> strs = {}) {
        for (auto &it : items)
            if (it.type == t && it.name == name) {
                it.nums = std::move(nums);
                it.strs = std::move(strs);
                return;
            }
        items.push_back(Item{t, name, std::move(nums), std::move(strs), false});
    }
    void AddInt(const std::string &n, const std::vector<int> &v) { Add(Type::Int, n, std::vector<double>(v.begin(), v.end())); }
    void AddFloat(const std::string &n, const std::vector<Float> &v) { Add(Type::Float, n, std::vector<double>(v.begin(), v.end())); }
    void AddPoint3f(const std::string &n, const std::vector<Float> &v) { Add(Type::Point3, n, std::vector<double>(v.begin(), v.end())); }
    void AddNormal3f(const std::string &n, const std::vector<Float> &v) { Add(Type::Normal, n, std::vector<double>(v.begin(), v.end())); }
    void AddRGBSpectrum(const std::string &n, const std::vector<Float> &v) { Add(Type::Rgb, n, std::vector<double>(v.begin(), v.end())); }
    void AddString(const std::string &n, const std::string &v) { Add(Type::String, n, {}, {v}); }
    void AddBool(const std::string &n, bool v) { Add(Type::Bool, n, {v ? 1.0 : 0.0}); }

    const Item *Find(Type t, const std::string &name) const {
        for (const auto &it : items)
            if (it.type == t && it.name == name) {
                it.lookedUp = true;
                return &it;
            }
        return nullptr;
    }
    Float FindOneFloat(const std::string &n, Float d) const {
        const Item *it = Find(Type::Float, n);
        return (it && it->nums.size() == 1) ? (Float)it->nums[0] : d;
    }
    int FindOneInt(const std::string &n, int d) const {
        const Item *it = Find(Type::Int, n);
        return (it && it->nums.size() == 1) ? (int)it->nums[0] : d;
    }
    bool FindOneBool(const std::string &n, bool d) const {
        const Item *it = Find(Type::Bool, n);
        return (it && it->nums.size() == 1) ? it->nums[0] != 0 : d;
    }
    std::string FindOneString(const std::string &n, const std::string &d) const {
        const Item *it = Find(Type::String, n);
        return (it && it->strs.size() == 1) ? it->strs[0] : d;
    }
    Spectrum FindOneSpectrum(const std::string &n, const Spectrum &d) const {
        const Item *it = Find(Type::Rgb, n);
        if (it && it->nums.size() == 3) return Spectrum((Float)it->nums[0], (Float)it->nums[1], (Float)it->nums[2]);
        return d;
    }
    Point3f FindOnePoint3f(const std::string &n, const Point3f &d) const {
        const Item *it = Find(Type::Point3, n);
        if (it && it->nums.size() == 3) return Point3f((Float)it->nums[0], (Float)it->nums[1], (Float)it->nums[2]);
        return d;
    }
    std::string FindTexture(const std::string &n) const {
        const Item *it = Find(Type::Texture, n);
        return (it && it->strs.size() == 1) ? it->strs[0] : std::string();
    }
    // Array lookups return flattened floats / ints (count = number of scalars).
    std::vector<Float> FindFloats(Type t, const std::string &n, bool *found = nullptr) const {
        const Item *it = Find(t, n);
        if (found) *found = it != nullptr;
        std::vector<Float> r;
        if (it) r.assign(it->nums.begin(), it->nums.end());
        return r;
    }
    std::vector<int> FindInts(const std::string &n, bool *found = nullptr) const {
        const Item *it = Find(Type::Int, n);
        if (found) *found = it != nullptr;
        std::vector<int> r;
        if (it) {
            r.reserve(it->nums.size());
            for (double d : it->nums) r.push_back((int)d);
        }
        return r;
    }
    void ReportUnused() const {
        for (const auto &it : items)
            if (!it.lookedUp) Warning("Parameter \"%s\" not used", it.name.c_str());
    }
    void Clear() { items.clear(); }
    std::vector<Item> items;
};

// TextureParams (src/core/paramset.h:142-190) restricted to constant textures: a material
// parameter is looked up in the shape's parameters first, then in the material's.
// Named textures whose value does not vary over a surface: "constant", and "scale" / "mix" of such (src/textures/
// constant.h, scale.h, mix.h).  They are the Texture directives that stay inside this path; a material parameter that
// names one is simply that value.
// An "imagemap" texture (imagemap.h:72-128) with a "uv" mapping: what the MIPMap constructor receives plus the mapping
// and filter parameters - one pb2_texture of the scene description.
struct ImageTexture {
    int channels = 3;                 // 1: float texture, 3: spectrum texture
    int width = 0, height = 0;
    std::vector<float> texels;        // channels * width * height, row 0 is t = 0, after scale / gamma / luminance
    int wrap = 0;                     // PB2_WRAP_*
    bool trilinear = false;
    Float maxAniso = 8.f;
    Float su = 1, sv = 1, du = 0, dv = 0;
    // A node that is not an image (PB2_TEXKIND_*): a constant, or ScaleTexture / MixTexture over other nodes (scale.h, mix.h).
    // Combinators of constants alone are folded at the directive; these exist when an operand varies.
    int kind = 0;
    std::shared_ptr<ImageTexture> child[3];   // SCALE: tex1, tex2; MIX: tex1, tex2, amount
    Float value[3] = {0, 0, 0};
};
struct ConstantTextures {
    std::map<std::string, Float> floats;
    std::map<std::string, Spectrum> spectra;
    // the named textures that do vary: image maps
    std::map<std::string, std::shared_ptr<ImageTexture>> floatImages, spectrumImages;
};
class TextureParams {
  public:
    TextureParams(const ParamSet &geom, const ParamSet &mat, const ConstantTextures *tex = nullptr)
        : geomParams(geom), materialParams(mat), textures(tex) {}
    // the texture a parameter names (the shape's list first, paramset.cpp:649-655), or ""
    std::string NamedTexture(const std::string &n) const {
        std::string name = geomParams.FindTexture(n);
        return name == "" ? materialParams.FindTexture(n) : name;
    }
    // true when the parameter names a texture that is not one of the constant ones (image maps, procedurals, unknown names)
    bool IsVaryingTexture(const std::string &n) const {
        std::string name = NamedTexture(n);
        if (name == "") return false;
        return !(textures && (textures->floats.count(name) || textures->spectra.count(name)));
    }
    // the image texture a parameter names (float or spectrum ones as asked), or null
    std::shared_ptr<ImageTexture> GetImageTexture(const std::string &n, bool spectrum) const {
        std::string name = NamedTexture(n);
        if (name == "" || !textures) return nullptr;
        const auto &m = spectrum ? textures->spectrumImages : textures->floatImages;
        auto it = m.find(name);
        return it == m.end() ? nullptr : it->second;
    }
    // true when the parameter names a texture that is neither constant nor an image map of the right kind
    bool IsUnsupportedTexture(const std::string &n, bool spectrum) const {
        return IsVaryingTexture(n) && !GetImageTexture(n, spectrum);
    }
    Float FindFloat(const std::string &n, Float d) const { return geomParams.FindOneFloat(n, materialParams.FindOneFloat(n, d)); }
    Spectrum GetSpectrumTexture(const std::string &n, const Spectrum &def, bool *isTexture = nullptr) const {
        if (isTexture) *isTexture = false;
        std::string name = NamedTexture(n);
        if (name != "") {
            if (textures) {
                auto it = textures->spectra.find(name);
                if (it != textures->spectra.end()) return it->second;
            }
            if (isTexture) *isTexture = true;
            return def;
        }
        Spectrum s = materialParams.FindOneSpectrum(n, def);
        return geomParams.FindOneSpectrum(n, s);
    }
    Float GetFloatTexture(const std::string &n, Float def, bool *isTexture = nullptr) const {
        if (isTexture) *isTexture = false;
        std::string name = NamedTexture(n);
        if (name != "") {
            if (textures) {
                auto it = textures->floats.find(name);
                if (it != textures->floats.end()) return it->second;
            }
            if (isTexture) *isTexture = true;
            return def;
        }
        return geomParams.FindOneFloat(n, materialParams.FindOneFloat(n, def));
    }
    // GetFloatTextureOrNull (paramset.cpp:703-732) for constant values: false when neither parameter list has it
    bool GetFloatOrNull(const std::string &n, Float *value) const {
        std::string name = NamedTexture(n);
        if (name != "") {
            auto it = textures ? textures->floats.find(name) : std::map<std::string, Float>::const_iterator();
            if (textures && it != textures->floats.end()) {
                *value = it->second;
                return true;
            }
            return false;   // a varying texture: reported by the material's factory
        }
        for (const ParamSet *ps : {&geomParams, &materialParams}) {
            const ParamSet::Item *it = ps->Find(ParamSet::Type::Float, n);
            if (it && !it->nums.empty()) {
                *value = (Float)it->nums[0];
                return true;
            }
        }
        return false;
    }
    bool FindBool(const std::string &n, bool d) const { return geomParams.FindOneBool(n, materialParams.FindOneBool(n, d)); }
    std::string FindString(const std::string &n, const std::string &d = "") const {
        return geomParams.FindOneString(n, materialParams.FindOneString(n, d));
    }
    void ReportUnused() const { materialParams.ReportUnused(); }
    const ParamSet &geomParams, &materialParams;
    const ConstantTextures *textures;
};

}  // namespace pbrt
#endif
