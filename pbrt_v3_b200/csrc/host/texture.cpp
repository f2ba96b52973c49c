// Host side of image textures: the readers behind ReadImage and CreateImage{Float,Spectrum}Texture.
//
// Reference (behaviour only):
//   ReadImage / ReadImagePFM / ReadImageTGA / ReadImagePNG      src/core/imageio.cpp:60-79, 216-290, 350-430
//   ImageTexture::GetTexture (flip in y, convertIn)             src/textures/imagemap.cpp:50-107, imagemap.h:97-106
//   CreateImageFloatTexture / CreateImageSpectrumTexture        src/textures/imagemap.cpp:113-197
// The MIP pyramid and the filtering live in the CUDA library (pb2_texture, include/pb2.h); what is built here is the texel
// array the MIPMap constructor receives.  OpenEXR input needs the PIZ / ZIP wavelet codecs and is reported as an error.
#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

#include "paramset.h"
#include "scene.h"

namespace pbrt {

extern std::string g_sceneDirectory;  // api.cpp: directory of the file being parsed

static bool hasExtension(const std::string &name, const char *ext) {
    size_t n = strlen(ext);
    if (name.size() < n) return false;
    for (size_t i = 0; i < n; ++i)
        if (std::tolower((unsigned char)name[name.size() - n + i]) != ext[i]) return false;
    return true;
}

static bool readFile(const std::string &name, std::vector<uint8_t> *bytes) {
    std::ifstream f(name, std::ios::binary);
    if (!f) return false;
    bytes->assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    return true;
}

// ---- PFM: "PF" (three channels) or "Pf" (one), width height, scale (negative = little endian), rows bottom to top
static bool readPFM(const std::string &name, std::vector<float> *rgb, int *w, int *h) {
    std::vector<uint8_t> b;
    if (!readFile(name, &b)) return false;
    size_t pos = 0;
    auto word = [&](std::string *out) {
        out->clear();
        while (pos < b.size() && !std::isspace(b[pos])) out->push_back((char)b[pos++]);
        if (pos >= b.size()) return false;
        ++pos;   // exactly one white-space byte ends a header word
        return !out->empty();
    };
    std::string tok;
    if (!word(&tok)) return false;
    int nChannels;
    if (tok == "Pf") nChannels = 1;
    else if (tok == "PF") nChannels = 3;
    else return false;
    if (!word(&tok)) return false;
    *w = atoi(tok.c_str());
    if (!word(&tok)) return false;
    *h = atoi(tok.c_str());
    if (!word(&tok)) return false;
    float scale = (float)atof(tok.c_str());
    if (*w <= 0 || *h <= 0) return false;
    const size_t nFloats = (size_t)nChannels * *w * *h;
    if (b.size() - pos < nFloats * 4) return false;
    std::vector<float> data(nFloats);
    for (int y = *h - 1; y >= 0; --y) {   // the file's first row is the image's last
        memcpy(&data[(size_t)y * nChannels * *w], &b[pos], (size_t)nChannels * *w * 4);
        pos += (size_t)nChannels * *w * 4;
    }
    const bool fileLittleEndian = scale < 0.f;
    const uint16_t probe = 1;
    const bool hostLittleEndian = *reinterpret_cast<const uint8_t *>(&probe) == 1;
    if (hostLittleEndian != fileLittleEndian)
        for (size_t i = 0; i < nFloats; ++i) {
            uint8_t t[4];
            memcpy(t, &data[i], 4);
            std::swap(t[0], t[3]);
            std::swap(t[1], t[2]);
            memcpy(&data[i], t, 4);
        }
    if (std::abs(scale) != 1.f)
        for (size_t i = 0; i < nFloats; ++i) data[i] *= std::abs(scale);
    rgb->resize((size_t)3 * *w * *h);
    for (size_t i = 0; i < (size_t)*w * *h; ++i)
        for (int c = 0; c < 3; ++c) (*rgb)[3 * i + c] = nChannels == 1 ? data[i] : data[3 * i + c];
    return true;
}

// ---- TGA: true-colour (24 / 32 bit), grey (8 bit) and colour-mapped images, raw or run-length encoded
static bool readTGA(const std::string &name, std::vector<float> *rgb, int *w, int *h) {
    std::vector<uint8_t> b;
    if (!readFile(name, &b) || b.size() < 18) return false;
    const int idLen = b[0], cmapType = b[1], imgType = b[2];
    const int cmapLen = b[5] | (b[6] << 8), cmapBits = b[7];
    *w = b[12] | (b[13] << 8);
    *h = b[14] | (b[15] << 8);
    const int bpp = b[16], desc = b[17];
    const bool rle = imgType >= 9;
    const int kind = imgType & 7;   // 1 colour-mapped, 2 true colour, 3 grey
    if (*w <= 0 || *h <= 0 || (kind != 1 && kind != 2 && kind != 3)) return false;
    if ((kind == 2 && bpp != 24 && bpp != 32) || (kind == 3 && bpp != 8) || (kind == 1 && (bpp != 8 || (cmapBits != 24 && cmapBits != 32)))) return false;
    size_t pos = 18 + (size_t)idLen;
    const uint8_t *cmap = nullptr;
    const int cmapBytes = cmapBits / 8;
    if (cmapType) {
        cmap = b.data() + pos;
        pos += (size_t)cmapLen * cmapBytes;
    }
    const int bytesPP = bpp / 8;
    std::vector<uint8_t> px((size_t)*w * *h * bytesPP);
    if (!rle) {
        if (b.size() < pos + px.size()) return false;
        memcpy(px.data(), &b[pos], px.size());
    } else {
        size_t o = 0;
        while (o < px.size()) {
            if (pos >= b.size()) return false;
            const int head = b[pos++], count = (head & 127) + 1;
            if (head & 128) {
                if (pos + bytesPP > b.size()) return false;
                for (int k = 0; k < count && o < px.size(); ++k, o += bytesPP) memcpy(&px[o], &b[pos], bytesPP);
                pos += bytesPP;
            } else {
                const size_t nb = (size_t)count * bytesPP;
                if (pos + nb > b.size() || o + nb > px.size()) return false;
                memcpy(&px[o], &b[pos], nb);
                pos += nb;
                o += nb;
            }
        }
    }
    const bool rightToLeft = (desc & 16) != 0, topToBottom = (desc & 32) != 0;
    rgb->resize((size_t)3 * *w * *h);
    for (int y = 0; y < *h; ++y)
        for (int x = 0; x < *w; ++x) {
            const int sx = rightToLeft ? *w - 1 - x : x, sy = topToBottom ? y : *h - 1 - y;
            const uint8_t *src = &px[((size_t)sy * *w + sx) * bytesPP];
            float *dst = &(*rgb)[3 * ((size_t)y * *w + x)];
            if (kind == 3) dst[0] = dst[1] = dst[2] = *src / 255.f;
            else {
                if (kind == 1) {
                    if (!cmap || *src >= cmapLen) return false;
                    src = cmap + (size_t)*src * cmapBytes;
                }
                dst[2] = src[0] / 255.f;   // stored blue, green, red
                dst[1] = src[1] / 255.f;
                dst[0] = src[2] / 255.f;
            }
        }
    return true;
}

// ---- PNG: every colour type and bit depth, not interlaced; decoded to 8-bit RGB as lodepng_decode24 does (alpha dropped)
static bool readPNG(const std::string &name, std::vector<float> *rgb, int *w, int *h) {
    std::vector<uint8_t> b;
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (!readFile(name, &b) || b.size() < 8 || memcmp(b.data(), sig, 8) != 0) return false;
    auto be32 = [&](size_t p) { return ((uint32_t)b[p] << 24) | ((uint32_t)b[p + 1] << 16) | ((uint32_t)b[p + 2] << 8) | b[p + 3]; };
    size_t pos = 8;
    int depth = 0, colorType = 0, interlace = 0;
    std::vector<uint8_t> idat, palette;
    while (pos + 8 <= b.size()) {
        const uint32_t len = be32(pos);
        const std::string type((const char *)&b[pos + 4], 4);
        if (pos + 12 + (size_t)len > b.size()) return false;
        const uint8_t *data = &b[pos + 8];
        if (type == "IHDR") {
            if (len < 13) return false;
            *w = (int)be32(pos + 8);
            *h = (int)be32(pos + 12);
            depth = data[8];
            colorType = data[9];
            interlace = data[12];
        } else if (type == "PLTE") palette.assign(data, data + len);
        else if (type == "IDAT") idat.insert(idat.end(), data, data + len);
        else if (type == "IEND") break;
        pos += 12 + (size_t)len;
    }
    if (*w <= 0 || *h <= 0 || interlace != 0 || idat.empty()) return false;
    const int channels = colorType == 0 ? 1 : colorType == 2 ? 3 : colorType == 3 ? 1 : colorType == 4 ? 2 : colorType == 6 ? 4 : 0;
    if (!channels || (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16)) return false;
    const size_t bitsPP = (size_t)channels * depth, stride = ((size_t)*w * bitsPP + 7) / 8, bpp = std::max<size_t>(1, bitsPP / 8);
    std::vector<uint8_t> raw((stride + 1) * *h);
    uLongf rawLen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size()) return false;
    // undo the scanline filters
    std::vector<uint8_t> img(stride * *h);
    for (int y = 0; y < *h; ++y) {
        const uint8_t *in = &raw[(stride + 1) * y];
        uint8_t *out = &img[stride * y];
        const uint8_t *up = y ? out - stride : nullptr;
        const int ft = in[0];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? out[i - bpp] : 0, bb = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
            int pred = 0;
            if (ft == 1) pred = a;
            else if (ft == 2) pred = bb;
            else if (ft == 3) pred = (a + bb) / 2;
            else if (ft == 4) {
                const int p = a + bb - c, pa = std::abs(p - a), pb = std::abs(p - bb), pc = std::abs(p - c);
                pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? bb : c);
            } else if (ft != 0)
                return false;
            out[i] = (uint8_t)(in[1 + i] + pred);
        }
    }
    auto sample = [&](int y, int x, int c) -> int {   // channel c of pixel (x, y) as an 8-bit value
        const uint8_t *row = &img[stride * y];
        if (depth == 8) return row[(size_t)x * channels + c];
        if (depth == 16) return row[((size_t)x * channels + c) * 2];   // the high byte
        const size_t bit = (size_t)x * depth;
        const int v = (row[bit / 8] >> (8 - depth - (bit % 8))) & ((1 << depth) - 1);
        return colorType == 3 ? v : v * 255 / ((1 << depth) - 1);
    };
    rgb->resize((size_t)3 * *w * *h);
    for (int y = 0; y < *h; ++y)
        for (int x = 0; x < *w; ++x) {
            int r, g, bl;
            if (colorType == 3) {
                const int idx = sample(y, x, 0);
                if ((size_t)idx * 3 + 2 >= palette.size()) return false;
                r = palette[idx * 3];
                g = palette[idx * 3 + 1];
                bl = palette[idx * 3 + 2];
            } else if (channels <= 2)
                r = g = bl = sample(y, x, 0);
            else {
                r = sample(y, x, 0);
                g = sample(y, x, 1);
                bl = sample(y, x, 2);
            }
            float *dst = &(*rgb)[3 * ((size_t)y * *w + x)];
            dst[0] = r / 255.f;
            dst[1] = g / 255.f;
            dst[2] = bl / 255.f;
        }
    return true;
}

// ReadImage (imageio.cpp:60-79): RGB per pixel, row 0 at the top
bool ReadImage(const std::string &name, std::vector<float> *rgb, int *w, int *h) {
    bool ok = false;
    if (hasExtension(name, ".pfm")) ok = readPFM(name, rgb, w, h);
    else if (hasExtension(name, ".tga")) ok = readTGA(name, rgb, w, h);
    else if (hasExtension(name, ".png")) ok = readPNG(name, rgb, w, h);
    else if (hasExtension(name, ".exr")) {
        Error("Unable to load \"%s\": OpenEXR input is outside the GPU path's scope (PFM, PNG and TGA are read)", name.c_str());
        return false;
    } else {
        Error("Unable to load image stored in format \"%s\" for filename \"%s\".",
              strrchr(name.c_str(), '.') ? (strrchr(name.c_str(), '.') + 1) : "(unknown)", name.c_str());
        return false;
    }
    if (!ok) Error("Unable to read image file \"%s\"", name.c_str());
    return ok;
}

static Float inverseGammaCorrect(Float value) {   // pbrt.h:298-301
    if (value <= 0.04045f) return value * 1.f / 12.92f;
    return std::pow((value + 0.055f) * 1.f / 1.055f, (Float)2.4f);
}

// CreateImageFloatTexture / CreateImageSpectrumTexture: the mapping and filter parameters, then ImageTexture::GetTexture
std::shared_ptr<ImageTexture> CreateImageTexture(const TextureParams &tp, bool spectrum) {
    auto tex = std::make_shared<ImageTexture>();
    std::string type = tp.FindString("mapping", "uv");
    if (type == "uv") {
        tex->su = tp.FindFloat("uscale", 1.);
        tex->sv = tp.FindFloat("vscale", 1.);
        tex->du = tp.FindFloat("udelta", 0.);
        tex->dv = tp.FindFloat("vdelta", 0.);
    } else if (type == "spherical" || type == "cylindrical" || type == "planar") {
        Error("2D texture mapping \"%s\" is outside the GPU path's scope (\"uv\" only); using \"uv\"", type.c_str());
    } else
        Error("2D texture mapping \"%s\" unknown", type.c_str());
    tex->maxAniso = tp.FindFloat("maxanisotropy", 8.f);
    tex->trilinear = tp.FindBool("trilinear", false);
    std::string wrap = tp.FindString("wrap", "repeat");
    tex->wrap = PB2_WRAP_REPEAT;
    if (wrap == "black") tex->wrap = PB2_WRAP_BLACK;
    else if (wrap == "clamp") tex->wrap = PB2_WRAP_CLAMP;
    Float scale = tp.FindFloat("scale", 1.f);
    std::string filename = tp.FindString("filename", "");
    const bool gamma = tp.FindBool("gamma", hasExtension(filename, ".tga") || hasExtension(filename, ".png"));
    std::string path = filename;
    if (!path.empty() && path[0] != '/' && !g_sceneDirectory.empty()) path = g_sceneDirectory + "/" + path;
    std::vector<float> rgb;
    int w = 0, h = 0;
    if (filename.empty() || !ReadImage(path, &rgb, &w, &h)) {
        // imagemap.cpp:67-74
        Warning("Creating a constant grey texture to replace \"%s\".", filename.c_str());
        w = h = 1;
        rgb.assign(3, 0.5f);
    }
    // flip in y: texture space has (0, 0) at the lower left corner (imagemap.cpp:77-84)
    for (int y = 0; y < h / 2; ++y)
        for (int x = 0; x < w; ++x)
            for (int c = 0; c < 3; ++c) std::swap(rgb[3 * ((size_t)y * w + x) + c], rgb[3 * ((size_t)(h - 1 - y) * w + x) + c]);
    tex->channels = spectrum ? 3 : 1;
    tex->width = w;
    tex->height = h;
    tex->texels.resize((size_t)tex->channels * w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        if (spectrum) {
            for (int c = 0; c < 3; ++c) tex->texels[3 * i + c] = scale * (gamma ? inverseGammaCorrect(rgb[3 * i + c]) : rgb[3 * i + c]);
        } else {
            // convertIn to Float: the luminance (RGBSpectrum::y, spectrum.h:462-465)
            const Float y = 0.212671f * rgb[3 * i] + 0.715160f * rgb[3 * i + 1] + 0.072169f * rgb[3 * i + 2];
            tex->texels[i] = scale * (gamma ? inverseGammaCorrect(y) : y);
        }
    }
    return tex;
}

// A ConstantTexture<Float> as a 1 x 1 image (the alpha mask of a mesh given "float alpha" 0, triangle.cpp:725-726)
std::shared_ptr<ImageTexture> ConstantFloatImage(Float v) {
    auto tex = std::make_shared<ImageTexture>();
    tex->channels = 1;
    tex->width = tex->height = 1;
    tex->texels.assign(1, v);
    return tex;
}

}  // namespace pbrt
