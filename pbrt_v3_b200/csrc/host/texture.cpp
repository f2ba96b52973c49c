// Host side of image textures: the readers behind ReadImage and CreateImage{Float,Spectrum}Texture.
//
// Reference (behaviour only):
//   ReadImage / ReadImagePFM / ReadImageTGA / ReadImagePNG      src/core/imageio.cpp:60-79, 216-290, 350-430
//   ImageTexture::GetTexture (flip in y, convertIn)             src/textures/imagemap.cpp:50-107, imagemap.h:97-106
//   CreateImageFloatTexture / CreateImageSpectrumTexture        src/textures/imagemap.cpp:113-197
// The MIP pyramid and the filtering live in the CUDA library (pb2_texture, include/pb2.h); what is built here is the texel
// array the MIPMap constructor receives.
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

// ---- OpenEXR: single-part scan-line files, half / float / uint channels R, G, B (or a lone Y), compression NONE, RLE, ZIPS,
// ZIP (what imgtool / pbrt's WriteImageEXR produce) and PIZ; the lossy codecs, tiled, deep and multi-part files are refused.
// ReadImageEXR (imageio.cpp:125-151) reads the data window through an RgbaInputFile: the same pixels.
static float halfToFloat(uint16_t h) {
    const uint32_t sign = (uint32_t)(h >> 15) << 31, exp = (h >> 10) & 31, man = h & 1023;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {   // subnormal half: normalise
            int e = -1;
            uint32_t m = man;
            do {
                ++e;
                m <<= 1;
            } while (!(m & 1024));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((m & 1023) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
    else bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}
// ---- the PIZ codec of OpenEXR (what most other tools write): a block's 16-bit words go through a value look-up table
// (only the values that occur are numbered), a 2D Haar-like wavelet per channel and a canonical Huffman code with a
// run-length symbol.  Decoder written from the format: range table, code lengths (6 bits each, zero runs escaped), codes
// assigned per length in symbol order, bits most significant first.
namespace {
struct PizBits {   // most-significant-bit-first reader
    const uint8_t *p, *end;
    uint64_t acc = 0;
    int n = 0;
    bool ok = true;
    uint64_t get(int bits) {
        while (n < bits) {
            if (p >= end) {
                ok = false;
                return 0;
            }
            acc = (acc << 8) | *p++;
            n += 8;
        }
        n -= bits;
        return (acc >> n) & ((bits == 64) ? ~0ull : ((1ull << bits) - 1));
    }
};
// one inverse wavelet step on a pair: 14-bit data (plain integer average / difference) or full 16-bit data (modulo form)
inline void pizUnpair(bool w14, uint16_t l, uint16_t hgh, uint16_t *a, uint16_t *b) {
    if (w14) {
        const int ls = (int16_t)l, hi = (int16_t)hgh;
        const int ai = ls + (hi & 1) + (hi >> 1);
        *a = (uint16_t)(int16_t)ai;
        *b = (uint16_t)(int16_t)(ai - hi);
    } else {
        const int m = l, d = hgh;
        const int bb = (m - (d >> 1)) & 0xffff;
        *b = (uint16_t)bb;
        *a = (uint16_t)((d + bb - 0x8000) & 0xffff);
    }
}
// inverse 2D wavelet of an nx x ny array with strides ox / oy, coarsest level first
void pizWaveletDecode(uint16_t *in, int nx, int ox, int ny, int oy, uint16_t maxValue) {
    const bool w14 = maxValue < (1 << 14);
    const int n = nx > ny ? ny : nx;
    int p = 1;
    while (p <= n) p <<= 1;
    p >>= 1;
    int p2 = p;
    p >>= 1;
    while (p >= 1) {
        const int oy1 = oy * p, oy2 = oy * p2, ox1 = ox * p, ox2 = ox * p2;
        int y = 0;
        for (; y <= ny - p2; y += p2) {
            uint16_t *row = in + (size_t)oy * y;
            int x = 0;
            for (; x <= nx - p2; x += p2) {
                uint16_t *px = row + (size_t)ox * x, *p01 = px + ox1, *p10 = px + oy1, *p11 = p10 + ox1;
                uint16_t i00, i01, i10, i11;
                pizUnpair(w14, *px, *p10, &i00, &i10);
                pizUnpair(w14, *p01, *p11, &i01, &i11);
                pizUnpair(w14, i00, i01, px, p01);
                pizUnpair(w14, i10, i11, p10, p11);
            }
            if (nx & p) {   // an odd column at this level: one vertical pair
                uint16_t *px = row + (size_t)ox * x, *p10 = px + oy1, i00;
                pizUnpair(w14, *px, *p10, &i00, p10);
                *px = i00;
            }
        }
        if (ny & p) {       // an odd row: horizontal pairs
            uint16_t *row = in + (size_t)oy * y;
            for (int x = 0; x <= nx - p2; x += p2) {
                uint16_t *px = row + (size_t)ox * x, *p01 = px + ox1, i00;
                pizUnpair(w14, *px, *p01, &i00, p01);
                *px = i00;
            }
        }
        (void)oy2;
        (void)ox2;
        p2 = p;
        p >>= 1;
    }
}
// Huffman: `data` = {first symbol, last symbol (= the run-length symbol), table length, bit count, reserved}, the packed
// code lengths, the bit stream; nOut 16-bit words come out
bool pizHuffmanDecode(const uint8_t *data, size_t size, uint16_t *out, size_t nOut) {
    if (size < 20) return nOut == 0;
    auto u32 = [&](size_t o) { uint32_t v; memcpy(&v, data + o, 4); return v; };
    const uint32_t im = u32(0), iM = u32(4), nBits = u32(12);
    const int kSymbols = (1 << 16) + 1;
    if (im >= (uint32_t)kSymbols || iM >= (uint32_t)kSymbols || im > iM) return false;
    std::vector<uint8_t> length((size_t)kSymbols, 0);
    PizBits tb{data + 20, data + size};
    for (uint32_t sym = im; sym <= iM; ++sym) {
        const int l = (int)tb.get(6);
        if (!tb.ok) return false;
        if (l == 63 || l >= 59) {   // a run of zero lengths: 63 -> 8 more bits + 6, 59..62 -> l - 59 + 2
            int run = l == 63 ? (int)tb.get(8) + 6 : l - 59 + 2;
            if (!tb.ok || sym + run > iM + 1) return false;
            sym += run - 1;
        } else
            length[sym] = (uint8_t)l;
    }
    // canonical codes: the numerically lowest code of each length, from the longest length down; then in symbol order
    uint64_t count[59] = {0}, first[59] = {0};
    for (int i = 0; i < kSymbols; ++i) count[length[i]]++;
    uint64_t c = 0;
    for (int l = 58; l > 0; --l) {
        first[l] = c;
        c = (c + count[l]) >> 1;
    }
    std::vector<uint32_t> offset(60, 0), sorted;
    sorted.reserve(iM - im + 1);
    for (int l = 1; l <= 58; ++l) {
        offset[l] = (uint32_t)sorted.size();
        if (count[l])
            for (uint32_t sym = im; sym <= iM; ++sym)
                if (length[sym] == l) sorted.push_back(sym);
    }
    // (the loop above is O(58 * symbols) only for lengths in use: a few dozen passes over at most 65537 entries)
    const uint8_t *bits = tb.n >= 8 ? tb.p - tb.n / 8 : tb.p;   // the bit stream starts at the next whole byte after the table
    PizBits in{bits, data + size};
    size_t o = 0;
    uint64_t used = 0;
    while (used < nBits) {
        uint64_t code = 0;
        int l = 0;
        uint32_t sym = 0;
        for (;;) {
            if (used >= nBits || l >= 58) return false;
            code = (code << 1) | in.get(1);
            if (!in.ok) return false;
            ++used;
            ++l;
            if (count[l] && code >= first[l] && code - first[l] < count[l]) {
                sym = sorted[offset[l] + (uint32_t)(code - first[l])];
                break;
            }
        }
        if (sym == iM) {   // run length: repeat the last word
            if (used + 8 > nBits || o == 0) return false;
            size_t run = (size_t)in.get(8);
            used += 8;
            if (!in.ok || o + run > nOut) return false;
            const uint16_t v = out[o - 1];
            while (run--) out[o++] = v;
        } else {
            if (o >= nOut) return false;
            out[o++] = (uint16_t)sym;
        }
    }
    return o == nOut;
}
}  // namespace

static bool readEXR(const std::string &name, std::vector<float> *rgb, int *w, int *h, std::string *why) {
    std::vector<uint8_t> b;
    if (!readFile(name, &b) || b.size() < 8) { *why = "cannot open"; return false; }
    auto i32 = [&](size_t p) { int32_t v; memcpy(&v, &b[p], 4); return v; };
    if ((uint32_t)i32(0) != 20000630u) { *why = "not an OpenEXR file"; return false; }
    const uint32_t version = (uint32_t)i32(4);
    if ((version & 0xff) != 2 || (version & (0x200 | 0x800 | 0x1000))) { *why = "tiled, deep or multi-part file"; return false; }
    size_t pos = 8;
    struct Channel { std::string name; int type, xs, ys; };
    std::vector<Channel> channels;
    int compression = -1, dw[4] = {0, 0, -1, -1};
    auto cstr = [&](std::string *out) {
        out->clear();
        while (pos < b.size() && b[pos]) out->push_back((char)b[pos++]);
        if (pos >= b.size()) return false;
        ++pos;
        return true;
    };
    for (;;) {
        std::string an, at;
        if (!cstr(&an)) { *why = "truncated header"; return false; }
        if (an.empty()) break;
        if (!cstr(&at) || pos + 4 > b.size()) { *why = "truncated header"; return false; }
        const int size = i32(pos);
        pos += 4;
        if (size < 0 || pos + (size_t)size > b.size()) { *why = "truncated header"; return false; }
        if (an == "channels") {
            size_t p = pos;
            while (p < pos + size && b[p]) {
                Channel c;
                while (b[p]) c.name.push_back((char)b[p++]);
                ++p;
                c.type = i32(p);
                c.xs = i32(p + 8);
                c.ys = i32(p + 12);
                p += 16;
                channels.push_back(c);
            }
        } else if (an == "compression") compression = b[pos];
        else if (an == "dataWindow")
            for (int k = 0; k < 4; ++k) dw[k] = i32(pos + 4 * k);
        pos += (size_t)size;
    }
    *w = dw[2] - dw[0] + 1;
    *h = dw[3] - dw[1] + 1;
    if (*w <= 0 || *h <= 0 || channels.empty()) { *why = "no data window / channels"; return false; }
    if (compression < 0 || compression > 4) { *why = "compression other than NONE / RLE / ZIPS / ZIP / PIZ (a lossy codec)"; return false; }
    size_t lineBytes = 0;
    std::vector<size_t> chOffset(channels.size());
    int src[3] = {-1, -1, -1};
    for (size_t c = 0; c < channels.size(); ++c) {
        if (channels[c].xs != 1 || channels[c].ys != 1) { *why = "sub-sampled channels"; return false; }
        if (channels[c].type < 0 || channels[c].type > 2) { *why = "unknown pixel type"; return false; }
        chOffset[c] = lineBytes;
        lineBytes += (size_t)*w * (channels[c].type == 1 ? 2 : 4);
        if (channels[c].name == "R") src[0] = (int)c;
        if (channels[c].name == "G") src[1] = (int)c;
        if (channels[c].name == "B") src[2] = (int)c;
    }
    if (src[0] < 0 && src[1] < 0 && src[2] < 0)
        for (size_t c = 0; c < channels.size(); ++c)
            if (channels[c].name == "Y") src[0] = src[1] = src[2] = (int)c;
    if (src[0] < 0 && src[1] < 0 && src[2] < 0) { *why = "no R, G, B or Y channel"; return false; }
    const int linesPerBlock = compression == 3 ? 16 : (compression == 4 ? 32 : 1);
    const int nBlocks = (*h + linesPerBlock - 1) / linesPerBlock;
    if (pos + (size_t)nBlocks * 8 > b.size()) { *why = "truncated offset table"; return false; }
    rgb->assign((size_t)3 * *w * *h, 0.f);
    std::vector<uint8_t> raw, tmp;
    for (int blk = 0; blk < nBlocks; ++blk) {
        uint64_t off;
        memcpy(&off, &b[pos + (size_t)blk * 8], 8);
        if (off + 8 > b.size()) { *why = "bad chunk offset"; return false; }
        const int y0 = i32(off) - dw[1], dataSize = i32(off + 4);
        const int lines = std::min(linesPerBlock, *h - y0);
        if (y0 < 0 || lines <= 0 || dataSize < 0 || off + 8 + (size_t)dataSize > b.size()) { *why = "bad chunk"; return false; }
        const size_t rawSize = lineBytes * lines;
        raw.resize(rawSize);
        const uint8_t *data = &b[off + 8];
        if ((size_t)dataSize >= rawSize || compression == 0) {   // stored as is (a block that did not shrink is not compressed)
            if ((size_t)dataSize < rawSize) { *why = "short block"; return false; }
            memcpy(raw.data(), data, rawSize);
        } else if (compression == 4) {
            // PIZ: [min, max non-zero byte of the value bitmap][that part of the bitmap][Huffman length][Huffman data]; the
            // decoded words are the block channel by channel (each channel: its rows), wavelet-transformed per channel
            if (dataSize < 4) { *why = "bad PIZ block"; return false; }
            uint16_t minNZ, maxNZ;
            memcpy(&minNZ, data, 2);
            memcpy(&maxNZ, data + 2, 2);
            size_t p = 4;
            std::vector<uint8_t> bitmap(8192, 0);
            if (maxNZ >= 8192) { *why = "bad PIZ block"; return false; }
            if (minNZ <= maxNZ) {
                const size_t nb = (size_t)maxNZ - minNZ + 1;
                if (p + nb > (size_t)dataSize) { *why = "bad PIZ block"; return false; }
                memcpy(bitmap.data() + minNZ, data + p, nb);
                p += nb;
            }
            std::vector<uint16_t> lut(65536, 0);
            int k = 0;
            for (int i = 0; i < 65536; ++i)
                if (i == 0 || (bitmap[i >> 3] & (1 << (i & 7)))) lut[k++] = (uint16_t)i;
            const uint16_t maxValue = (uint16_t)(k - 1);
            if (p + 4 > (size_t)dataSize) { *why = "bad PIZ block"; return false; }
            int32_t hufLen;
            memcpy(&hufLen, data + p, 4);
            p += 4;
            if (hufLen < 0 || p + (size_t)hufLen > (size_t)dataSize) { *why = "bad PIZ block"; return false; }
            const size_t nWords = rawSize / 2;
            std::vector<uint16_t> words(nWords);
            if (!pizHuffmanDecode(data + p, (size_t)hufLen, words.data(), nWords)) { *why = "bad PIZ block (Huffman)"; return false; }
            size_t start = 0;
            std::vector<size_t> chStart(channels.size());
            for (size_t c = 0; c < channels.size(); ++c) {
                const int size = channels[c].type == 1 ? 1 : 2;   // 16-bit words per sample
                chStart[c] = start;
                for (int j = 0; j < size; ++j) pizWaveletDecode(words.data() + start + j, *w, size, lines, *w * size, maxValue);
                start += (size_t)*w * lines * size;
            }
            for (size_t i = 0; i < nWords; ++i) words[i] = lut[words[i]];
            // back to scan lines: for every line, every channel's samples
            uint8_t *dstp = raw.data();
            for (int l = 0; l < lines; ++l)
                for (size_t c = 0; c < channels.size(); ++c) {
                    const size_t n = (size_t)*w * (channels[c].type == 1 ? 1 : 2);
                    memcpy(dstp, words.data() + chStart[c] + (size_t)l * n, n * 2);
                    dstp += n * 2;
                }
        } else {
            tmp.resize(rawSize);
            if (compression == 1) {   // run-length: count < 0 -> -count literal bytes, else count + 1 copies of the next byte
                size_t o = 0, p = 0;
                while (p < (size_t)dataSize && o < rawSize) {
                    const int count = (int8_t)data[p++];
                    if (count < 0) {
                        const size_t n = (size_t)(-count);
                        if (p + n > (size_t)dataSize || o + n > rawSize) { *why = "bad RLE block"; return false; }
                        memcpy(&tmp[o], &data[p], n);
                        p += n;
                        o += n;
                    } else {
                        const size_t n = (size_t)count + 1;
                        if (p >= (size_t)dataSize || o + n > rawSize) { *why = "bad RLE block"; return false; }
                        memset(&tmp[o], data[p++], n);
                        o += n;
                    }
                }
                if (o != rawSize) { *why = "bad RLE block"; return false; }
            } else {
                uLongf n = (uLongf)rawSize;
                if (uncompress(tmp.data(), &n, data, (uLong)dataSize) != Z_OK || n != rawSize) { *why = "bad ZIP block"; return false; }
            }
            // undo the byte-delta predictor, then the split into even and odd bytes
            for (size_t i = 1; i < rawSize; ++i) tmp[i] = (uint8_t)(tmp[i - 1] + tmp[i] - 128);
            const size_t half = (rawSize + 1) / 2;
            for (size_t i = 0; i < rawSize; ++i) raw[i] = (i & 1) ? tmp[half + i / 2] : tmp[i / 2];
        }
        for (int l = 0; l < lines; ++l)
            for (int c = 0; c < 3; ++c) {
                if (src[c] < 0) continue;
                const Channel &ch = channels[(size_t)src[c]];
                const uint8_t *line = raw.data() + (size_t)l * lineBytes + chOffset[(size_t)src[c]];
                float *dst = rgb->data() + 3 * (size_t)(y0 + l) * *w + c;
                for (int x = 0; x < *w; ++x) {
                    float v;
                    if (ch.type == 1) {
                        uint16_t hv;
                        memcpy(&hv, line + 2 * (size_t)x, 2);
                        v = halfToFloat(hv);
                    } else if (ch.type == 2) memcpy(&v, line + 4 * (size_t)x, 4);
                    else {
                        uint32_t u;
                        memcpy(&u, line + 4 * (size_t)x, 4);
                        v = (float)u;
                    }
                    dst[3 * (size_t)x] = v;
                }
            }
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
        std::string why;
        ok = readEXR(name, rgb, w, h, &why);
        if (!ok) {
            Error("Unable to read OpenEXR file \"%s\": %s (scan-line files with NONE / RLE / ZIPS / ZIP / PIZ compression are read)", name.c_str(), why.c_str());
            return false;
        }
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
