// Image textures on the device: the MIP pyramid of one ImageTexture and MIPMap::Lookup.
//
// Reference: src/core/mipmap.h (class MIPMap: Texel :205-225, Lookup(st, width) :227-245, triangle :247-258,
// Lookup(st, dst0, dst1) :260-287, EWA :289-350), src/textures/imagemap.h:83-89 (ImageTexture::Evaluate),
// src/core/texture.cpp:93-99 (UVMapping2D::Map).  The pyramid itself is built on the host by the library
// (pb2_cuda.cu, buildTexturePyramid) the way the MIPMap constructor does.
//
// Layout in HBM: one float pool per scene (DScene::texels).  Its first 128 floats are MIPMap::weightLut; every texture's
// levels follow, finest first, each level row-major (t * width + s) with `channels` floats per texel - the reference blocks
// its levels (BlockedArray) for the CPU cache, the values are the same.
#ifndef PB2_TEXTURE_CUH
#define PB2_TEXTURE_CUH

#include "pb2_math.cuh"

namespace pb2 {

enum { TEX_MAX_LEVELS = 16, TEX_LUT_SIZE = 128 };

struct DTexture {
    int channels, nLevels;
    int w, h;                 // level 0 (a power of two in each direction)
    int wrap, doTrilinear;
    float maxAniso;
    float su, sv, du, dv;
    int kind;                 // PB2_TEXKIND_*: an image (the fields above), a constant, or a combinator over other textures
    long long levelOfs[TEX_MAX_LEVELS];   // float index of each level in the pool
    int child[3];             // 1 + index into the texture array (SCALE: two factors; MIX: two operands and the amount)
    float value[3];           // CONSTANT
    int pad2[2];
};

// pbrt's Mod (pbrt.h:291-296): the remainder is never negative
PB2_HD int texMod(int a, int b) {
    int r = a - (a / b) * b;
    return r < 0 ? r + b : r;
}

// MIPMap::Texel
PB2_HD V3 texTexel(const DTexture &tx, const float *pool, int level, int s, int t) {
    const int lw = (tx.w >> level) > 1 ? (tx.w >> level) : 1, lh = (tx.h >> level) > 1 ? (tx.h >> level) : 1;
    if (tx.wrap == PB2_WRAP_REPEAT) {
        s = texMod(s, lw);
        t = texMod(t, lh);
    } else if (tx.wrap == PB2_WRAP_CLAMP) {
        s = s < 0 ? 0 : (s > lw - 1 ? lw - 1 : s);
        t = t < 0 ? 0 : (t > lh - 1 ? lh - 1 : t);
    } else if (s < 0 || s >= lw || t < 0 || t >= lh)
        return mk3(0, 0, 0);
    const float *p = pool + tx.levelOfs[level] + ((long long)t * lw + s) * tx.channels;
    if (tx.channels == 3) return mk3(p[0], p[1], p[2]);
    return mk3(p[0], p[0], p[0]);
}

// MIPMap::triangle: the bilinear filter at one level
PB2_HD V3 texTriangle(const DTexture &tx, const float *pool, int level, V2 st) {
    level = level < 0 ? 0 : (level > tx.nLevels - 1 ? tx.nLevels - 1 : level);
    const int lw = (tx.w >> level) > 1 ? (tx.w >> level) : 1, lh = (tx.h >> level) > 1 ? (tx.h >> level) : 1;
    float s = st.x * lw - 0.5f, t = st.y * lh - 0.5f;
    int s0 = (int)floorf(s), t0 = (int)floorf(t);
    float ds = s - s0, dt = t - t0;
    return ((1 - ds) * (1 - dt)) * texTexel(tx, pool, level, s0, t0) + ((1 - ds) * dt) * texTexel(tx, pool, level, s0, t0 + 1) +
           (ds * (1 - dt)) * texTexel(tx, pool, level, s0 + 1, t0) + (ds * dt) * texTexel(tx, pool, level, s0 + 1, t0 + 1);
}

PB2_HD float texLog2(float x) { return plogf(x) * 1.442695040888963387004650940071f; }   // pbrt.h:306-309
PB2_HD V3 texLerp(float t, V3 a, V3 b) { return (1 - t) * a + t * b; }

// MIPMap::Lookup(st, width): trilinear
PB2_HD V3 texLookupWidth(const DTexture &tx, const float *pool, V2 st, float width) {
    float level = tx.nLevels - 1 + texLog2(pmax(width, 1e-8f));
    if (level < 0) return texTriangle(tx, pool, 0, st);
    if (level >= tx.nLevels - 1) return texTexel(tx, pool, tx.nLevels - 1, 0, 0);
    int iLevel = (int)floorf(level);
    float delta = level - iLevel;
    return texLerp(delta, texTriangle(tx, pool, iLevel, st), texTriangle(tx, pool, iLevel + 1, st));
}

// MIPMap::EWA: the elliptically weighted average at one level
PB2_HDN V3 texEwa(const DTexture &tx, const float *pool, int level, V2 st, V2 dst0, V2 dst1) {
    if (level >= tx.nLevels) return texTexel(tx, pool, tx.nLevels - 1, 0, 0);
    const int lw = (tx.w >> level) > 1 ? (tx.w >> level) : 1, lh = (tx.h >> level) > 1 ? (tx.h >> level) : 1;
    st.x = st.x * lw - 0.5f;
    st.y = st.y * lh - 0.5f;
    dst0.x *= lw;
    dst0.y *= lh;
    dst1.x *= lw;
    dst1.y *= lh;
    float A = dst0.y * dst0.y + dst1.y * dst1.y + 1;
    float B = -2 * (dst0.x * dst0.y + dst1.x * dst1.y);
    float C = dst0.x * dst0.x + dst1.x * dst1.x + 1;
    float invF = 1 / (A * C - B * B * 0.25f);
    A *= invF;
    B *= invF;
    C *= invF;
    float det = -B * B + 4 * A * C;
    float invDet = 1 / det;
    float uSqrt = sqrtf(det * C), vSqrt = sqrtf(A * det);
    int s0 = (int)ceilf(st.x - 2 * invDet * uSqrt);
    int s1 = (int)floorf(st.x + 2 * invDet * uSqrt);
    int t0 = (int)ceilf(st.y - 2 * invDet * vSqrt);
    int t1 = (int)floorf(st.y + 2 * invDet * vSqrt);
    V3 sum = mk3(0, 0, 0);
    float sumWts = 0;
    for (int it = t0; it <= t1; ++it) {
        float tt = it - st.y;
        for (int is = s0; is <= s1; ++is) {
            float ss = is - st.x;
            float r2 = A * ss * ss + B * ss * tt + C * tt * tt;
            if (r2 < 1) {
                int index = (int)(r2 * TEX_LUT_SIZE);
                if (index > TEX_LUT_SIZE - 1) index = TEX_LUT_SIZE - 1;
                float weight = pool[index];
                sum = sum + texTexel(tx, pool, level, is, it) * weight;
                sumWts += weight;
            }
        }
    }
    return mk3(sum.x / sumWts, sum.y / sumWts, sum.z / sumWts);
}

// MIPMap::Lookup(st, dst0, dst1)
PB2_HD V3 texLookup(const DTexture &tx, const float *pool, V2 st, V2 dst0, V2 dst1) {
    if (tx.doTrilinear) {
        float width = pmax(pmax(fabsf(dst0.x), fabsf(dst0.y)), pmax(fabsf(dst1.x), fabsf(dst1.y)));
        return texLookupWidth(tx, pool, st, width);
    }
    if (dst0.x * dst0.x + dst0.y * dst0.y < dst1.x * dst1.x + dst1.y * dst1.y) {
        V2 t = dst0;
        dst0 = dst1;
        dst1 = t;
    }
    float majorLength = sqrtf(dst0.x * dst0.x + dst0.y * dst0.y);
    float minorLength = sqrtf(dst1.x * dst1.x + dst1.y * dst1.y);
    if (minorLength * tx.maxAniso < majorLength && minorLength > 0) {
        float scale = majorLength / (minorLength * tx.maxAniso);
        dst1.x *= scale;
        dst1.y *= scale;
        minorLength *= scale;
    }
    if (minorLength == 0) return texTriangle(tx, pool, 0, st);
    float lod = pmax(0.f, tx.nLevels - 1.f + texLog2(minorLength));
    int ilod = (int)floorf(lod);
    return texLerp(lod - ilod, texEwa(tx, pool, ilod, st, dst0, dst1), texEwa(tx, pool, ilod + 1, st, dst0, dst1));
}

// The (u, v) differentials of a shaded point (SurfaceInteraction::dudx ..., interaction.h:120-121); all zero for a point
// that was not reached by a camera ray.
struct DUvDiff { float dudx, dvdx, dudy, dvdy; };

// ImageTexture::Evaluate with a UVMapping2D
PB2_HD V3 texEvaluate(const DTexture &tx, const float *pool, V2 uv, const DUvDiff &d) {
    V2 dstdx = mk2(tx.su * d.dudx, tx.sv * d.dvdx), dstdy = mk2(tx.su * d.dudy, tx.sv * d.dvdy);
    V2 st = mk2(tx.su * uv.x + tx.du, tx.sv * uv.y + tx.dv);
    return texLookup(tx, pool, st, dstdx, dstdy);
}

// (out of line: the combinator recursion below would otherwise inline the whole filter once per node position)
PB2_HDN V3 texEvaluateImage(const DTexture &tx, const float *pool, V2 uv, const DUvDiff &d) { return texEvaluate(tx, pool, uv, d); }

// Texture::Evaluate for any node of the scene's texture array: an image, a constant, ScaleTexture (scale.h:56-58) or
// MixTexture (mix.h:57-61).  The library checks at upload that children precede their parents and that no chain is deeper
// than TEX_MAX_DEPTH, so the recursion unrolls at compile time.
enum { TEX_MAX_DEPTH = 3 };
// UVTexture::Evaluate (uv.h:53-59): the fractional parts of (s, t) as red and green
PB2_HD V3 texUv(const DTexture &tx, V2 uv) {
    const float s = tx.su * uv.x + tx.du, t = tx.sv * uv.y + tx.dv;
    return mk3(s - floorf(s), t - floorf(t), 0.f);
}
// Checkerboard2DTexture::Evaluate (checkerboard.h:63-101): how much of tex2 the point sees - 0 or 1 for a point sample, the
// box-filtered fraction in closed form when the footprint straddles a check
PB2_HD float texCheckerWeight(const DTexture &tx, V2 uv, const DUvDiff &d) {
    const float s = tx.su * uv.x + tx.du, t = tx.sv * uv.y + tx.dv;
    const float point = (((int)floorf(s) + (int)floorf(t)) % 2 == 0) ? 0.f : 1.f;
    if (tx.value[0] == 0) return point;   // "aamode" "none"
    const float ds = pmax(fabsf(tx.su * d.dudx), fabsf(tx.su * d.dudy)), dt = pmax(fabsf(tx.sv * d.dvdx), fabsf(tx.sv * d.dvdy));
    const float s0 = s - ds, s1 = s + ds, t0 = t - dt, t1 = t + dt;
    if (floorf(s0) == floorf(s1) && floorf(t0) == floorf(t1)) return point;
    auto bumpInt = [](float x) { return (int)floorf(x / 2) + 2 * pmax(x / 2 - (int)floorf(x / 2) - 0.5f, 0.f); };
    const float sint = (bumpInt(s1) - bumpInt(s0)) / (2 * ds);
    const float tint = (bumpInt(t1) - bumpInt(t0)) / (2 * dt);
    float area2 = sint + tint - 2 * sint * tint;
    if (ds > 1 || dt > 1) area2 = .5f;
    return area2;
}
template <int DEPTH>
struct TexEval {
    static PB2_HD V3 node(const DTexture *textures, const float *pool, int id, V2 uv, const DUvDiff &d) {
        const DTexture &tx = textures[id];
        if (tx.kind == PB2_TEXKIND_IMAGE) return texEvaluateImage(tx, pool, uv, d);
        if (tx.kind == PB2_TEXKIND_CONSTANT) return mk3(tx.value[0], tx.value[1], tx.value[2]);
        if (tx.kind == PB2_TEXKIND_UV) return texUv(tx, uv);
        const V3 a = TexEval<DEPTH - 1>::node(textures, pool, tx.child[0] - 1, uv, d);
        const V3 b = TexEval<DEPTH - 1>::node(textures, pool, tx.child[1] - 1, uv, d);
        if (tx.kind == PB2_TEXKIND_SCALE) return a * b;
        if (tx.kind == PB2_TEXKIND_CHECKERBOARD) {
            const float area2 = texCheckerWeight(tx, uv, d);
            if (area2 == 0) return a;          // (the reference evaluates only the texture of the check it is in)
            if (area2 == 1) return b;
            return (1 - area2) * a + area2 * b;
        }
        const float amt = TexEval<DEPTH - 1>::node(textures, pool, tx.child[2] - 1, uv, d).x;
        return (1 - amt) * a + amt * b;
    }
};
template <>
struct TexEval<0> {
    static PB2_HD V3 node(const DTexture *textures, const float *pool, int id, V2 uv, const DUvDiff &d) {
        const DTexture &tx = textures[id];
        if (tx.kind == PB2_TEXKIND_CONSTANT) return mk3(tx.value[0], tx.value[1], tx.value[2]);
        if (tx.kind == PB2_TEXKIND_UV) return texUv(tx, uv);
        return texEvaluateImage(tx, pool, uv, d);
    }
};
PB2_HD V3 texEvaluateNode(const DTexture *textures, const float *pool, int id, V2 uv, const DUvDiff &d) {
    return TexEval<TEX_MAX_DEPTH>::node(textures, pool, id, uv, d);
}

}  // namespace pb2
#endif
