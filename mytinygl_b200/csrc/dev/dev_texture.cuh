/*
 * dev_texture.cuh -- texture sampling of the sm_100a back end (src/textures.c:272-557), shared by the tile kernels
 * (per fragment) and the set-up kernel (per point vertex).
 */
#ifndef MTGL_DEV_TEXTURE_CUH
#define MTGL_DEV_TEXTURE_CUH

#include "dev_common.cuh"

namespace mtgl_dev_impl {

/* ---------------------------------------------------------------- texture sampling (textures.c)
 * The sampler is split in two: tex_taps() resolves the filter, wraps the coordinates and fetches the (up to 8)
 * texels once; tex_channel() then filters ONE 8-bit channel.  Every stage of the reference works per channel
 * (bilinear_filter textures.c:294-307, the trilinear blend 512-515, all truncating to 8 bits), so evaluating
 * alpha first and the colour channels only for fragments that survive the alpha test is bit-identical. */
struct LevelTaps {
    uint32_t t00, t10, t01, t11;
    float fx, fy;
    uint32_t mode;          /* 0 = single texel in t00, 1 = bilinear, 2 = opaque white (missing level 1), 3 = not sampled (weight 0) */
};

struct TexTaps {
    LevelTaps a, b;
    float cl;               /* trilinear weight min(lod, 1) */
    bool tri;
};

__device__ __forceinline__ int wrap_coord(int x, int n, bool repeat)
{   /* get_texel_wrapped / get_mip1_texel_wrapped, textures.c:272-291, 357-376 */
    if (repeat) {
        if ((n & (n - 1)) == 0) return x & (n - 1);                 /* power-of-two size: the euclidean modulo is a mask */
        if ((unsigned)(x + n) < (unsigned)(3 * n)) {            /* x in [-n, 2n): one conditional add == the euclidean modulo */
            if (x < 0) x += n; else if (x >= n) x -= n;
        } else x = ((x % n) + n) % n;
    } else x = min(max(x, 0), n - 1);      /* n >= 1 */
    return x;
}

__device__ __forceinline__ void level_taps(LevelTaps &L, const uint32_t *px, int w, int h, bool rep_s, bool rep_t,
                                           float u, float v, bool linear)
{   /* texture_sample_base / _mip1 / tail of texture_sample_lod, textures.c:379-451, 524-556 */
    float tx = u * (float)w - 0.5f;
    float ty = v * (float)h - 0.5f;
    if (linear) {
        int x0 = f2i_x86(floorf(tx)), y0 = f2i_x86(floorf(ty));
        L.fx = tx - (float)x0; L.fy = ty - (float)y0;
        int xa = wrap_coord(x0, w, rep_s), xb = wrap_coord(x0 + 1, w, rep_s);
        int ya = wrap_coord(y0, h, rep_t), yb = wrap_coord(y0 + 1, h, rep_t);
        L.t00 = __ldg(px + ya * w + xa); L.t10 = __ldg(px + ya * w + xb);
        L.t01 = __ldg(px + yb * w + xa); L.t11 = __ldg(px + yb * w + xb);
        L.mode = 1;
    } else {
        int x = f2i_x86(floorf(tx + 0.5f)), y = f2i_x86(floorf(ty + 0.5f));
        if (x < 0) x = 0;
        if (x >= w) x = w - 1;
        if (y < 0) y = 0;
        if (y >= h) y = h - 1;
        L.t00 = __ldg(px + y * w + x);
        L.mode = 0;
    }
}

__device__ __forceinline__ void mip1_taps(LevelTaps &L, const RasterCfg *c, float u, float v, uint32_t filter)
{   /* texture_sample_mip1, textures.c:413-451: a level that cannot exist samples as opaque white */
    if (!c->tex_l1) { L.mode = 2; return; }
    bool linear = (filter == G_LINEAR || filter == G_LINEAR_MIPMAP_NEAREST || filter == G_LINEAR_MIPMAP_LINEAR);
    level_taps(L, c->tex_l1, c->tex_w1, c->tex_h1, c->tex_wrap_s == G_REPEAT, c->tex_wrap_t == G_REPEAT, u, v, linear);
}

__device__ __forceinline__ void tex_taps(TexTaps &T, const RasterCfg *c, float u, float v, float lod)
{   /* texture_sample_lod, textures.c:457-557 */
    const bool rep_s = c->tex_wrap_s == G_REPEAT, rep_t = c->tex_wrap_t == G_REPEAT;
    if (rep_s) { u = u - (float)f2i_x86(u); if (u < 0) u += 1.0f; }
    else { if (u < 0.0f) u = 0.0f; if (u > 1.0f) u = 1.0f; }
    if (rep_t) { v = v - (float)f2i_x86(v); if (v < 0) v += 1.0f; }
    else { if (v < 0.0f) v = 0.0f; if (v > 1.0f) v = 1.0f; }

    T.tri = false;
    uint32_t filter = (lod > 0.0f) ? c->tex_min : c->tex_mag;
    if (filter == G_NEAREST_MIPMAP_NEAREST || filter == G_LINEAR_MIPMAP_NEAREST) {
        if (lod >= 0.5f) { mip1_taps(T.a, c, u, v, filter); return; }
        filter = (filter == G_NEAREST_MIPMAP_NEAREST) ? G_NEAREST : G_LINEAR;
    } else if (filter == G_NEAREST_MIPMAP_LINEAR || filter == G_LINEAR_MIPMAP_LINEAR) {
        if (lod > 0.0f) {
            T.cl = (lod > 1.0f) ? 1.0f : lod;
            T.tri = true;
            /* lod >= 1: the blend of textures.c:512-515 is c0 * (1 - 1) + c1 * 1 with c0 in [0, 1]: c0 * 0 is +0 and
             * +0 + c1 is c1, bit for bit -- level 0 does not have to be fetched or filtered at all */
            if (T.cl == 1.0f) T.a.mode = 3;
            else level_taps(T.a, c->tex_l0, c->tex_w, c->tex_h, rep_s, rep_t, u, v, filter != G_NEAREST_MIPMAP_LINEAR);
            mip1_taps(T.b, c, u, v, filter);
            return;
        }
        filter = (filter == G_NEAREST_MIPMAP_LINEAR) ? G_NEAREST : G_LINEAR;
    }
    level_taps(T.a, c->tex_l0, c->tex_w, c->tex_h, rep_s, rep_t, u, v, filter == G_LINEAR);
}

/* one channel of color_to_rgba32 (graphics.h:337-348).  __saturatef differs from the reference's two ternaries only for
 * NaN (0 instead of NaN), and NaN * 255 converts to 0 as well (x86 cvttss2si yields 0x80000000, low byte 0), so the
 * packed byte is the same for every input -- in one instruction instead of four. */
__device__ __forceinline__ uint32_t pack1(float x) { return __float2uint_rz(__saturatef(x) * 255.0f); }

__device__ __forceinline__ uint32_t level_channel(const LevelTaps &L, int sh, const float *un)
{   /* bilinear_filter, textures.c:294-307: lerp horizontally, then vertically, truncate to 8 bits */
    if (L.mode == 0) return (L.t00 >> sh) & 0xFFu;
    if (L.mode == 2) return 0xFFu;
    if (L.mode == 3) return 0u;         /* not sampled: its weight in the trilinear blend is exactly 0 (tex_taps), any finite value does */
    float c00 = un[(L.t00 >> sh) & 0xFFu], c10 = un[(L.t10 >> sh) & 0xFFu];
    float c01 = un[(L.t01 >> sh) & 0xFFu], c11 = un[(L.t11 >> sh) & 0xFFu];
    float sx = 1.0f - L.fx, sy = 1.0f - L.fy;
    float top = c00 * sx + c10 * L.fx;
    float bot = c01 * sx + c11 * L.fx;
    return pack1(top * sy + bot * L.fy);
}

__device__ __forceinline__ float tex_channel(const TexTaps &T, int sh, const float *un)
{   /* one channel of color_from_rgba32(texture_sample_lod(...)) */
    uint32_t v0 = level_channel(T.a, sh, un);
    if (!T.tri) return un[v0];
    uint32_t v1 = level_channel(T.b, sh, un);
    float s = 1.0f - T.cl;
    return un[pack1(un[v0] * s + un[v1] * T.cl)];      /* textures.c:512-515 */
}

} // namespace mtgl_dev_impl

#endif
