/*
 * dev_fasttex.cuh -- texture_sample_lod (src/textures.c:457-557) from a texture STAGED IN SHARED MEMORY, shared by the
 * tile kernels that shade many fragments against one small texture (k_fill.cu, k_shade.cu).
 *
 * The tile's CTA converts the texture once into float4 texels n / 255 (level 0, then level 1): a bilinear tap is then one
 * 16-byte LDS with no per-channel conversion, where the general sampler (dev_texture.cuh) does a global load plus four
 * shift / mask / table look-ups per tap.  The filter decision, which depends only on the state and the PER-TRIANGLE lod
 * (raster.c:505-529), is made once per triangle (sampler_plan) instead of once per fragment; the 8-bit truncation
 * after every filter stage uses the arithmetic forms of dev_fragment.cuh.  Values are the reference's, bit for bit.
 *
 * Preconditions of the fast path (checked once per triangle by fast_texture_ok, everything else takes the general
 * sampler): the triangle's texture is the staged one; REPEAT only with power-of-two sizes (a mask); u, v, 1/w finite and
 * bounded, so that no NaN can reach the wrap or the floor.
 */
#ifndef MTGL_DEV_FASTTEX_CUH
#define MTGL_DEV_FASTTEX_CUH

#include "dev_common.cuh"
#include "dev_fragment.cuh"

namespace mtgl_dev_impl {

constexpr int STAGED_TEXELS = 64 * 64 + 32 * 32;        /* float4 texels: a 64x64 texture with its level 1 = 80 KB */

/* sampler plan: what texture_sample_lod decides from the state and the per-triangle LOD */
constexpr uint32_t TP_A_KIND = 3u;              /* level sampled first: 0 level 0, 1 level 1, 2 opaque white (missing level 1), 3 none (weight 0) */
constexpr uint32_t TP_A_LINEAR = 1u << 2;
constexpr uint32_t TP_TRI = 1u << 3;            /* blend with a second level, weight cl */
constexpr uint32_t TP_B_SHIFT = 4;              /* second level: kind in bits 4-5 */
constexpr uint32_t TP_B_LINEAR = 1u << 6;
constexpr uint32_t TP_REP_S = 1u << 7, TP_REP_T = 1u << 8;

struct StagedTex {              /* in shared memory, next to the texels */
    const uint32_t *id;         /* level-0 pointer of the staged texture = its identity; NULL: nothing staged */
    int w, h, w1, h1, n0;       /* level 1 starts at texel n0 */
};

__device__ __forceinline__ uint32_t sampler_plan(const RasterCfg *c, float lod, float &cl)
{
    uint32_t plan = (c->tex_wrap_s == G_REPEAT ? TP_REP_S : 0u) | (c->tex_wrap_t == G_REPEAT ? TP_REP_T : 0u);
    cl = 0.0f;
    uint32_t filter = (lod > 0.0f) ? c->tex_min : c->tex_mag;
    const uint32_t l1_kind = c->tex_l1 ? 1u : 2u;           /* textures.c:413-419: a level that cannot exist samples as white */
    if (filter == G_NEAREST_MIPMAP_NEAREST || filter == G_LINEAR_MIPMAP_NEAREST) {
        if (lod >= 0.5f) return plan | l1_kind | (filter == G_LINEAR_MIPMAP_NEAREST ? TP_A_LINEAR : 0u);
        filter = (filter == G_NEAREST_MIPMAP_NEAREST) ? G_NEAREST : G_LINEAR;
    } else if (filter == G_NEAREST_MIPMAP_LINEAR || filter == G_LINEAR_MIPMAP_LINEAR) {
        if (lod > 0.0f) {
            cl = (lod > 1.0f) ? 1.0f : lod;
            plan |= TP_TRI | (l1_kind << TP_B_SHIFT) | (filter == G_LINEAR_MIPMAP_LINEAR ? TP_B_LINEAR : 0u);
            if (cl == 1.0f) return plan | 3u;               /* weight of level 0 is exactly 0 (dev_texture.cuh, tex_taps) */
            return plan | 0u | (filter != G_NEAREST_MIPMAP_LINEAR ? TP_A_LINEAR : 0u);
        }
        filter = (filter == G_NEAREST_MIPMAP_LINEAR) ? G_NEAREST : G_LINEAR;
    }
    return plan | 0u | (filter == G_LINEAR ? TP_A_LINEAR : 0u);
}

__device__ __forceinline__ bool bounded_abs(float v, float lim) { return fabsf(v) <= lim; }   /* false for NaN */

/* may a triangle with these per-vertex texture coordinates and 1/w values take the fast path? */
__device__ __forceinline__ bool fast_texture_ok(const StagedTex &st, const RasterCfg *cfg, const float (&u)[3], const float (&v)[3], const float (&w)[3])
{
    if (!(cfg->flags & RC_TEXTURED) || st.id == nullptr || cfg->tex_l0 != st.id) return false;
    if (cfg->tex_wrap_s == G_REPEAT && (st.w & (st.w - 1))) return false;
    if (cfg->tex_wrap_t == G_REPEAT && (st.h & (st.h - 1))) return false;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; k++)
        ok = ok && bounded_abs(u[k], 1048576.0f) && bounded_abs(v[k], 1048576.0f) && w[k] >= 9.094947e-13f && w[k] <= 1.0995116e12f;   /* 2^-40 .. 2^40 */
    return ok;
}

/* all threads of the CTA: texture of state `tc` -> float4 texels (the caller synchronises before use) */
__device__ __forceinline__ void stage_texture(float4 *tex, StagedTex &st, const RasterCfg *tc, const float *un, int nthreads)
{
    const int n0 = tc->tex_w * tc->tex_h, n1 = tc->tex_l1 ? tc->tex_w1 * tc->tex_h1 : 0;
    if (!tc->tex_l0 || n0 + n1 > STAGED_TEXELS) return;
    for (int i = threadIdx.x; i < n0 + n1; i += nthreads) {
        const uint32_t t = (i < n0) ? __ldg(tc->tex_l0 + i) : __ldg(tc->tex_l1 + (i - n0));
        tex[i] = make_float4(un[t & 0xFFu], un[(t >> 8) & 0xFFu], un[(t >> 16) & 0xFFu], un[t >> 24]);
    }
    if (threadIdx.x == 0) { st.id = tc->tex_l0; st.w = tc->tex_w; st.h = tc->tex_h; st.w1 = tc->tex_w1; st.h1 = tc->tex_h1; st.n0 = n0; }
}

/* wrap of texture_sample_lod (textures.c:463-486) for finite coordinates */
__device__ __forceinline__ float fast_wrap(float u, bool repeat)
{
    if (repeat) { u = u - truncf(u); if (u < 0) u += 1.0f; return u; }
    return __saturatef(u);
}

struct FTaps { float4 t00, t10, t01, t11; float fx, fy; };

/* bilinear taps of one level (texture_sample_base / _mip1, textures.c:379-451).  floor(t) for |t| < 2^22 by a
 * round-down add of 1.5 * 2^23: the sum is 2^23 + 2^22 + floor(t) exactly, its low mantissa bits are the integer.
 * Texel coordinates: REPEAT with a power-of-two size is a mask, CLAMP a min / max -- one branch-free form for both. */
__device__ __forceinline__ void fast_taps(FTaps &T, const float4 *px, int w, int h, bool rep_s, bool rep_t, float u, float v)
{
    const float tx = u * (float)w - 0.5f, ty = v * (float)h - 0.5f;
    const float M = 12582912.0f;
    const float mx = __fadd_rd(tx, M), my = __fadd_rd(ty, M);
    const int x0 = __float_as_int(mx) - 0x4B400000, y0 = __float_as_int(my) - 0x4B400000;
    T.fx = tx - (mx - M); T.fy = ty - (my - M);
    const int wm = w - 1, hm = h - 1;
    const int ax = rep_s ? wm : -1, ay = rep_t ? hm : -1;
    const int xa = min(max(x0 & ax, 0), wm), xb = min(max((x0 + 1) & ax, 0), wm);
    const int ya = min(max(y0 & ay, 0), hm) * w, yb = min(max((y0 + 1) & ay, 0), hm) * w;
    T.t00 = px[ya + xa]; T.t10 = px[ya + xb]; T.t01 = px[yb + xa]; T.t11 = px[yb + xb];
}

/* bilinear_filter (textures.c:294-307) of one channel, truncated to 8 bits, as the float n / 255 */
__device__ __forceinline__ float fast_channel(float c00, float c10, float c01, float c11, float fx, float fy, float sx, float sy)
{
    const float top = c00 * sx + c10 * fx;
    const float bot = c01 * sx + c11 * fx;
    return unorm_of(byte_of(top * sy + bot * fy));
}

/* all four channels of one mip level as n / 255 floats; u, v already wrapped */
__device__ __forceinline__ float4 fast_level(const float4 *tex, const StagedTex &st, uint32_t kind, bool linear, bool rep_s, bool rep_t, float u, float v)
{
    if (kind == 2u) return make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    if (kind == 3u) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const float4 *px = kind ? tex + st.n0 : tex;
    const int w = kind ? st.w1 : st.w, h = kind ? st.h1 : st.h;
    if (!linear) {          /* nearest: floor(u w), clamped, never wrapped (textures.c:394-403) */
        int x = f2i_x86(floorf(u * (float)w - 0.5f + 0.5f)), y = f2i_x86(floorf(v * (float)h - 0.5f + 0.5f));
        x = min(max(x, 0), w - 1); y = min(max(y, 0), h - 1);
        return px[y * w + x];
    }
    FTaps T;
    fast_taps(T, px, w, h, rep_s, rep_t, u, v);
    const float sx = 1.0f - T.fx, sy = 1.0f - T.fy;
    return make_float4(fast_channel(T.t00.x, T.t10.x, T.t01.x, T.t11.x, T.fx, T.fy, sx, sy),
                       fast_channel(T.t00.y, T.t10.y, T.t01.y, T.t11.y, T.fx, T.fy, sx, sy),
                       fast_channel(T.t00.z, T.t10.z, T.t01.z, T.t11.z, T.fx, T.fy, sx, sy),
                       fast_channel(T.t00.w, T.t10.w, T.t01.w, T.t11.w, T.fx, T.fy, sx, sy));
}

/* texture_sample_lod for one fragment with the triangle's plan: the texel as four n / 255 floats.  u, v unwrapped. */
__device__ __forceinline__ float4 fast_sample(const float4 *tex, const StagedTex &st, uint32_t plan, float cl, float u, float v)
{
    const bool rep_s = (plan & TP_REP_S) != 0u, rep_t = (plan & TP_REP_T) != 0u;
    u = fast_wrap(u, rep_s); v = fast_wrap(v, rep_t);
    float4 t = fast_level(tex, st, plan & TP_A_KIND, (plan & TP_A_LINEAR) != 0u, rep_s, rep_t, u, v);
    if (plan & TP_TRI) {        /* textures.c:512-515: per channel, truncated to 8 bits once more */
        const float4 t1 = fast_level(tex, st, (plan >> TP_B_SHIFT) & 3u, (plan & TP_B_LINEAR) != 0u, rep_s, rep_t, u, v);
        const float s = 1.0f - cl;
        t.x = unorm_of(byte_of(t.x * s + t1.x * cl)); t.y = unorm_of(byte_of(t.y * s + t1.y * cl));
        t.z = unorm_of(byte_of(t.z * s + t1.z * cl)); t.w = unorm_of(byte_of(t.w * s + t1.w * cl));
    }
    return t;
}

} // namespace mtgl_dev_impl

#endif
