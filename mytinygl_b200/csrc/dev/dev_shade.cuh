/*
 * dev_shade.cuh -- colour of one fragment from its record: interpolation, per-fragment lighting, texturing with the
 * general (global-memory) sampler, alpha test, texenv, fog (src/raster.c:581-705).  Shared by k_raster.cu (in-order
 * and resolve paths) and k_shade.cu (the path for fragments that cannot take the staged-texture sampler).
 */
#ifndef MTGL_DEV_SHADE_CUH
#define MTGL_DEV_SHADE_CUH

#include "dev_common.cuh"
#include "dev_texture.cuh"
#include "dev_fragment.cuh"

namespace mtgl_dev_impl {

/* the interpolants of one record (rows 3-9), loaded uniformly per triangle or per lane when shading is deferred */
struct TriAttr {
    float4 col0, col1, col2;
    float u0, v0, u1, v1, u2, v2;
    float w0, w1, w2;
    float ez0, ez1, ez2;
    float lod;
};

__device__ __forceinline__ void load_attr(TriAttr &A, const TriRecord *rec)
{
    const float4 row3 = __ldg(reinterpret_cast<const float4 *>(rec) + 3);
    const float4 row4 = __ldg(reinterpret_cast<const float4 *>(rec) + 4);
    A.col0 = __ldg(reinterpret_cast<const float4 *>(rec) + 5);
    A.col1 = __ldg(reinterpret_cast<const float4 *>(rec) + 6);
    A.col2 = __ldg(reinterpret_cast<const float4 *>(rec) + 7);
    const float4 row8 = __ldg(reinterpret_cast<const float4 *>(rec) + 8);
    const float4 row9 = __ldg(reinterpret_cast<const float4 *>(rec) + 9);
    A.lod = row3.w;
    A.w0 = row4.x; A.w1 = row4.y; A.w2 = row4.z; A.ez0 = row4.w;
    A.u0 = row8.x; A.v0 = row8.y; A.u1 = row8.z; A.v1 = row8.w;
    A.u2 = row9.x; A.v2 = row9.y; A.ez1 = row9.z; A.ez2 = row9.w;
}

/* Colour of one fragment: interpolation, per-fragment lighting, texturing, alpha test, texenv, fog
 * (raster.c:581-705).  Returns false when the alpha test discards the fragment. */
__device__ __forceinline__ bool shade_color(const BatchDev &b, const float *un, uint32_t r, uint32_t state_flags,
                                            const TriAttr &A, const RasterCfg *cfg, float b0, float b1, float b2, Color4 &c)
{
    const uint32_t flags = cfg->flags;
    if (flags & RC_FLAT) c = { A.col2.x, A.col2.y, A.col2.z, A.col2.w };        /* third vertex of the sub-triangle (raster.c:583-585) */
    else {
        c.r = A.col0.x * b0 + A.col1.x * b1 + A.col2.x * b2;
        c.g = A.col0.y * b0 + A.col1.y * b1 + A.col2.y * b2;
        c.b = A.col0.z * b0 + A.col1.z * b1 + A.col2.z * b2;
        c.a = A.col0.w * b0 + A.col1.w * b1 + A.col2.w * b2;
    }

    if (flags & RC_LIGHTING) {                              /* raster.c:592-615 */
        const bool back_facing = (state_flags >> 31) != 0;
        const bool flip = back_facing && (flags & RC_TWO_SIDE);
        if ((flags & RC_PHONG) || flip) {
            const TriEye *eye = b.rec_eye + r;
            float ep[3], en[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                ep[k] = eye->ep0[k] * b0 + eye->ep1[k] * b1 + eye->ep2[k] * b2;
                en[k] = eye->en0[k] * b0 + eye->en1[k] * b1 + eye->en2[k] * b2;
            }
            const mtgl_state *st = b.states + (state_flags & STATE_INDEX_MASK);
            MaterialRegs mat;
            if (flip) {
                en[0] *= -1.0f; en[1] *= -1.0f; en[2] *= -1.0f;
                load_material(mat, &st->material_back);
            } else load_material(mat, &st->material_front);
            c = compute_lighting(st, ep[0], ep[1], ep[2], en[0], en[1], en[2], mat);
        }
    }

    if (flags & RC_TEXTURED) {                              /* raster.c:618-669 */
        float u, v;
        if (flags & RC_PERSPECTIVE) {
            /* u/w, v/w per vertex (raster.c:501-503) */
            const float u0w = A.u0 * A.w0, v0w = A.v0 * A.w0, u1w = A.u1 * A.w1, v1w = A.v1 * A.w1, u2w = A.u2 * A.w2, v2w = A.v2 * A.w2;
            float uw = b0 * u0w + b1 * u1w + b2 * u2w;
            float vw = b0 * v0w + b1 * v1w + b2 * v2w;
            float ow = b0 * A.w0 + b1 * A.w1 + b2 * A.w2;
            float w = 1.0f / ow;
            u = uw * w;
            v = vw * w;
        } else {
            u = b0 * A.u0 + b1 * A.u1 + b2 * A.u2;
            v = b0 * A.v0 + b1 * A.v1 + b2 * A.v2;
        }
        TexTaps T;
        tex_taps(T, cfg, u, v, A.lod);
        Color4 t;
        t.a = tex_channel(T, 24, un);
        /* alpha test exists only here and tests the TEXEL alpha (raster.c:640-643) */
        if ((flags & RC_ALPHA_TEST) && !compare_f(cfg->alpha_func, t.a, cfg->alpha_ref)) return false;
        t.r = tex_channel(T, 0, un); t.g = tex_channel(T, 8, un); t.b = tex_channel(T, 16, un);
        switch (cfg->tex_env_mode) {
        case G_REPLACE: c = t; break;
        case G_DECAL: c = color_lerp_rgb(c, t, t.a); break;
        case G_BLEND: {
            const float *e = cfg->tex_env_color;
            c = { c.r * (1.0f - t.r) + e[0] * t.r, c.g * (1.0f - t.g) + e[1] * t.g, c.b * (1.0f - t.b) + e[2] * t.b, c.a * t.a };
            break;
        }
        case G_ADD: c = { c.r + t.r, c.g + t.g, c.b + t.b, c.a * t.a }; break;
        default: c = { c.r * t.r, c.g * t.g, c.b * t.b, c.a * t.a }; break;
        }
    }

    if (flags & RC_FOG) {                                   /* raster.c:672-705; result alpha = fog colour alpha */
        float fc = b0 * A.ez0 + b1 * A.ez1 + b2 * A.ez2;
        Color4 fogc = { cfg->fog_color[0], cfg->fog_color[1], cfg->fog_color[2], cfg->fog_color[3] };
        c = color_lerp_rgb(fogc, c, fog_factor(cfg, fc));
    }
    return true;
}

} // namespace mtgl_dev_impl

#endif
