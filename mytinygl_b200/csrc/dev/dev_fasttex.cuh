/*
 * dev_fasttex.cuh -- texture_sample_lod (src/textures.c:457-557) from FLOAT4 TEXELS, shared by the tile kernels that shade
 * many fragments (k_fill.cu, k_shade.cu).
 *
 * Small textures keep, next to their RGBA8 words, a float4 copy in HBM in which every texel already is what
 * color_from_rgba32 returns (n / 255, built once at glTexImage2D time: k_tex_f4).  A bilinear tap is then one 16-byte load
 * with no per-channel conversion, where the general sampler (dev_texture.cuh) does a 4-byte load plus four shift / mask /
 * table look-ups per tap.  k_shade reads the copy through L1 (GLOBAL = true); k_fill stages it in shared memory with one
 * bulk-asynchronous copy (cp.async.bulk + mbarrier, below) because it samples it tens of thousands of times per tile.
 * The filter decision, which depends only on the state and the PER-TRIANGLE lod (raster.c:505-529), is made once per
 * triangle (sampler_plan); the 8-bit truncation after every filter stage uses the arithmetic forms of dev_fragment.cuh.
 * Values are the reference's, bit for bit.
 *
 * Preconditions of the fast path (everything else takes the general sampler): the state's texture has a float4 copy and
 * every REPEAT axis a power-of-two size (RasterCfg::fast_tex); u, v, 1/w of the triangle are finite and bounded
 * (STATE_BOUNDED_BIT, attr_bounded), so that no NaN can reach the wrap or the floor.
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

/* where the float4 texels are: the float4 copy in HBM (RasterCfg::tex_f4, read through L1) or its image in shared memory */
struct TexView {
    const float4 *tex;          /* level 0, level 1 right behind it */
    int w, h, w1, h1, n0;       /* level 1 starts at texel n0 = w * h */
};

struct StagedTex {              /* in shared memory, next to the staged texels */
    const float4 *id;           /* the float4 copy that was staged = its identity; NULL: nothing staged */
    int w, h, w1, h1, n0;
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

template <bool GLOBAL>
__device__ __forceinline__ float4 texel_at(const float4 *px, int i)
{
    if (GLOBAL) return __ldg(px + i);
    return px[i];
}

/* ---- bulk-asynchronous copy global -> shared (sm_90+: cp.async.bulk, SASS UBLKCP) completing on an mbarrier ---- */
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");      /* visible to the asynchronous proxy */
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}

/* dst, src 16-byte aligned, bytes a multiple of 16 */
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

/* every consumer thread: wait for phase `parity` of the barrier; a copy that never lands traps instead of hanging the GPU */
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    for (uint32_t spins = 0;; spins++) {
        uint32_t done;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
        if (done) return;
        if (spins > (1u << 24)) __trap();
    }
}

/* one thread: queue the float4 copy of state `tc`'s texture into `tex` (at most STAGED_TEXELS texels) */
__device__ __forceinline__ void stage_texture_async(float4 *tex, StagedTex &st, const RasterCfg *tc, unsigned long long *bar)
{
    const int n0 = tc->tex_w * tc->tex_h, n1 = tc->tex_l1 ? tc->tex_w1 * tc->tex_h1 : 0;
    if (!tc->tex_f4 || n0 + n1 > STAGED_TEXELS) return;
    const uint32_t bytes = (uint32_t)(n0 + n1) * (uint32_t)sizeof(float4);
    mbar_expect_tx(bar, bytes);
    for (uint32_t off = 0; off < bytes; off += 16384u)
        bulk_copy_g2s(reinterpret_cast<unsigned char *>(tex) + off, reinterpret_cast<const unsigned char *>(tc->tex_f4) + off, min(16384u, bytes - off), bar);
    st.id = tc->tex_f4; st.w = tc->tex_w; st.h = tc->tex_h; st.w1 = tc->tex_w1; st.h1 = tc->tex_h1; st.n0 = n0;
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
template <bool GLOBAL>
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
    T.t00 = texel_at<GLOBAL>(px, ya + xa); T.t10 = texel_at<GLOBAL>(px, ya + xb);
    T.t01 = texel_at<GLOBAL>(px, yb + xa); T.t11 = texel_at<GLOBAL>(px, yb + xb);
}

/* bilinear_filter (textures.c:294-307) of one channel, truncated to 8 bits, as the float n / 255 */
__device__ __forceinline__ float fast_channel(float c00, float c10, float c01, float c11, float fx, float fy, float sx, float sy)
{
    const float top = c00 * sx + c10 * fx;
    const float bot = c01 * sx + c11 * fx;
    return unorm_of(byte_of(top * sy + bot * fy));
}

/* all four channels of one mip level as n / 255 floats; u, v already wrapped */
template <bool GLOBAL>
__device__ __forceinline__ float4 fast_level(const TexView &tv, uint32_t kind, bool linear, bool rep_s, bool rep_t, float u, float v)
{
    if (kind == 2u) return make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    if (kind == 3u) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const float4 *px = kind ? tv.tex + tv.n0 : tv.tex;
    const int w = kind ? tv.w1 : tv.w, h = kind ? tv.h1 : tv.h;
    if (!linear) {          /* nearest: floor(u w), clamped, never wrapped (textures.c:394-403) */
        int x = f2i_x86(floorf(u * (float)w - 0.5f + 0.5f)), y = f2i_x86(floorf(v * (float)h - 0.5f + 0.5f));
        x = min(max(x, 0), w - 1); y = min(max(y, 0), h - 1);
        return texel_at<GLOBAL>(px, y * w + x);
    }
    FTaps T;
    fast_taps<GLOBAL>(T, px, w, h, rep_s, rep_t, u, v);
    const float sx = 1.0f - T.fx, sy = 1.0f - T.fy;
    return make_float4(fast_channel(T.t00.x, T.t10.x, T.t01.x, T.t11.x, T.fx, T.fy, sx, sy),
                       fast_channel(T.t00.y, T.t10.y, T.t01.y, T.t11.y, T.fx, T.fy, sx, sy),
                       fast_channel(T.t00.z, T.t10.z, T.t01.z, T.t11.z, T.fx, T.fy, sx, sy),
                       fast_channel(T.t00.w, T.t10.w, T.t01.w, T.t11.w, T.fx, T.fy, sx, sy));
}

/* texture_sample_lod for one fragment with the triangle's plan: the texel as four n / 255 floats.  u, v unwrapped. */
template <bool GLOBAL>
__device__ __forceinline__ float4 fast_sample(const TexView &tv, uint32_t plan, float cl, float u, float v)
{
    const bool rep_s = (plan & TP_REP_S) != 0u, rep_t = (plan & TP_REP_T) != 0u;
    u = fast_wrap(u, rep_s); v = fast_wrap(v, rep_t);
    float4 t = fast_level<GLOBAL>(tv, plan & TP_A_KIND, (plan & TP_A_LINEAR) != 0u, rep_s, rep_t, u, v);
    if (plan & TP_TRI) {        /* textures.c:512-515: per channel, truncated to 8 bits once more */
        const float4 t1 = fast_level<GLOBAL>(tv, (plan >> TP_B_SHIFT) & 3u, (plan & TP_B_LINEAR) != 0u, rep_s, rep_t, u, v);
        const float s = 1.0f - cl;
        t.x = unorm_of(byte_of(t.x * s + t1.x * cl)); t.y = unorm_of(byte_of(t.y * s + t1.y * cl));
        t.z = unorm_of(byte_of(t.z * s + t1.z * cl)); t.w = unorm_of(byte_of(t.w * s + t1.w * cl));
    }
    return t;
}

} // namespace mtgl_dev_impl

#endif
