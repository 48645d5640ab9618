/*
 * k_fill.cu -- K4d: the in-order tile kernel for LARGE triangles (the fill-rate case, BASELINE config C3: 64 full-screen
 * quads with texture, alpha test, stencil and blending).
 *
 * Replaces, for the tiles it owns, the scan loop and the per-fragment block of rasterize_triangle_smooth
 * (src/raster.c:532-724) with edge_function (299-302), depth / alpha / stencil tests and ops (344-448), blending
 * (360-387), write_pixel_masked (20-45), texture_sample_lod (src/textures.c:457-557) and glClear (src/gl_api.c:409-457)
 * -- the same functions as k_raster<false> (k_raster.cu), with the work turned inside out:
 *
 *   k_raster<false>: a warp owns a 16x16 region, walks the sorted list and visits every 8x4 block of every triangle's
 *                    box; the tile's planes live in shared memory and every fragment reads its triangle's 160-byte
 *                    record from global memory (128 registers, ~580 instructions per fragment on C3).
 *   k_fill:          a THREAD owns pixels.  Colour, depth and stencil of its pixels stay in registers while it walks the
 *                    sorted list; what a triangle contributes to every pixel -- edge coefficients relative to the tile,
 *                    u/w, v/w, colours, the resolved sampler plan -- is prepared once per (triangle, tile) by one thread
 *                    and broadcast from shared memory; the texture is staged in shared memory as float4 texels, i.e.
 *                    already divided by 255, so a bilinear tap is one 16-byte LDS; the two 8-bit conversions that follow
 *                    every filter stage are done with two FP32 adds / an FMA pair (dev_fragment.cuh) instead of F2I and
 *                    a bank-conflicting table look-up.
 *
 * Work decomposition: one CTA (8 warps) per 64x64 tile; in pass g warp w owns row 8 g + w of the tile and lane l the
 * pixels x = l and x = l + 32 of that row: two independent fragments per thread for instruction-level parallelism,
 * 128-byte coalesced plane accesses per warp, and a triangle's rows are spread evenly over the warps.  The list is sorted
 * by submission id (shared memory rank sort), then handled in windows of 64 prepared triangles.
 *
 * Edge functions: when every product and every edge value over the tile is an integer below 2^24 (always for a
 * 3840x2160 full-screen triangle), the reference's float expression is exact, so e = A x + B y + C with tile-relative
 * integer x, y evaluated with FMAs gives the same bits; a triangle too large for that (7680x4320) uses the reference's
 * expression operation for operation.
 *
 * A tile goes to this kernel when its list is in-order, holds only triangles and its mean clamped box covers at
 * least a quarter of the tile (dev_fill.cuh); everything else stays with k_raster<false>.  States with per-fragment
 * lighting are not handled here (the host does not launch the kernel for such batches).
 *
 * Roofline: algorithmically HBM-bound (stencil 2 B, depth 4 (+4) B, blend read 4 B, colour write 4 B per
 * fragment in the reference's immediate-mode formulation); here the planes are read and written once per tile and
 * window, and the kernel is bound by FP32 issue.
 */
#include "dev_common.cuh"
#include "dev_texture.cuh"
#include "dev_fragment.cuh"
#include "dev_fasttex.cuh"
#include "dev_fill.cuh"

namespace mtgl_dev_impl {

void note_launch();

constexpr int FILL_THREADS = 256;
constexpr int FILL_WINDOW = 64;                         /* prepared triangles per window */
constexpr int FILL_PX = 2;                              /* pixels per thread: x = lane and x = lane + 32 of one row */

/* bits 27..31 of PrepTri::eb.w (bits 0..26 = state / cfg index) */
constexpr uint32_t PT_EXACT = 1u << 27;         /* e = A x + B y + C is exact over the tile */
constexpr uint32_t PT_NEG = 1u << 28;           /* (inexact form only) negative area: edge values are negated */
constexpr uint32_t PT_FASTTEX = 1u << 29;       /* textured from the staged texture (dev_fasttex.cuh) */
constexpr uint32_t PT_COINCIDENT = 1u << 30;    /* same geometry, texture coordinates and sampler as the previous triangle of the window */
constexpr uint32_t PT_COINCIDENT2 = 1u << 31;   /* ... as the triangle before the previous one (the two halves of stacked quads alternate) */

/* what one triangle contributes to every pixel of the tile: 14 x 16 B, read as broadcasts */
struct __align__(16) PrepTri {
    float4 ea;      /* A0 A1 A2 | 1/area            (all negated for clockwise triangles: one inclusive test e >= 0) */
    float4 eb;      /* B0 B1 B2 | cfg index + PT_*  */
    float4 ec;      /* C0 C1 C2 | box x0 | y0 << 8 | x1 << 16 | y1 << 24 (tile-relative, inclusive) */
    float4 zz;      /* z0 z1 z2 | lod */
    float4 c0, c1, c2;
    float4 tu;      /* u0 u1 u2 (times 1/w when perspective-correct, raster.c:501-503) | 1/w0 */
    float4 tv;      /* v0 v1 v2 (likewise) | 1/w1 */
    float4 te;      /* eye z0 z1 z2 | 1/w2 */
    float4 p0;      /* x0 y0 x1 y1 as floats: the reference form of the edge functions */
    float4 p1;      /* x2 y2 | sampler plan | trilinear weight */
    uint4 s0;       /* state, decoded once per (triangle, tile): RasterCfg flags | PS_* word | masked stencil reference | zpass op: hi */
    uint4 s1;       /* stencil fail op | stencil zfail op | alpha reference (float bits) | blend_src << 16 | blend_dst */
    int4 so;        /* the stencil zpass op, decoded (dev_fragment.cuh StencilOp): and-mask | xor-mask | add | lo */
};
static_assert(sizeof(PrepTri) == 240, "PrepTri layout");

/* PrepTri::s0.y: comparison masks (dev_fragment.cuh, compare_mask) and small enums */
constexpr uint32_t PS_STENCIL_CMP_SHIFT = 0, PS_DEPTH_CMP_SHIFT = 4, PS_ALPHA_CMP_SHIFT = 8;     /* 4 bits each */
constexpr uint32_t PS_STENCIL_MASK_SHIFT = 12, PS_STENCIL_WMASK_SHIFT = 20;                       /* 8 bits each */
constexpr uint32_t PS_COLOR_MASK_SHIFT = 28;                                                      /* 4 bits */

struct FillSmem {
    float4 tex[STAGED_TEXELS];
    PrepTri tri[FILL_WINDOW];
    uint32_t key[FILL_MAX_LIST];
    uint32_t rec[FILL_MAX_LIST];
    uint32_t sorted[FILL_MAX_LIST];
    float un[256];
    uint8_t cls[FILL_WINDOW];       /* bits 0-5: the triangle's geometry class = index of the first triangle of the window it coincides
                                     * with (through the chain of PT_COINCIDENT / PT_COINCIDENT2 links); bit 7: the class has one member */
    uint32_t acc;               /* scratch of fill_owns_tile */
    uint32_t tex_cfg;           /* state index whose texture is staged (lowest textured state of the list), ~0 = none */
    StagedTex st;
    __align__(8) unsigned long long tex_bar;    /* mbarrier the bulk copy of the texture completes on */
};
/* two CTAs per SM */
static_assert(2 * (sizeof(FillSmem) + 1024) <= 227 * 1024, "k_fill: two tiles per SM");

/* ---------------------------------------------------------------- per-tile preparation */
__device__ __forceinline__ uint32_t tile_box(uint32_t bbox_min, uint32_t bbox_max, int px0, int py0)
{
    const int x0 = max((int)(bbox_min & 0xFFFFu) - px0, 0), y0 = max((int)(bbox_min >> 16) - py0, 0);
    const int x1 = min((int)(bbox_max & 0xFFFFu) - px0, TILE_W - 1), y1 = min((int)(bbox_max >> 16) - py0, TILE_H - 1);
    return (uint32_t)x0 | ((uint32_t)y0 << 8) | ((uint32_t)x1 << 16) | ((uint32_t)y1 << 24);
}

__device__ __forceinline__ bool same16(const uint4 &a, const uint4 &b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }

/* Everything a triangle hands to the per-pixel stages BEFORE its own colours, depth and per-fragment state come into play
 * -- coverage, barycentrics, texture coordinates, the sampled texel -- is a pure function of its snapped vertices, its
 * clamped box, its per-vertex (u, v, 1/w), its LOD and its sampler state.  Multi-pass rendering draws the same geometry
 * again and again (C3: 64 coincident quads); a triangle that repeats all of those inputs of its predecessor in the
 * sorted list is marked PT_COINCIDENT and reuses the predecessor's per-pixel intermediates (fill_shared). */
__device__ __forceinline__ bool coincident(const BatchDev &b, const TriRecord *rec, const TriRecord *prev, const RasterCfg *cfg)
{
    const uint4 *a = reinterpret_cast<const uint4 *>(rec), *p = reinterpret_cast<const uint4 *>(prev);
    const uint4 a0 = __ldg(a + 0), p0 = __ldg(p + 0), a1 = __ldg(a + 1), p1 = __ldg(p + 1), a2 = __ldg(a + 2), p2 = __ldg(p + 2);
    if (!same16(a0, p0) || !same16(a1, p1) || a2.x != p2.x || a2.y != p2.y) return false;         /* vertices, area, box */
    if ((a2.z & STATE_KIND_MASK) != (p2.z & STATE_KIND_MASK)) return false;
    const RasterCfg *pc = b.cfgs + (p2.z & STATE_INDEX_MASK);
    const uint32_t tf = RC_TEXTURED | RC_PERSPECTIVE;
    if ((cfg->flags & tf) != (pc->flags & tf)) return false;
    if (!(cfg->flags & RC_TEXTURED)) return true;
    if (cfg->tex_l0 != pc->tex_l0 || cfg->tex_l1 != pc->tex_l1 || cfg->tex_min != pc->tex_min || cfg->tex_mag != pc->tex_mag ||
        cfg->tex_wrap_s != pc->tex_wrap_s || cfg->tex_wrap_t != pc->tex_wrap_t) return false;
    const uint4 a4 = __ldg(a + 4), p4 = __ldg(p + 4), a8 = __ldg(a + 8), p8 = __ldg(p + 8), a9 = __ldg(a + 9), p9 = __ldg(p + 9);
    const uint32_t alod = __ldg(&reinterpret_cast<const uint32_t *>(rec)[15]), plod = __ldg(&reinterpret_cast<const uint32_t *>(prev)[15]);
    return a4.x == p4.x && a4.y == p4.y && a4.z == p4.z && same16(a8, p8) && a9.x == p9.x && a9.y == p9.y && alod == plod;   /* 1/w, u, v, lod (bitwise) */
}

__device__ void prep_triangle(PrepTri &P, const BatchDev &b, const FillSmem &sm, uint32_t r, uint32_t r_prev, uint32_t r_prev2, int px0, int py0)
{
    const TriRecord *rec = b.records + r;
    const int4 row0 = __ldg(reinterpret_cast<const int4 *>(rec) + 0);
    const int4 row1 = __ldg(reinterpret_cast<const int4 *>(rec) + 1);
    const uint4 row2 = __ldg(reinterpret_cast<const uint4 *>(rec) + 2);
    const float4 row3 = __ldg(reinterpret_cast<const float4 *>(rec) + 3);
    const float4 row4 = __ldg(reinterpret_cast<const float4 *>(rec) + 4);
    const float4 row8 = __ldg(reinterpret_cast<const float4 *>(rec) + 8);
    const float4 row9 = __ldg(reinterpret_cast<const float4 *>(rec) + 9);
    P.c0 = __ldg(reinterpret_cast<const float4 *>(rec) + 5);
    P.c1 = __ldg(reinterpret_cast<const float4 *>(rec) + 6);
    P.c2 = __ldg(reinterpret_cast<const float4 *>(rec) + 7);

    const uint32_t cfg_index = row2.z & STATE_INDEX_MASK;
    const RasterCfg *cfg = b.cfgs + cfg_index;
    const uint32_t cflags = cfg->flags;
    const float area = __int_as_float(row1.z), inv_area = __int_as_float(row1.w);
    const bool pos = area > 0;
    uint32_t word = cfg_index;
    if (r_prev != 0xFFFFFFFFu && coincident(b, rec, b.records + r_prev, cfg)) word |= PT_COINCIDENT;
    else if (r_prev2 != 0xFFFFFFFFu && coincident(b, rec, b.records + r_prev2, cfg)) word |= PT_COINCIDENT2;

    /* raster.c:536-538: w0 = edge(v1, v2, p), w1 = edge(v2, v0, p), w2 = edge(v0, v1, p) */
    const int vx[3] = { row0.x, row0.z, row1.x }, vy[3] = { row0.y, row0.w, row1.y };
    float A[3] = { 0.0f, 0.0f, 0.0f }, B[3] = { 0.0f, 0.0f, 0.0f }, C[3] = { 0.0f, 0.0f, 0.0f };
    bool exact = true;
#pragma unroll
    for (int k = 0; k < 3; k++) exact = exact && abs((long long)vx[k]) < (1ll << 24) && abs((long long)vy[k]) < (1ll << 24);
    if (exact) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const long long ax = vx[(k + 1) % 3], ay = vy[(k + 1) % 3], bx = vx[(k + 2) % 3], by = vy[(k + 2) % 3];
            const long long dx = bx - ax, dy = by - ay;
            /* over the tile's pixels: |px - ax| <= mx, |py - ay| <= my; both products and their difference must stay
             * below 2^24 for the float expression of raster.c:299-302 to be exact */
            const long long mx = max(abs((long long)px0 - ax), abs((long long)px0 + TILE_W - 1 - ax));
            const long long my = max(abs((long long)py0 - ay), abs((long long)py0 + TILE_H - 1 - ay));
            if (mx * abs(dy) + my * abs(dx) >= (1ll << 24)) exact = false;
            const long long c = ((long long)px0 - ax) * dy - ((long long)py0 - ay) * dx;       /* the edge value at the tile's origin */
            A[k] = (float)dy; B[k] = (float)-dx; C[k] = (float)c;
        }
    }
    if (exact) {
        word |= PT_EXACT;
        if (!pos) {
#pragma unroll
            for (int k = 0; k < 3; k++) { A[k] = -A[k]; B[k] = -B[k]; C[k] = -C[k]; }
        }
    } else if (!pos) word |= PT_NEG;
    P.ea = make_float4(A[0], A[1], A[2], pos ? inv_area : -inv_area);
    P.ec = make_float4(C[0], C[1], C[2], __uint_as_float(tile_box(row2.x, row2.y, px0, py0)));
    P.zz = row3;
    P.p0 = make_float4((float)row0.x, (float)row0.y, (float)row0.z, (float)row0.w);

    /* u/w, v/w per vertex (raster.c:501-503) */
    float u[3] = { row8.x, row8.z, row9.x }, v[3] = { row8.y, row8.w, row9.y };
    const float w[3] = { row4.x, row4.y, row4.z };
    if ((cflags & RC_TEXTURED) && cfg->fast_tex && sm.st.id != nullptr && cfg->tex_f4 == sm.st.id && (row2.z & STATE_BOUNDED_BIT)) word |= PT_FASTTEX;
    if (cflags & RC_PERSPECTIVE) {
#pragma unroll
        for (int k = 0; k < 3; k++) { u[k] = u[k] * w[k]; v[k] = v[k] * w[k]; }
    }
    P.tu = make_float4(u[0], u[1], u[2], w[0]);
    P.tv = make_float4(v[0], v[1], v[2], w[1]);
    P.te = make_float4(row4.w, row9.z, row9.w, w[2]);
    float cl = 0.0f;
    const uint32_t plan = (cflags & RC_TEXTURED) ? sampler_plan(cfg, row3.w, cl) : 0u;
    P.p1 = make_float4((float)row1.x, (float)row1.y, __uint_as_float(plan), cl);
    P.eb = make_float4(B[0], B[1], B[2], __uint_as_float(word));
    const uint32_t ps = (compare_mask(cfg->stencil_func) << PS_STENCIL_CMP_SHIFT) | (compare_mask(cfg->depth_func) << PS_DEPTH_CMP_SHIFT) |
                        (compare_mask(cfg->alpha_func) << PS_ALPHA_CMP_SHIFT) | ((cfg->stencil_mask & 0xFFu) << PS_STENCIL_MASK_SHIFT) |
                        ((cfg->stencil_writemask & 0xFFu) << PS_STENCIL_WMASK_SHIFT) | ((cfg->color_mask & 0xFu) << PS_COLOR_MASK_SHIFT);
    /* raster.c:407-422 compares (ref & mask) with (value & mask); the value has 8 bits, the reference and the mask need not */
    const StencilOp zp = stencil_op_decode(stencil_op_encode(cfg->stencil_zpass, cfg->stencil_ref));
    P.s0 = make_uint4(cflags, ps, (uint32_t)cfg->stencil_ref & cfg->stencil_mask, (uint32_t)zp.hi);
    P.so = make_int4((int)zp.amask, (int)zp.xmask, zp.add, zp.lo);
    P.s1 = make_uint4(stencil_op_encode(cfg->stencil_fail, cfg->stencil_ref), stencil_op_encode(cfg->stencil_zfail, cfg->stencil_ref),
                      __float_as_uint(cfg->alpha_ref), (cfg->blend_src << 16) | (cfg->blend_dst & 0xFFFFu));
}

/* the general sampler (any texture size, any attribute values) from global memory: dev_texture.cuh, out of line */
__device__ __noinline__ void slow_texel(const RasterCfg *cfg, const float *un, float u, float v, float lod, float4 *out)
{
    TexTaps T;
    tex_taps(T, cfg, u, v, lod);
    *out = make_float4(tex_channel(T, 0, un), tex_channel(T, 8, un), tex_channel(T, 16, un), tex_channel(T, 24, un));
}

/* ---------------------------------------------------------------- one class of coincident triangles over this thread's pixels */
/* Pixel state of a thread: colour channels as integral floats 0..255 (the byte the plane holds), depth, stencil. */
struct PixelState {
    float r[FILL_PX], g[FILL_PX], b[FILL_PX], a[FILL_PX];
    float depth[FILL_PX];
    uint32_t stencil[FILL_PX];
};

/* what the triangles of a geometry class share, per pixel */
struct Shared {
    bool cov[FILL_PX];                                  /* inside the triangle, its box and the framebuffer */
    float b0[FILL_PX], b1[FILL_PX], b2[FILL_PX];        /* barycentrics (raster.c:541-543) */
    float tr[FILL_PX], tg[FILL_PX], tb[FILL_PX], ta[FILL_PX];   /* the sampled texel */
};

/* coverage (raster.c:536-540), barycentrics, texture coordinates (618-637) and the texel of a class, from its first
 * triangle that reaches this thread.  `single`: the class has one member, so its alpha test may discard before the colour
 * channels are filtered.  Returns false -- and leaves H alone -- when none of this thread's pixels is covered. */
template <uint32_t ON, uint32_t OFF>
__device__ __forceinline__ bool fill_shared(const BatchDev &b, const FillSmem &sm, const PrepTri &T, int px0, int py0, int Y,
                                            const int (&X)[FILL_PX], const bool (&inb)[FILL_PX], bool single, Shared &H)
{
    constexpr int P = FILL_PX;
    const uint32_t box = __float_as_uint(T.ec.w);
    if (Y < (int)((box >> 8) & 0xFFu) || Y > (int)(box >> 24)) return false;          /* warp-uniform */
    const int bx0 = (int)(box & 0xFFu), bx1 = (int)((box >> 16) & 0xFFu);
    const uint32_t word = __float_as_uint(T.eb.w);
    const uint32_t fl = T.s0.x;
    auto has = [&](uint32_t bit) -> bool { return (ON & bit) ? true : ((OFF & bit) ? false : (fl & bit) != 0u); };

    float e0[P], e1[P], e2[P];
    if (word & PT_EXACT) {
        const float4 ea = T.ea, eb = T.eb, ec = T.ec;
        const float fy = (float)Y;
        const float t0 = __fmaf_rn(eb.x, fy, ec.x), t1 = __fmaf_rn(eb.y, fy, ec.y), t2 = __fmaf_rn(eb.z, fy, ec.z);
#pragma unroll
        for (int p = 0; p < P; p++) {
            const float fx = (float)X[p];
            e0[p] = __fmaf_rn(ea.x, fx, t0); e1[p] = __fmaf_rn(ea.y, fx, t1); e2[p] = __fmaf_rn(ea.z, fx, t2);
        }
    } else {
        const float4 p0 = T.p0, p1 = T.p1;
        const float py = (float)(py0 + Y);
#pragma unroll
        for (int p = 0; p < P; p++) {
            const float px = (float)(px0 + X[p]);
            e0[p] = edge_at(p0.z, p0.w, p1.x, p1.y, px, py);
            e1[p] = edge_at(p1.x, p1.y, p0.x, p0.y, px, py);
            e2[p] = edge_at(p0.x, p0.y, p0.z, p0.w, px, py);
            if (word & PT_NEG) { e0[p] = -e0[p]; e1[p] = -e1[p]; e2[p] = -e2[p]; }
        }
    }
    bool any = false, cov[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
        cov[p] = inb[p] && X[p] >= bx0 && X[p] <= bx1 && fminf(fminf(e0[p], e1[p]), e2[p]) >= 0.0f;
        any = any || cov[p];
    }
    if (!any) return false;             /* H still holds what it held: the intermediates of another class stay usable */
#pragma unroll
    for (int p = 0; p < P; p++) H.cov[p] = cov[p];
    const float inv_area = T.ea.w;
#pragma unroll
    for (int p = 0; p < P; p++) { H.b0[p] = e0[p] * inv_area; H.b1[p] = e1[p] * inv_area; H.b2[p] = e2[p] * inv_area; }
    if (!has(RC_TEXTURED)) return true;

    const float4 tu = T.tu, tv = T.tv;
    const float w2 = T.te.w;
    float u[P], v[P];
    if (has(RC_PERSPECTIVE)) {
#pragma unroll
        for (int p = 0; p < P; p++) {
            const float uw = H.b0[p] * tu.x + H.b1[p] * tu.y + H.b2[p] * tu.z;
            const float vw = H.b0[p] * tv.x + H.b1[p] * tv.y + H.b2[p] * tv.z;
            const float ow = H.b0[p] * tu.w + H.b1[p] * tv.w + H.b2[p] * w2;
            const float w = 1.0f / ow;
            u[p] = uw * w; v[p] = vw * w;
        }
    } else {
#pragma unroll
        for (int p = 0; p < P; p++) {
            u[p] = H.b0[p] * tu.x + H.b1[p] * tu.y + H.b2[p] * tu.z;
            v[p] = H.b0[p] * tv.x + H.b1[p] * tv.y + H.b2[p] * tv.z;
        }
    }
    if (word & PT_FASTTEX) {
        const uint32_t plan = __float_as_uint(T.p1.z);
        const uint32_t kind_a = plan & TP_A_KIND;
        if (!(plan & TP_TRI) && (plan & TP_A_LINEAR) && kind_a <= 1u) {
            /* one bilinear level (the magnified / non-mipmapped case): alpha first; a lone triangle whose alpha test
             * (raster.c:640-643: on the texel's alpha) discards both fragments never filters the colour channels */
            const bool rep_s = (plan & TP_REP_S) != 0u, rep_t = (plan & TP_REP_T) != 0u;
            const float4 *px = kind_a ? sm.tex + sm.st.n0 : sm.tex;
            const int w = kind_a ? sm.st.w1 : sm.st.w, h = kind_a ? sm.st.h1 : sm.st.h;      /* (the staged copy: shared-memory loads) */
            FTaps F[P];
            float sx[P], sy[P];
#pragma unroll
            for (int p = 0; p < P; p++) {
                fast_taps<false>(F[p], px, w, h, rep_s, rep_t, fast_wrap(u[p], rep_s), fast_wrap(v[p], rep_t));
                sx[p] = 1.0f - F[p].fx; sy[p] = 1.0f - F[p].fy;
                H.ta[p] = fast_channel(F[p].t00.w, F[p].t10.w, F[p].t01.w, F[p].t11.w, F[p].fx, F[p].fy, sx[p], sy[p]);
            }
            if (single && has(RC_ALPHA_TEST)) {
                const uint32_t acmp = (T.s0.y >> PS_ALPHA_CMP_SHIFT) & 15u;
                const float aref = __uint_as_float(T.s1.z);
                any = false;
#pragma unroll
                for (int p = 0; p < P; p++) any = any || (H.cov[p] && ((ON & FILL_ALPHA_GREATER) ? (H.ta[p] > aref) : compare_f_mask(acmp, H.ta[p], aref)));
                if (!any) {         /* the stencil / depth stages of the triangle still run (fill_one), the colour stages cannot be reached */
#pragma unroll
                    for (int p = 0; p < P; p++) { H.tr[p] = 0.0f; H.tg[p] = 0.0f; H.tb[p] = 0.0f; }
                    return true;
                }
            }
#pragma unroll
            for (int p = 0; p < P; p++) {
                H.tr[p] = fast_channel(F[p].t00.x, F[p].t10.x, F[p].t01.x, F[p].t11.x, F[p].fx, F[p].fy, sx[p], sy[p]);
                H.tg[p] = fast_channel(F[p].t00.y, F[p].t10.y, F[p].t01.y, F[p].t11.y, F[p].fx, F[p].fy, sx[p], sy[p]);
                H.tb[p] = fast_channel(F[p].t00.z, F[p].t10.z, F[p].t01.z, F[p].t11.z, F[p].fx, F[p].fy, sx[p], sy[p]);
            }
        } else {
            const float cl = T.p1.w;
            const TexView tv = { sm.tex, sm.st.w, sm.st.h, sm.st.w1, sm.st.h1, sm.st.n0 };
#pragma unroll
            for (int p = 0; p < P; p++) {
                const float4 t = fast_sample<false>(tv, plan, cl, u[p], v[p]);
                H.tr[p] = t.x; H.tg[p] = t.y; H.tb[p] = t.z; H.ta[p] = t.w;
            }
        }
    } else {
        const RasterCfg *cfg = b.cfgs + (word & STATE_INDEX_MASK);
        const float lod = T.zz.w;
#pragma unroll
        for (int p = 0; p < P; p++) {
            H.tr[p] = H.tg[p] = H.tb[p] = H.ta[p] = 0.0f;
            if (!H.cov[p]) continue;
            float4 t;
            slow_texel(cfg, sm.un, u[p], v[p], lod, &t);
            H.tr[p] = t.x; H.tg[p] = t.y; H.tb[p] = t.z; H.ta[p] = t.w;
        }
    }
    return true;
}

/* One triangle of the class H holds: depth value, stencil test + ops, depth test (raster.c:546-587), colour (581-591), alpha test
 * (640-643), texenv (645-669), fog (672-705), late depth write (707-710), blending (712-717), masked write (719-721, 20-45). */
template <uint32_t ON, uint32_t OFF>
__device__ __forceinline__ void fill_one(const BatchDev &b, const PrepTri &T, const Shared &H, PixelState &S)
{
    constexpr int P = FILL_PX;
    const uint4 s0 = T.s0;
    const uint32_t fl = s0.x, ps = s0.y;
    auto has = [&](uint32_t bit) -> bool { return (ON & bit) ? true : ((OFF & bit) ? false : (fl & bit) != 0u); };
    const RasterCfg *cfg = b.cfgs + (__float_as_uint(T.eb.w) & STATE_INDEX_MASK);      /* only the rarely used fields are read from it */

    bool act[P];
#pragma unroll
    for (int p = 0; p < P; p++) act[p] = H.cov[p];
    float depth[P];
#pragma unroll
    for (int p = 0; p < P; p++) depth[p] = 0.0f;
    const bool depth_test = has(RC_DEPTH_TEST);
    const uint32_t depth_cmp = (ps >> PS_DEPTH_CMP_SHIFT) & 15u;
    if (depth_test) {
        const float4 zz = T.zz;
        const bool r01 = has(RC_DEPTH_RANGE_01);
#pragma unroll
        for (int p = 0; p < P; p++) {
            const float z = H.b0[p] * zz.x + H.b1[p] * zz.y + H.b2[p] * zz.z;
            if (r01) depth[p] = (z + 1.0f) * 0.5f;
            else depth[p] = (float)((double)((z + 1.0f) * 0.5f) * (cfg->depth_far - cfg->depth_near) + cfg->depth_near);   /* raster.c:548 */
        }
    }
    if (has(RC_STENCIL)) {
        const uint4 s1 = T.s1;
        const uint32_t scmp = (ps >> PS_STENCIL_CMP_SHIFT) & 15u, smask = (ps >> PS_STENCIL_MASK_SHIFT) & 0xFFu, swm = (ps >> PS_STENCIL_WMASK_SHIFT) & 0xFFu;
        const int32_t mref = (int32_t)s0.z;
        const int4 so = T.so;
        const StencilOp zpass_op = { (uint32_t)so.x, (uint32_t)so.y, so.z, so.w, (int)s0.w };
        if (((ON & FILL_STENCIL_ALWAYS) || (scmp & 7u) == 7u) && !depth_test) {      /* GL_ALWAYS without a depth test: every covered fragment takes the zpass op */
#pragma unroll
            for (int p = 0; p < P; p++) {
                const uint32_t sval = S.stencil[p];
                if (act[p]) S.stencil[p] = (sval & ~swm) | (stencil_op_apply(zpass_op, sval) & swm);
            }
        } else {
            const StencilOp fail_op = stencil_op_decode(s1.x), zfail_op = stencil_op_decode(s1.y);
#pragma unroll
            for (int p = 0; p < P; p++) {
                const uint32_t sval = S.stencil[p];
                const bool spass = compare_i_mask(scmp, mref, (int32_t)(sval & smask));
                const bool zpass = !depth_test || compare_f_mask(depth_cmp, depth[p], S.depth[p]);
                const uint32_t nv = !spass ? stencil_op_apply(fail_op, sval) : (!zpass ? stencil_op_apply(zfail_op, sval) : stencil_op_apply(zpass_op, sval));
                if (act[p]) S.stencil[p] = (sval & ~swm) | (nv & swm);
                act[p] = act[p] && spass && zpass;
            }
        }
    } else if (depth_test) {
#pragma unroll
        for (int p = 0; p < P; p++) act[p] = act[p] && compare_f_mask(depth_cmp, depth[p], S.depth[p]);
    }
    const bool textured = has(RC_TEXTURED);
    if (textured && has(RC_ALPHA_TEST)) {       /* on the texel's alpha, and only when textured (raster.c:640-643) */
        const uint32_t acmp = (ps >> PS_ALPHA_CMP_SHIFT) & 15u;
        const float aref = __uint_as_float(T.s1.z);
#pragma unroll
        for (int p = 0; p < P; p++) act[p] = act[p] && ((ON & FILL_ALPHA_GREATER) ? (H.ta[p] > aref) : compare_f_mask(acmp, H.ta[p], aref));
    }
    bool any = false;
#pragma unroll
    for (int p = 0; p < P; p++) any = any || act[p];
    if (!any) return;
    const bool depth_write = depth_test && has(RC_DEPTH_WRITE);

    float cr[P], cg[P], cb[P], ca[P];
    {
        const float4 c2 = T.c2;
        if (has(RC_FLAT)) {
#pragma unroll
            for (int p = 0; p < P; p++) { cr[p] = c2.x; cg[p] = c2.y; cb[p] = c2.z; ca[p] = c2.w; }
        } else {
            const float4 c0 = T.c0, c1 = T.c1;
#pragma unroll
            for (int p = 0; p < P; p++) {
                cr[p] = c0.x * H.b0[p] + c1.x * H.b1[p] + c2.x * H.b2[p];
                cg[p] = c0.y * H.b0[p] + c1.y * H.b1[p] + c2.y * H.b2[p];
                cb[p] = c0.z * H.b0[p] + c1.z * H.b1[p] + c2.z * H.b2[p];
                ca[p] = c0.w * H.b0[p] + c1.w * H.b1[p] + c2.w * H.b2[p];
            }
        }
    }
    if (textured) {
        const uint32_t env = (ON & FILL_MODULATE) ? (uint32_t)G_MODULATE : cfg->tex_env_mode;
        if (env == G_MODULATE) {
#pragma unroll
            for (int p = 0; p < P; p++) { cr[p] = cr[p] * H.tr[p]; cg[p] = cg[p] * H.tg[p]; cb[p] = cb[p] * H.tb[p]; ca[p] = ca[p] * H.ta[p]; }
        } else {
#pragma unroll
            for (int p = 0; p < P; p++) {
                const float tr = H.tr[p], tg = H.tg[p], tb = H.tb[p], ta = H.ta[p];
                switch (env) {
                case G_REPLACE: cr[p] = tr; cg[p] = tg; cb[p] = tb; ca[p] = ta; break;
                case G_DECAL: cr[p] = cr[p] + (tr - cr[p]) * ta; cg[p] = cg[p] + (tg - cg[p]) * ta; cb[p] = cb[p] + (tb - cb[p]) * ta; break;
                case G_BLEND: {
                    const float *ec = cfg->tex_env_color;
                    cr[p] = cr[p] * (1.0f - tr) + ec[0] * tr; cg[p] = cg[p] * (1.0f - tg) + ec[1] * tg;
                    cb[p] = cb[p] * (1.0f - tb) + ec[2] * tb; ca[p] = ca[p] * ta;
                    break;
                }
                case G_ADD: cr[p] = cr[p] + tr; cg[p] = cg[p] + tg; cb[p] = cb[p] + tb; ca[p] = ca[p] * ta; break;
                default: cr[p] = cr[p] * tr; cg[p] = cg[p] * tg; cb[p] = cb[p] * tb; ca[p] = ca[p] * ta; break;
                }
            }
        }
    }
    if (has(RC_FOG)) {
        const float4 te = T.te;
        const float fr = cfg->fog_color[0], fg = cfg->fog_color[1], fbl = cfg->fog_color[2], fa = cfg->fog_color[3];
#pragma unroll
        for (int p = 0; p < P; p++) {
            const float f = fog_factor(cfg, H.b0[p] * te.x + H.b1[p] * te.y + H.b2[p] * te.z);
            cr[p] = fr + (cr[p] - fr) * f; cg[p] = fg + (cg[p] - fg) * f; cb[p] = fbl + (cb[p] - fbl) * f; ca[p] = fa;   /* color_lerp_rgb(fog, c, f) */
        }
    }

    const bool blend = has(RC_BLEND);
    const uint32_t bfunc = T.s1.w, bsrc = bfunc >> 16, bdst = bfunc & 0xFFFFu;
    const bool blend_alpha = (ON & FILL_BLEND_ALPHA) || bfunc == ((G_SRC_ALPHA << 16) | G_ONE_MINUS_SRC_ALPHA);       /* the usual transparency blend, without the switches */
    const uint32_t cm = (ON & FILL_FULL_MASK) ? 0xFu : ps >> PS_COLOR_MASK_SHIFT;
#pragma unroll
    for (int p = 0; p < P; p++) {
        if (!act[p]) continue;
        if (depth_write) S.depth[p] = depth[p];
        Color4 c = { cr[p], cg[p], cb[p], ca[p] };
        if (blend) {
            const Color4 d = { unorm_of(S.r[p]), unorm_of(S.g[p]), unorm_of(S.b[p]), unorm_of(S.a[p]) };
            /* blend_colors clamps, raster.c:719 clamps again, color_to_rgba32 clamps a third time: byte_of saturates once */
            if (blend_alpha) {
                const float sa = c.a, da = 1 - c.a;
                c = { c.r * sa + d.r * da, c.g * sa + d.g * da, c.b * sa + d.b * da, c.a * sa + d.a * da };
            } else {
                const Color4 sf = blend_factor(bsrc, c, d), df = blend_factor(bdst, c, d);
                c = { c.r * sf.r + d.r * df.r, c.g * sf.g + d.g * df.g, c.b * sf.b + d.b * df.b, c.a * sf.a + d.a * df.a };
            }
        }
        if (cm == 0xFu) { S.r[p] = byte_of(c.r); S.g[p] = byte_of(c.g); S.b[p] = byte_of(c.b); S.a[p] = byte_of(c.a); }
        else if (cm != 0u) {
            /* partial mask: the reference unpacks the pixel, replaces the enabled channels and packs ALL of them again */
            S.r[p] = byte_of((cm & 1u) ? c.r : unorm_of(S.r[p]));
            S.g[p] = byte_of((cm & 2u) ? c.g : unorm_of(S.g[p]));
            S.b[p] = byte_of((cm & 4u) ? c.b : unorm_of(S.b[p]));
            S.a[p] = byte_of((cm & 8u) ? c.a : unorm_of(S.a[p]));
        }
    }
}

/* ---------------------------------------------------------------- the kernel */
template <uint32_t ON, uint32_t OFF>
__global__ void __launch_bounds__(FILL_THREADS, 2) k_fill(BatchDev b, FrameTargets fb, ClearOp clr, uint32_t planes, uint32_t fill_mode, uint32_t split)
{
    extern __shared__ __align__(16) unsigned char fill_smem_raw[];
    FillSmem &sm = *reinterpret_cast<FillSmem *>(fill_smem_raw);
    if (!lists_fit(b) || !b.tile_count) return;

    /* split CTAs per tile, TILE_H / split rows each: a grid below one wave (the band of a multi-GPU frame) lasts as long
     * as its heaviest tile, so there the tile is cut into up to eight slices that sort and prepare the same list */
    const uint32_t slot = blockIdx.x / split, sub = blockIdx.x % split;
    const uint32_t tile = b.tile_order ? b.tile_order[slot] : slot;
    const int tx = (int)(tile % (uint32_t)fb.tiles_x), ty = (int)(tile / (uint32_t)fb.tiles_x) + fb.tile_y0;
    const int px0 = tx << TILE_LOG, py0t = ty << TILE_LOG;
    const int py0 = max(py0t, fb.band_y0);          /* row 0 of the tile's coordinate system: its first row inside the band */
    const int vw = min(TILE_W, fb.width - px0);
    const int vh = min(py0t + TILE_H, fb.band_y1) - py0;
    if (vw <= 0 || vh <= 0) return;
    const uint32_t L = b.tile_count[tile];
    if (L == 0u) return;
    const uint32_t *list = b.tile_list + b.tile_offset[tile];
    if (!fill_owns_tile(b, fill_mode, L, b.tile_flags[tile], list, px0, py0, vh, &sm.acc)) return;

    /* ---- list -> shared memory, sorted by submission id; the texture to stage ---- */
    sm.un[threadIdx.x] = b.unorm8[threadIdx.x];
    if (threadIdx.x == 0) { sm.tex_cfg = 0xFFFFFFFFu; sm.st.id = nullptr; mbar_init(&sm.tex_bar, 1u); }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < L; i += FILL_THREADS) {
        const uint32_t r = list[i];
        const uint4 row = __ldg(b.bin_rows + r);
        sm.key[i] = row.w;
        sm.rec[i] = r;
        const uint32_t ci = row.z & STATE_INDEX_MASK;
        if (b.cfgs[ci].flags & RC_TEXTURED) atomicMin(&sm.tex_cfg, ci);
    }
    __syncthreads();
    /* the texture goes to shared memory as one bulk-asynchronous copy of its float4 image while the list is being sorted */
    if (threadIdx.x == 0 && sm.tex_cfg != 0xFFFFFFFFu) stage_texture_async(sm.tex, sm.st, b.cfgs + sm.tex_cfg, &sm.tex_bar);
    for (uint32_t i = threadIdx.x; i < L; i += FILL_THREADS) {      /* ids are unique: the rank is the sorted position */
        const uint32_t mine = sm.key[i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < L; j++) rank += (sm.key[j] < mine) ? 1u : 0u;
        sm.sorted[rank] = sm.rec[i];
    }

    /* ---- clear rectangle relative to the tile (gl_api.c:409-457) ---- */
    const int cx0 = max(clr.x0 - px0, 0), cy0 = max(clr.y0 - py0, 0);
    const int cx1 = min(clr.x1 - px0, vw), cy1 = min(clr.y1 - py0, vh);
    const bool clr_any = clr.mask && cx0 < cx1 && cy0 < cy1;

    const int lane = (int)(threadIdx.x & 31), warp = (int)(threadIdx.x >> 5);
    int X[FILL_PX];
#pragma unroll
    for (int p = 0; p < FILL_PX; p++) X[p] = lane + 32 * p;

    for (uint32_t w0 = 0; w0 < L; w0 += FILL_WINDOW) {
        const uint32_t n = min((uint32_t)FILL_WINDOW, L - w0);
        __syncthreads();                /* sorted list complete, texture queued; the previous window is no longer read */
        if (w0 == 0u && sm.st.id != nullptr) mbar_wait(&sm.tex_bar, 0u);        /* the staged texels have landed */
        if (threadIdx.x < n)        /* (a class of coincident triangles does not continue across windows) */
            prep_triangle(sm.tri[threadIdx.x], b, sm, sm.sorted[w0 + threadIdx.x], threadIdx.x ? sm.sorted[w0 + threadIdx.x - 1] : 0xFFFFFFFFu,
                          threadIdx.x > 1u ? sm.sorted[w0 + threadIdx.x - 2] : 0xFFFFFFFFu, px0, py0);
        __syncthreads();
        if (threadIdx.x < n) {      /* geometry classes: follow the links back to the first triangle of the chain */
            auto link = [&](uint32_t t) -> uint32_t { const uint32_t wd = __float_as_uint(sm.tri[t].eb.w); return (wd & PT_COINCIDENT) ? 1u : ((wd & PT_COINCIDENT2) ? 2u : 0u); };
            const uint32_t t = threadIdx.x;
            uint32_t c = t;
            for (uint32_t l = link(c); l; l = link(c)) c -= l;
            const bool followed = (t + 1u < n && link(t + 1u) == 1u) || (t + 2u < n && link(t + 2u) == 2u);
            sm.cls[t] = (uint8_t)(c | ((c == t && !followed) ? 0x80u : 0u));
        }
        __syncthreads();

        const int g_per = (TILE_H / 8) / (int)split;
        for (int g = (int)sub * g_per; g < ((int)sub + 1) * g_per; g++) {
            const int Y = g * 8 + warp;
            if (Y >= vh) break;
            const size_t rowp = (size_t)(py0 + Y) * fb.width + px0;
            bool inb[FILL_PX];
            PixelState S;
            const bool first = (w0 == 0u);
#pragma unroll
            for (int p = 0; p < FILL_PX; p++) {
                inb[p] = X[p] < vw;
                const bool in_clr = first && clr_any && X[p] >= cx0 && X[p] < cx1 && Y >= cy0 && Y < cy1;
                uint32_t c = 0;
                S.depth[p] = 0.0f; S.stencil[p] = 0u;
                if (inb[p]) {
                    if (planes & 1u) c = (in_clr && (clr.mask & G_COLOR_BUFFER_BIT)) ? clr.color : fb.color[rowp + X[p]];
                    if (planes & 2u) S.depth[p] = (in_clr && (clr.mask & G_DEPTH_BUFFER_BIT)) ? clr.depth : fb.depth[rowp + X[p]];
                    if (planes & 4u) S.stencil[p] = (in_clr && (clr.mask & G_STENCIL_BUFFER_BIT)) ? (clr.stencil & 0xFFu) : (uint32_t)fb.stencil[rowp + X[p]];
                }
                S.r[p] = (float)(c & 0xFFu); S.g[p] = (float)((c >> 8) & 0xFFu); S.b[p] = (float)((c >> 16) & 0xFFu); S.a[p] = (float)(c >> 24);
            }
            /* The intermediates H belong to a geometry class; a triangle of the class H holds skips straight to its own
             * stages, a triangle of a class that covered none of this thread's pixels is skipped altogether.  Stacked
             * quads (C3) give runs of one class where a tile lies inside one half of the quad, and two alternating
             * classes in the tiles on the diagonal -- there a thread's pixels belong to one half (the other half misses
             * and leaves H alone), except for the pixels ON the diagonal, which recompute. */
            Shared H;
            uint32_t held = 0xFFFFFFFFu, missed = 0xFFFFFFFFu;
            constexpr int UNROLL = ((ON & FILL_PSEUDO) == FILL_PSEUDO) ? 2 : 1;     /* the small instance: two triangles in flight */
#pragma unroll UNROLL
            for (uint32_t t = 0; t < n; t++) {
                const uint32_t ce = sm.cls[t], c = ce & 63u;
                if (c == missed) continue;
                if (c != held) {
                    if (!fill_shared<ON, OFF>(b, sm, sm.tri[t], px0, py0, Y, X, inb, (ce & 0x80u) != 0u, H)) { missed = c; continue; }
                    held = c;
                }
                fill_one<ON, OFF>(b, sm.tri[t], H, S);
            }
#pragma unroll
            for (int p = 0; p < FILL_PX; p++) {
                if (!inb[p]) continue;
                if (planes & 1u) {
                    const uint32_t c = __float2uint_rz(S.r[p]) | (__float2uint_rz(S.g[p]) << 8) | (__float2uint_rz(S.b[p]) << 16) | (__float2uint_rz(S.a[p]) << 24);
                    fb.color[rowp + X[p]] = c;
                    if (fb.present && w0 + FILL_WINDOW >= L) fb.present[rowp + X[p]] = c;       /* fused gather: the finished row goes to the presenting GPU */
                }
                if (planes & 2u) fb.depth[rowp + X[p]] = S.depth[p];
                if (planes & 4u) fb.stencil[rowp + X[p]] = (uint8_t)S.stencil[p];
            }
        }
    }
}

/* The fill-rate state mix (C3): textured, no depth test, no fog, smooth shading -- the dead tests and their registers
 * leave the kernel.  Everything else runs the fully dynamic instance. */
constexpr uint32_t FILL_FAST_ON = RC_TEXTURED;
constexpr uint32_t FILL_FAST_OFF = RC_DEPTH_TEST | RC_FOG | RC_FLAT;
/* ... and the mix itself: textured, alpha-tested (GREATER), stencil (ALWAYS), blended (SRC_ALPHA, ONE_MINUS_SRC_ALPHA),
 * MODULATE, full colour mask */
constexpr uint32_t FILL_FASTER_ON = RC_TEXTURED | RC_ALPHA_TEST | RC_STENCIL | RC_BLEND | FILL_PSEUDO;

void launch_fill(const BatchDev &b, const FrameTargets &fb, const ClearOp &clear, uint32_t planes, uint32_t fill_mode, uint32_t all_on, uint32_t any_on,
                 cudaStream_t s)
{
    static bool configured[64] = { false };
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaFuncSetAttribute(k_fill<0u, 0u>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FillSmem));
        cudaFuncSetAttribute(k_fill<FILL_FAST_ON, FILL_FAST_OFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FillSmem));
        cudaFuncSetAttribute(k_fill<FILL_FASTER_ON, FILL_FAST_OFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FillSmem));
        configured[dev] = true;
    }
    const uint32_t tiles = (uint32_t)(fb.tiles_x * fb.tile_rows);
    if (tiles == 0 || fill_mode == FILL_OFF) return;
    /* (1, 2, 4 or 8: the tile is walked in eight passes of eight rows.)  Tiles differ in cost by up to 4x -- layers whose
     * alpha test passes, lists that hold both halves of stacked quads -- so on the band of a multi-GPU frame a few
     * whole-tile CTAs outlast everything else: up to FILL_SPLIT_TILES tiles they are cut in four.  Measured on C3
     * (emulated bands): 1020 tiles 0.82 -> 0.71 ms, 540 tiles 0.79 -> 0.45; the full frame (2040 tiles) 1.18 -> 1.34, hence
     * the bound. */
    constexpr uint32_t FILL_SPLIT_TILES = 1100u;
    const uint32_t split = small_grid(tiles, FILL_SPLIT_TILES) ? 4u : 1u;
    /* all_on / any_on: AND / OR of the RasterCfg flags of the pass's in-order states */
    if ((all_on & FILL_FASTER_ON) == FILL_FASTER_ON && (any_on & FILL_FAST_OFF) == 0u)
        k_fill<FILL_FASTER_ON, FILL_FAST_OFF><<<tiles * split, FILL_THREADS, sizeof(FillSmem), s>>>(b, fb, clear, planes, fill_mode, split);
    else if ((all_on & FILL_FAST_ON) == FILL_FAST_ON && (any_on & FILL_FAST_OFF) == 0u)
        k_fill<FILL_FAST_ON, FILL_FAST_OFF><<<tiles * split, FILL_THREADS, sizeof(FillSmem), s>>>(b, fb, clear, planes, fill_mode, split);
    else
        k_fill<0u, 0u><<<tiles * split, FILL_THREADS, sizeof(FillSmem), s>>>(b, fb, clear, planes, fill_mode, split);
    note_launch();
}

} // namespace mtgl_dev_impl
