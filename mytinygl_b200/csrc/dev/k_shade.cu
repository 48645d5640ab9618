/*
 * k_shade.cu -- K4b, the shade pass: one colour per VISIBLE pixel of the tiles resolved by the visibility kernels
 * (k_vis.cu, k_raster<true>).
 *
 * Replaces interpolation, texture_sample_lod, texenv, fog and the colour write of rasterize_triangle_smooth
 * (src/raster.c:581-705, 719-721; src/textures.c:457-557) for fragments of deferrable states (no blending, no alpha
 * test, full colour mask): only the last surviving fragment of a pixel has to be shaded, and K4a has already found it.
 *
 * One CTA per tile (or per 16-row quarter of a tile, SPLIT = 4, for grids below one wave):
 *   pass 1  the tile's visibility entries -> a colour tile in shared memory: the record index for visible pixels, the
 *           clear colour / the plane's colour for the others; visible pixels are compacted into a row-major list (one
 *           block scan), so that pass 2 runs with every lane busy and neighbouring lanes on neighbouring pixels;
 *   texels  small textures have a float4 copy in HBM (n / 255 per channel, built at glTexImage2D time; dev_fasttex.cuh):
 *           a bilinear tap is one 16-byte load through L1, the trilinear sample of a C4 fragment 8 of them, instead of 8
 *           4-byte loads and 32 shift / mask / table look-ups;
 *   pass 2  P pixels per thread and iteration (shipped: one, at 64 registers and 4 CTAs per SM; two independent streams
 *           per thread at 2-3 CTAs per SM measured slower): record -> barycentrics with the
 *           reference's expressions (raster.c:534-544) -> colour, texel (sampler plan from the record's LOD), texenv, fog
 *           -> packed colour into the tile.  Fragments that cannot take the float4 path -- per-fragment lighting, a large
 *           texture, REPEAT with a size that is no power of two, non-finite attributes -- go through the general code
 *           of dev_shade.cuh, out of line;
 *   pass 3  the tile leaves as whole rows in 16-byte stores: to the colour plane and, for a band of a multi-GPU frame,
 *           straight into the presenting GPU's plane over NVLink (fused gather).
 *
 * Algorithmic bytes: 4 B colour write per depth-passing fragment + 4 B per cleared pixel; real traffic:
 * 4 B visibility in, 4 B colour out per pixel, 160 B record per visible pixel through L1 / L2.
 */
#include "dev_common.cuh"
#include "dev_shade.cuh"
#include "dev_fasttex.cuh"
#include "dev_fill.cuh"


namespace mtgl_dev_impl {

void note_launch();

constexpr uint32_t VIS_NONE = 0xFFFFFFFFu;
constexpr int SHADE_THREADS = 256;

template <int SPLIT>
struct ShadeSmem {
    uint32_t tile[TILE_W * (TILE_H / SPLIT)];       /* record index of a visible pixel until it is shaded, packed colour otherwise */
    uint16_t list[TILE_W * (TILE_H / SPLIT)];
    float un[256];
    uint32_t warp_total[SHADE_THREADS / 32];
    RasterCfg cfg0;                                 /* the batch's only configuration block, if it has just one */
};

/* the general path (dev_shade.cuh), out of line so that its registers do not count against the fast path */
__device__ __noinline__ uint32_t shade_general(const BatchDev &b, const float *un, uint32_t r, float b0, float b1, float b2)
{
    const TriRecord *rec = b.records + r;
    const uint32_t state_flags = __ldg(&rec->state_flags);
    TriAttr A;
    load_attr(A, rec);
    Color4 c;
    shade_color(b, un, r, state_flags, A, b.cfgs + (state_flags & STATE_INDEX_MASK), b0, b1, b2, c);
    return color_pack(c);      /* raster.c:719-721: color_pack clamps */
}

/* P: pixels per thread and iteration of pass 2; CTAS: CTAs per SM the register budget is set for.
 * ON / OFF: RasterCfg flags (and dev_fill.cuh pseudo-flags) that every state of the pass has / that none has -- decided at
 * compile time instead of per pixel (the instance for the plain textured mix: smooth shading, GL_MODULATE, no fog, no
 * per-fragment lighting; everything else runs <0, 0>). */
template <int SPLIT, int P, int CTAS, uint32_t ON, uint32_t OFF>
__global__ void __launch_bounds__(SHADE_THREADS, CTAS) k_shade(BatchDev b, FrameTargets fb, ClearOp clr)
{
    constexpr int ROWS = TILE_H / SPLIT;        /* rows of the tile this CTA owns */
    constexpr int PX = 16 / SPLIT;              /* consecutive pixels per thread in pass 1 */
    constexpr int TPR = TILE_W / PX;            /* threads per row */
    __shared__ __align__(16) ShadeSmem<SPLIT> sm;
    if (!lists_fit(b)) return;

    const uint32_t slot = blockIdx.x / SPLIT, sub = blockIdx.x % SPLIT;
    const uint32_t tile = b.tile_order ? b.tile_order[slot] : slot;
    const int tx = (int)(tile % (uint32_t)fb.tiles_x), ty = (int)(tile / (uint32_t)fb.tiles_x) + fb.tile_y0;
    const int row0 = (ty << TILE_LOG) + (int)sub * ROWS;
    const int px0 = tx << TILE_LOG, py0 = max(row0, fb.band_y0);
    const int vw = min(TILE_W, fb.width - px0), vh = min(row0 + ROWS, fb.band_y1) - py0;
    if (vw <= 0 || vh <= 0) return;
    const uint32_t L = b.tile_count ? b.tile_count[tile] : 0u;
    if (L && (b.tile_flags[tile] & 1u)) return;             /* an in-order kernel owns this tile */
    const bool clr_here = clr.mask && clr.x0 < px0 + vw && clr.x1 > px0 && clr.y0 < py0 + vh && clr.y1 > py0;   /* as in k_raster */
    if (L == 0 && !clr_here) return;
    sm.un[threadIdx.x] = b.unorm8[threadIdx.x];
    /* a batch with one state block (C4, C5): its configuration is read from shared memory, so a pixel's chain of dependent
     * loads is list -> record and not list -> record -> configuration */
    const bool one_cfg = b.n_states == 1u;
    if (one_cfg && threadIdx.x < sizeof(RasterCfg) / 4u)
        reinterpret_cast<uint32_t *>(&sm.cfg0)[threadIdx.x] = __ldg(reinterpret_cast<const uint32_t *>(b.cfgs) + threadIdx.x);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool vec = (vw == TILE_W) && ((fb.width & 3) == 0);

    /* ---- pass 1: visibility entries + clear / existing colour -> the tile; compaction of the pixels to shade.  A
     * thread owns PX consecutive pixels of one row, so its loads are 16-byte loads issued back to back and thread order
     * is row-major pixel order: the compacted list keeps neighbouring pixels next to each other for pass 2. ---- */
    const int y = (int)threadIdx.x / TPR, xq = ((int)threadIdx.x % TPR) * PX;
    uint32_t has_mask = 0;
    uint32_t v[PX];
    if (y < vh) {
        const size_t p0 = (size_t)(py0 + y) * fb.width + px0 + xq;
        if (!L) {
#pragma unroll
            for (int k = 0; k < PX; k++) v[k] = VIS_NONE;
        } else if (vec) {
#pragma unroll
            for (int q = 0; q < PX / 4; q++) {
                const uint4 t = *reinterpret_cast<const uint4 *>(b.vis_plane + p0 + q * 4);
                v[q * 4 + 0] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < PX; k++) v[k] = (xq + k < vw) ? b.vis_plane[p0 + k] : VIS_NONE;
        }
        const bool clr_row = (clr.mask & G_COLOR_BUFFER_BIT) && py0 + y >= clr.y0 && py0 + y < clr.y1;
#pragma unroll
        for (int k = 0; k < PX; k++) {
            uint32_t word = v[k];
            if (v[k] != VIS_NONE) {
                has_mask |= 1u << k;
                /* pass 2 starts two barriers from here: its records (160 B, two 128-byte lines at most) travel L2 -> L1
                 * meanwhile; a run of pixels of one triangle asks once */
                if (k == 0 || v[k] != v[k > 0 ? k - 1 : 0]) {
                    const char *rp = reinterpret_cast<const char *>(b.records + v[k]);
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(rp));
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(rp + 128));
                }
            } else if (xq + k < vw) word = (clr_row && px0 + xq + k >= clr.x0 && px0 + xq + k < clr.x1) ? clr.color : fb.color[p0 + k];
            sm.tile[y * TILE_W + xq + k] = word;
        }
    }
    const uint32_t mine = (uint32_t)__popc(has_mask);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += up;
    }
    if (lane == 31) sm.warp_total[warp] = incl;
    __syncthreads();
    uint32_t at = incl - mine, n = 0;
#pragma unroll
    for (uint32_t w = 0; w < SHADE_THREADS / 32; w++) {
        const uint32_t wt = sm.warp_total[w];
        if (w < warp) at += wt;
        n += wt;
    }
#pragma unroll
    for (int k = 0; k < PX; k++)
        if (has_mask & (1u << k)) sm.list[at++] = (uint16_t)(y * TILE_W + xq + k);

    __syncthreads();

    /* ---- pass 2: shade the compacted pixels, P per thread and iteration ---- */
    for (uint32_t i0 = threadIdx.x; i0 < n; i0 += P * SHADE_THREADS) {
        int ci[P];
        uint32_t r[P], out[P];
        bool on[P], general[P];
        float b0[P], b1[P], b2[P];
        uint32_t sflags[P];
#pragma unroll
        for (int p = 0; p < P; p++) {
            const uint32_t i = i0 + (uint32_t)p * SHADE_THREADS;
            on[p] = i < n;
            ci[p] = sm.list[on[p] ? i : i0];
            r[p] = sm.tile[ci[p]];
            const TriRecord *rec = b.records + r[p];
            const int4 row0 = __ldg(reinterpret_cast<const int4 *>(rec) + 0);
            const int4 row1 = __ldg(reinterpret_cast<const int4 *>(rec) + 1);
            sflags[p] = __ldg(&rec->state_flags);
            const float fx0 = (float)row0.x, fy0 = (float)row0.y, fx1 = (float)row0.z, fy1 = (float)row0.w;
            const float fx2 = (float)row1.x, fy2 = (float)row1.y;
            const float inv_area = __int_as_float(row1.w);
            const float px = (float)(px0 + ci[p] % TILE_W), py = (float)(py0 + ci[p] / TILE_W);
            b0[p] = edge_at(fx1, fy1, fx2, fy2, px, py) * inv_area;
            b1[p] = edge_at(fx2, fy2, fx0, fy0, px, py) * inv_area;
            b2[p] = edge_at(fx0, fy0, fx1, fy1, px, py) * inv_area;
        }
#pragma unroll
        for (int p = 0; p < P; p++) {
            const TriRecord *rec = b.records + r[p];
            const RasterCfg *cfg = one_cfg ? &sm.cfg0 : b.cfgs + (sflags[p] & STATE_INDEX_MASK);
            const uint32_t flags = cfg->flags;
            auto has = [&](uint32_t bit) -> bool { return (ON & bit) ? true : ((OFF & bit) ? false : (flags & bit) != 0u); };
            TriAttr A;
            load_attr(A, rec);
            /* per-fragment lighting (raster.c:592-615) and textures without a usable float4 copy take the general path */
            general[p] = has(RC_LIGHTING) && (has(RC_PHONG) || ((sflags[p] >> 31) && has(RC_TWO_SIDE)));
            if (has(RC_TEXTURED)) general[p] = general[p] || !cfg->fast_tex || !(sflags[p] & STATE_BOUNDED_BIT);
            Color4 c;
            if (has(RC_FLAT)) c = { A.col2.x, A.col2.y, A.col2.z, A.col2.w };        /* third vertex of the sub-triangle (raster.c:583-585) */
            else {
                c.r = A.col0.x * b0[p] + A.col1.x * b1[p] + A.col2.x * b2[p];
                c.g = A.col0.y * b0[p] + A.col1.y * b1[p] + A.col2.y * b2[p];
                c.b = A.col0.z * b0[p] + A.col1.z * b1[p] + A.col2.z * b2[p];
                c.a = A.col0.w * b0[p] + A.col1.w * b1[p] + A.col2.w * b2[p];
            }
            if (has(RC_TEXTURED) && !general[p]) {          /* raster.c:618-669 */
                float u, v2;
                if (has(RC_PERSPECTIVE)) {
                    const float u0w = A.u0 * A.w0, v0w = A.v0 * A.w0, u1w = A.u1 * A.w1, v1w = A.v1 * A.w1, u2w = A.u2 * A.w2, v2w = A.v2 * A.w2;
                    const float uw = b0[p] * u0w + b1[p] * u1w + b2[p] * u2w;
                    const float vw_ = b0[p] * v0w + b1[p] * v1w + b2[p] * v2w;
                    const float ow = b0[p] * A.w0 + b1[p] * A.w1 + b2[p] * A.w2;
                    const float w = 1.0f / ow;
                    u = uw * w; v2 = vw_ * w;
                } else {
                    u = b0[p] * A.u0 + b1[p] * A.u1 + b2[p] * A.u2;
                    v2 = b0[p] * A.v0 + b1[p] * A.v1 + b2[p] * A.v2;
                }
                float cl;
                const uint32_t plan = sampler_plan(cfg, A.lod, cl);
                const TexView tv = { cfg->tex_f4, cfg->tex_w, cfg->tex_h, cfg->tex_w1, cfg->tex_h1, cfg->tex_w * cfg->tex_h };
                const float4 t = fast_sample<true>(tv, plan, cl, u, v2);
                switch ((ON & FILL_MODULATE) ? (uint32_t)G_MODULATE : cfg->tex_env_mode) {
                case G_REPLACE: c = { t.x, t.y, t.z, t.w }; break;
                case G_DECAL: c = color_lerp_rgb(c, { t.x, t.y, t.z, t.w }, t.w); break;
                case G_BLEND: {
                    const float *e = cfg->tex_env_color;
                    c = { c.r * (1.0f - t.x) + e[0] * t.x, c.g * (1.0f - t.y) + e[1] * t.y, c.b * (1.0f - t.z) + e[2] * t.z, c.a * t.w };
                    break;
                }
                case G_ADD: c = { c.r + t.x, c.g + t.y, c.b + t.z, c.a * t.w }; break;
                default: c = { c.r * t.x, c.g * t.y, c.b * t.z, c.a * t.w }; break;
                }
            }
            if (has(RC_FOG)) {                                   /* raster.c:672-705; result alpha = fog colour alpha */
                const float fc = b0[p] * A.ez0 + b1[p] * A.ez1 + b2[p] * A.ez2;
                const Color4 fogc = { cfg->fog_color[0], cfg->fog_color[1], cfg->fog_color[2], cfg->fog_color[3] };
                c = color_lerp_rgb(fogc, c, fog_factor(cfg, fc));
            }
            out[p] = color_pack(c);        /* raster.c:719-721: color_pack clamps */
        }
#pragma unroll
        for (int p = 0; p < P; p++) {
            if (!on[p]) continue;
            if (general[p]) out[p] = shade_general(b, sm.un, r[p], b0[p], b1[p], b2[p]);
            sm.tile[ci[p]] = out[p];
        }
    }
    __syncthreads();

    /* ---- pass 3: the finished rows leave in 16-byte stores (north_star 5); the band of a multi-GPU frame also goes to
     * the presenting GPU's plane over NVLink (fused gather) ---- */
    if (vec) {
        for (int i = threadIdx.x; i < vh * (TILE_W / 4); i += SHADE_THREADS) {
            const int yy = i >> 4, q = (i & 15) * 4;
            const uint4 c = *reinterpret_cast<const uint4 *>(&sm.tile[yy * TILE_W + q]);
            const size_t p = (size_t)(py0 + yy) * fb.width + px0 + q;
            *reinterpret_cast<uint4 *>(fb.color + p) = c;
            if (fb.present) *reinterpret_cast<uint4 *>(fb.present + p) = c;
        }
    } else {
        for (int i = threadIdx.x; i < vh * TILE_W; i += SHADE_THREADS) {
            const int yy = i >> 6, x = i & 63;
            if (x >= vw) continue;
            const size_t p = (size_t)(py0 + yy) * fb.width + px0 + x;
            fb.color[p] = sm.tile[yy * TILE_W + x];
            if (fb.present) fb.present[p] = sm.tile[yy * TILE_W + x];
        }
    }
}

/* the plain textured mix (C4 / C5): every state textured with GL_MODULATE, smooth-shaded, no fog, no per-fragment lighting */
constexpr uint32_t SHADE_PLAIN_ON = RC_TEXTURED | FILL_MODULATE;
constexpr uint32_t SHADE_PLAIN_OFF = RC_FLAT | RC_FOG | RC_PHONG | RC_TWO_SIDE;

/* all_on / any_on: AND / OR of the RasterCfg flags (+ pseudo-flags, dev_fill.cuh) of the pass's states.
 * One pixel per thread at 64 registers and 4 CTAs per SM; measured slower on C4: two pixels at 2 CTAs (0.168 ms against
 * 0.148), one pixel at 3 CTAs (0.153), two pixels at 3 CTAs with a few spilled registers (0.157). */
void launch_shade(const BatchDev &b, const FrameTargets &fb, const ClearOp &clear, uint32_t all_on, uint32_t any_on, cudaStream_t s)
{
    const uint32_t tiles = (uint32_t)(fb.tiles_x * fb.tile_rows);
    if (tiles == 0) return;
    const bool plain = (all_on & SHADE_PLAIN_ON) == SHADE_PLAIN_ON && (any_on & SHADE_PLAIN_OFF) == 0u;
    if (small_grid(tiles)) {
        if (plain) k_shade<4, 1, 4, SHADE_PLAIN_ON, SHADE_PLAIN_OFF><<<tiles * 4u, SHADE_THREADS, 0, s>>>(b, fb, clear);
        else k_shade<4, 1, 4, 0u, 0u><<<tiles * 4u, SHADE_THREADS, 0, s>>>(b, fb, clear);
    } else {
        if (plain) k_shade<1, 1, 4, SHADE_PLAIN_ON, SHADE_PLAIN_OFF><<<tiles, SHADE_THREADS, 0, s>>>(b, fb, clear);
        else k_shade<1, 1, 4, 0u, 0u><<<tiles, SHADE_THREADS, 0, s>>>(b, fb, clear);
    }
    note_launch();
}

} // namespace mtgl_dev_impl
