/*
 * k_raster.cu -- K4: fine rasterisation, texture sampling and per-fragment operations on
 * tile-resident colour / depth / stencil, plus the fused glClear and the tile write-back (K5).
 *
 * Replaces the scan loop and fragment block of rasterize_triangle_smooth (src/raster.c:532-724),
 * edge_function (299-302), depth_test / alpha_test / stencil_test / stencil_op /
 * write_stencil_masked (344-448), get_blend_factor / blend_colors (360-387), write_pixel_masked
 * (20-45), texture_sample_lod and helpers (src/textures.c:272-557), the per-fragment
 * compute_lighting calls (raster.c:593-614) and glClear (src/gl_api.c:409-457).
 *
 * One CTA per 64x64 tile.  The three planes of the tile live in shared memory for the whole
 * batch: loaded once with 128-bit loads (or initialised from the clear values when the batch
 * starts with a clear that covers the tile -- a cleared tile is never read from HBM), updated in
 * place by every fragment, written back once with 128-bit stores.  The tile's triangle references
 * are sorted by submission id first.  The tile is cut into 16 regions of 16x16 pixels; each of the 8
 * warps repeatedly takes the next region, walks the sorted list 32 references at a time (one
 * bounding-box test per lane, warp ballot), and rasterises the hits in order over 8x4-pixel
 * blocks -- so every pixel sees its fragments in exactly the reference's order (blending, stencil
 * counting, depth ties, double-shaded shared edges) without per-pixel atomics, and the regions
 * balance the load between warps.  Fragments of states that neither blend nor alpha-test only
 * update depth/stencil and a per-pixel visibility entry; their colour is computed once per pixel,
 * with full warps, when the region is resolved (deferred shading).
 *
 * Shared-memory rows are padded (72 words / 80 bytes) so that the 32 lanes of an 8x4 block hit
 * 32 distinct banks.
 *
 * Roofline: algorithmically HBM-bound (2-18 B of framebuffer traffic per covered
 * fragment in the reference's immediate-mode formulation); in this tile-resident formulation the
 * real DRAM traffic is one load + one store per touched tile and the kernel is issue-bound.
 */
#include "dev_common.cuh"
#include "dev_texture.cuh"
#include "dev_fragment.cuh"
#include "dev_fill.cuh"
#include "dev_shade.cuh"

namespace mtgl_dev_impl {

void note_launch();

constexpr uint32_t VIS_NONE = 0xFFFFFFFFu;

struct RasterSmem {
    float depth[TILE_H * COLOR_PITCH];
    /* visibility buffer: record of the last fragment that passed the stencil/depth tests and whose colour has
     * not been computed yet (deferred shading), VIS_NONE otherwise */
    uint32_t vis[TILE_H * COLOR_PITCH];
    uint8_t stencil[TILE_H * STENCIL_PITCH];
    union {
        uint32_t key[LIST_WINDOW];      /* sort keys: dead once the window is sorted ... */
        uint16_t pend[RASTER_THREADS / 32][REGION_W * REGION_H];   /* ... reused for the per-warp lists of pixels to shade */
    };
    uint32_t rec[LIST_WINDOW];
    uint32_t box[LIST_WINDOW];
    uint16_t pos[LIST_WINDOW];          /* sort payload: original position of the key */
    float unorm8[256];
    uint32_t count;
    uint32_t next_region;       /* dynamic region scheduler of the current window */
    uint32_t region_work[NUM_REGIONS];      /* estimated work per region (8x4 blocks touched + visits) */
    uint32_t region_order[NUM_REGIONS];     /* regions sorted by decreasing work: heaviest first */
    uint32_t scratch[RASTER_THREADS / 32];
    /* last: the visibility-only variant of the kernel (K4a) does not allocate the colour plane */
    __align__(16) uint32_t color[TILE_H * COLOR_PITCH];
};
static_assert(sizeof(uint16_t) * (RASTER_THREADS / 32) * REGION_W * REGION_H <= sizeof(uint32_t) * LIST_WINDOW, "pending lists must fit in the key array");
static_assert((LIST_WINDOW & (LIST_WINDOW - 1)) == 0, "the bitonic sort pads to a power of two");

/* Deferred shading of one 16x16 region: every pixel whose visibility entry is set gets the colour of that
 * (last, in submission order) fragment.  Only fragments of states that neither blend nor alpha-test nor mask
 * colour channels are deferred, so earlier fragments of the pixel would have been overwritten anyway: the work
 * is done once per covered pixel instead of once per passing fragment, and with full warps -- the pixels to
 * shade are first compacted into a per-warp list.  Barycentrics are recomputed from the record with the
 * reference's expressions (raster.c:534-544), which is deterministic. */
__device__ __noinline__ void resolve_region(const BatchDev &b, RasterSmem &sm, int rx0, int ry0, int tile_px, int tile_py)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t count = 0;
#pragma unroll 1
    for (int blk = 0; blk < (REGION_W / 8) * (REGION_H / 4); blk++) {
        const int x = rx0 + (blk % (REGION_W / 8)) * 8 + (int)(lane & 7), y = ry0 + (blk / (REGION_W / 8)) * 4 + (int)(lane >> 3);
        const int ci = y * COLOR_PITCH + x;
        const bool has = sm.vis[ci] != VIS_NONE;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, has);
        if (has) sm.pend[warp][count + __popc(m & lt_mask)] = (uint16_t)ci;
        count += __popc(m);
    }
    __syncwarp();
#pragma unroll 1
    for (uint32_t base = 0; base < count; base += 32) {
        if (base + lane < count) {
            const int ci = sm.pend[warp][base + lane];
            const uint32_t r = sm.vis[ci];
            sm.vis[ci] = VIS_NONE;
            const TriRecord *rec = b.records + r;
            const int4 row0 = __ldg(reinterpret_cast<const int4 *>(rec) + 0);
            const int4 row1 = __ldg(reinterpret_cast<const int4 *>(rec) + 1);
            const uint32_t state_flags = __ldg(&rec->state_flags);
            const float fx0 = (float)row0.x, fy0 = (float)row0.y, fx1 = (float)row0.z, fy1 = (float)row0.w;
            const float fx2 = (float)row1.x, fy2 = (float)row1.y;
            const float inv_area = __int_as_float(row1.w);
            const float px = (float)(tile_px + ci % COLOR_PITCH), py = (float)(tile_py + ci / COLOR_PITCH);
            const float b0 = edge_at(fx1, fy1, fx2, fy2, px, py) * inv_area;
            const float b1 = edge_at(fx2, fy2, fx0, fy0, px, py) * inv_area;
            const float b2 = edge_at(fx0, fy0, fx1, fy1, px, py) * inv_area;
            const RasterCfg *cfg = b.cfgs + (state_flags & STATE_INDEX_MASK);
            TriAttr A;
            load_attr(A, rec);
            Color4 c;
            shade_color(b, sm.unorm8, r, state_flags, A, cfg, b0, b1, b2, c);
            sm.color[ci] = color_pack(c);      /* raster.c:719-721: color_pack clamps */
        }
    }
    __syncwarp();
}

/* Immediate (in-order) shading of one fragment for states that blend, alpha-test or mask channels:
 * colour, then the late depth write (raster.c:707-710), blending and the masked write (712-721). */
__device__ __noinline__ void shade_now(const BatchDev &b, RasterSmem &sm, uint32_t r, uint32_t state_flags, const RasterCfg *cfg,
                                       float b0, float b1, float b2, int ci, float depth, bool depth_write)
{
    const float *un = sm.unorm8;
    const uint32_t flags = cfg->flags, cm = cfg->color_mask;
    TriAttr A;
    load_attr(A, b.records + r);
    Color4 c;
    if (!shade_color(b, un, r, state_flags, A, cfg, b0, b1, b2, c)) return;
    if (depth_write) sm.depth[ci] = depth;
    if (flags & RC_BLEND) {                     /* raster.c:712-717 */
        Color4 d = color_unpack(sm.color[ci], un);
        Color4 sf = blend_factor(cfg->blend_src, c, d), df = blend_factor(cfg->blend_dst, c, d);
        c = color_clamp({ c.r * sf.r + d.r * df.r, c.g * sf.g + d.g * df.g, c.b * sf.b + d.b * df.b, c.a * sf.a + d.a * df.a });
    }
    /* (the clamp of raster.c:719 is part of color_pack) */
    if (cm == 0xFu) sm.color[ci] = color_pack(c);       /* write_pixel_masked, raster.c:20-45 */
    else if (cm != 0u) {
        Color4 d = color_unpack(sm.color[ci], un);
        if (cm & 1u) d.r = c.r;
        if (cm & 2u) d.g = c.g;
        if (cm & 4u) d.b = c.b;
        if (cm & 8u) d.a = c.a;
        sm.color[ci] = color_pack(d);
    }
}

/* ---------------------------------------------------------------- points and lines (general kernel only) */
/* depth test, blend, depth write, masked colour write of one line / point pixel: write_line_pixel (raster.c:66-104)
 * and the pixel loops of flush_points (1123-1161) / draw_point_at_screen (795-843).  These paths have no stencil. */
__device__ __forceinline__ void simple_pixel(RasterSmem &sm, const RasterCfg *cfg, int ci, float depth, Color4 c)
{
    const float *un = sm.unorm8;
    const uint32_t flags = cfg->flags, cm = cfg->color_mask;
    if ((flags & RC_DEPTH_TEST) && !compare_f(cfg->depth_func, depth, sm.depth[ci])) return;
    if (flags & RC_BLEND) {
        Color4 d = color_unpack(sm.color[ci], un);
        Color4 sf = blend_factor(cfg->blend_src, c, d), df = blend_factor(cfg->blend_dst, c, d);
        c = color_clamp({ c.r * sf.r + d.r * df.r, c.g * sf.g + d.g * df.g, c.b * sf.b + d.b * df.b, c.a * sf.a + d.a * df.a });
    }
    if ((flags & (RC_DEPTH_TEST | RC_DEPTH_WRITE)) == (RC_DEPTH_TEST | RC_DEPTH_WRITE)) sm.depth[ci] = depth;
    if (cm == 0xFu) sm.color[ci] = color_pack(c);
    else if (cm != 0u) {
        Color4 d = color_unpack(sm.color[ci], un);
        if (cm & 1u) d.r = c.r;
        if (cm & 2u) d.g = c.g;
        if (cm & 4u) d.b = c.b;
        if (cm & 8u) d.a = c.a;
        sm.color[ci] = color_pack(d);
    }
}

/* a POINT record over the part [X0,X1]x[Y0,Y1] (tile-relative, already clamped to the record's box) of a region */
__device__ __noinline__ void raster_point(const BatchDev &b, RasterSmem &sm, uint32_t r, int X0, int Y0, int X1, int Y1)
{
    const uint32_t lane = threadIdx.x & 31;
    const TriRecord *rec = b.records + r;
    const uint32_t state_flags = __ldg(&rec->state_flags);
    const RasterCfg *cfg = b.cfgs + (state_flags & STATE_INDEX_MASK);
    const float depth = __ldg(&rec->z0);
    const float4 col = __ldg(reinterpret_cast<const float4 *>(rec) + 5);
    for (int by = Y0; by <= Y1; by += 4)
        for (int bx = X0; bx <= X1; bx += 8) {
            const int x = bx + (int)(lane & 7), y = by + (int)(lane >> 3);
            if (x <= X1 && y <= Y1) simple_pixel(sm, cfg, y * COLOR_PITCH + x, depth, { col.x, col.y, col.z, col.w });
        }
}

/* A LINE record: the pixels of the reference's Bresenham walk (raster.c:154-240) in closed form.  Step k of the walk
 * (k = 0 .. max(|dx|,|dy|)) sits at major = start + k * sign and minor = start + sign * floor((2 k a_minor + a_major - 1)
 * / (2 a_major)) (verified exhaustively against the loop); the width replicates perpendicular to the major axis, so
 * all pixels of one line are distinct and the steps can be processed in parallel, 32 per warp iteration. */
__device__ __noinline__ void raster_line(const BatchDev &b, RasterSmem &sm, uint32_t r, int tile_px, int tile_py,
                                         int X0, int Y0, int X1, int Y1)
{
    const uint32_t lane = threadIdx.x & 31;
    const float *un = sm.unorm8;
    const TriRecord *rec = b.records + r;
    const int4 row0 = __ldg(reinterpret_cast<const int4 *>(rec) + 0);
    const int lw = __ldg(&rec->x2);
    const uint32_t state_flags = __ldg(&rec->state_flags);
    const RasterCfg *cfg = b.cfgs + (state_flags & STATE_INDEX_MASK);
    const uint32_t flags = cfg->flags;
    const int x0 = row0.x, y0 = row0.y, x1 = row0.z, y1 = row0.w;
    const long long dx = (long long)x1 - x0, dy = (long long)y1 - y0;
    const long long adx = dx < 0 ? -dx : dx, ady = dy < 0 ? -dy : dy;
    const int sx = dx < 0 ? -1 : 1, sy = dy < 0 ? -1 : 1;
    const long long total = adx > ady ? adx : ady;
    const float total_steps = (float)(total == 0 ? 1 : (int)total);
    const bool x_major = adx > ady;                 /* also the replication axis: vertical when mostly horizontal */
    const int half = lw / 2;
    /* steps whose major coordinate falls into the rectangle (absolute pixel coordinates) */
    const int lo = x_major ? tile_px + X0 : tile_py + Y0, hi = x_major ? tile_px + X1 : tile_py + Y1;
    const int start = x_major ? x0 : y0, sgn = x_major ? sx : sy;
    long long k0, k1;                               /* inclusive step range with lo <= start + sgn * k <= hi */
    if (sgn > 0) { k0 = (long long)lo - start; k1 = (long long)hi - start; } else { k0 = (long long)start - hi; k1 = (long long)start - lo; }
    if (k0 < 0) k0 = 0;
    if (k1 > total) k1 = total;
    const float z0 = __ldg(&rec->z0), z1 = __ldg(&rec->z1);
    const float4 c0 = __ldg(reinterpret_cast<const float4 *>(rec) + 5), c1 = __ldg(reinterpret_cast<const float4 *>(rec) + 6);
    const float4 uv = __ldg(reinterpret_cast<const float4 *>(rec) + 8);
    const float ez0 = __ldg(&rec->ez0), ez1 = __ldg(&rec->ez1);

    for (long long kb = k0; kb <= k1; kb += 32) {
        const long long k = kb + lane;
        if (k > k1) continue;
        int px, py;
        if (adx >= ady) {
            px = x0 + sx * (int)k;
            py = y0 + (adx ? sy * (int)((2 * k * ady + adx - 1) / (2 * adx)) : 0);
        } else {
            px = x0 + sx * (int)((2 * k * adx + ady - 1) / (2 * ady));
            py = y0 + sy * (int)k;
        }
        const float t = (float)(int)k / total_steps;            /* raster.c:157 */
        const float z = z0 + t * (z1 - z0);
        float depth;
        if (flags & RC_DEPTH_RANGE_01) depth = (z + 1.0f) * 0.5f;
        else depth = (float)((double)((z + 1.0f) * 0.5f) * (cfg->depth_far - cfg->depth_near) + cfg->depth_near);
        Color4 c = color_lerp({ c0.x, c0.y, c0.z, c0.w }, { c1.x, c1.y, c1.z, c1.w }, t);
        if (flags & RC_TEXTURED) {                              /* raster.c:167-179: lod 0, GL_MODULATE only */
            const float u = uv.x + t * (uv.z - uv.x), v = uv.y + t * (uv.w - uv.y);
            TexTaps T;
            tex_taps(T, cfg, u, v, 0.0f);
            Color4 tc;
            tc.a = tex_channel(T, 24, un);
            /* the reference never terminates when this test fails (`continue` without stepping); skip the pixel instead */
            if ((flags & RC_ALPHA_TEST) && !compare_f(cfg->alpha_func, tc.a, cfg->alpha_ref)) continue;
            tc.r = tex_channel(T, 0, un); tc.g = tex_channel(T, 8, un); tc.b = tex_channel(T, 16, un);
            c = { c.r * tc.r, c.g * tc.g, c.b * tc.b, c.a * tc.a };
        } else if ((flags & RC_ALPHA_TEST) && !compare_f(cfg->alpha_func, c.a, cfg->alpha_ref)) continue;
        if (flags & RC_FOG) {                                   /* raster.c:188-211: negated coordinate */
            const float fc = -(ez0 + t * (ez1 - ez0));
            Color4 fogc = { cfg->fog_color[0], cfg->fog_color[1], cfg->fog_color[2], cfg->fog_color[3] };
            c = color_lerp_rgb(fogc, c, fog_factor(cfg, fc));
        }
        for (int w = -half; w < lw - half; w++) {               /* raster.c:214-226 */
            const int qx = (x_major ? px : px + w) - tile_px, qy = (x_major ? py + w : py) - tile_py;
            if (qx >= X0 && qx <= X1 && qy >= Y0 && qy <= Y1) simple_pixel(sm, cfg, qy * COLOR_PITCH + qx, depth, c);
        }
    }
}

/* ---------------------------------------------------------------- one triangle over one warp region */
/* the part of a record the coverage / depth stage needs (rows 0, 1, state word of row 2, z of row 3) */
struct TriHead {
    int4 row0, row1;
    uint32_t state_flags;
    float z0, z1, z2;
};

__device__ __forceinline__ void load_head(TriHead &h, const TriRecord *rec)
{
    h.row0 = __ldg(reinterpret_cast<const int4 *>(rec) + 0);
    h.row1 = __ldg(reinterpret_cast<const int4 *>(rec) + 1);
    h.state_flags = __ldg(&rec->state_flags);
    const float4 row3 = __ldg(reinterpret_cast<const float4 *>(rec) + 3);
    h.z0 = row3.x; h.z1 = row3.y; h.z2 = row3.z;
}

__device__ __forceinline__ TriHead broadcast_head(const TriHead &h, int src)
{
    TriHead o;
    o.row0.x = __shfl_sync(0xFFFFFFFFu, h.row0.x, src); o.row0.y = __shfl_sync(0xFFFFFFFFu, h.row0.y, src);
    o.row0.z = __shfl_sync(0xFFFFFFFFu, h.row0.z, src); o.row0.w = __shfl_sync(0xFFFFFFFFu, h.row0.w, src);
    o.row1.x = __shfl_sync(0xFFFFFFFFu, h.row1.x, src); o.row1.y = __shfl_sync(0xFFFFFFFFu, h.row1.y, src);
    o.row1.z = __shfl_sync(0xFFFFFFFFu, h.row1.z, src); o.row1.w = __shfl_sync(0xFFFFFFFFu, h.row1.w, src);
    o.state_flags = __shfl_sync(0xFFFFFFFFu, h.state_flags, src);
    o.z0 = __shfl_sync(0xFFFFFFFFu, h.z0, src); o.z1 = __shfl_sync(0xFFFFFFFFu, h.z1, src); o.z2 = __shfl_sync(0xFFFFFFFFu, h.z2, src);
    return o;
}

template <bool VIS, bool PLAIN>
__device__ void raster_triangle(const BatchDev &b, RasterSmem &sm, uint32_t r, const TriHead &h, int tile_px, int tile_py,
                                int X0, int Y0, int X1, int Y1, int rx0, int ry0, bool &pending)
{
    const uint32_t lane = threadIdx.x & 31;
    const float fx0 = (float)h.row0.x, fy0 = (float)h.row0.y, fx1 = (float)h.row0.z, fy1 = (float)h.row0.w;
    const float fx2 = (float)h.row1.x, fy2 = (float)h.row1.y;
    const float area = __int_as_float(h.row1.z), inv_area = __int_as_float(h.row1.w);
    const uint32_t state_flags = h.state_flags;
    const RasterCfg *cfg = b.cfgs + (state_flags & STATE_INDEX_MASK);
    const uint32_t flags = cfg->flags;
    const float z0 = h.z0, z1 = h.z1, z2 = h.z2;
    const bool area_pos = area > 0;
    /* shading can be deferred when nothing between the depth test and the colour write depends on or discards
     * per fragment state: no blending, no alpha test, full colour mask (RC_DEFER, mirrored in the record).
     * The visibility-only kernel only ever sees such records. */
    const bool defer = VIS || (!PLAIN && (state_flags & STATE_DEFER_BIT) != 0);

    if (!VIS && !PLAIN) {
        if (defer) pending = true;
        else if (pending) {     /* deferred colours of earlier triangles must land before an in-order triangle reads them */
            resolve_region(b, sm, rx0, ry0, tile_px, tile_py);
            pending = false;
        }
    }

    for (int by = Y0; by <= Y1; by += 4) {
        for (int bx = X0; bx <= X1; bx += 8) {
            const int x = bx + (int)(lane & 7), y = by + (int)(lane >> 3);
            bool active = (x <= X1) && (y <= Y1);
            const float px = (float)(tile_px + x), py = (float)(tile_py + y);
            const float e0 = edge_at(fx1, fy1, fx2, fy2, px, py);
            const float e1 = edge_at(fx2, fy2, fx0, fy0, px, py);
            const float e2 = edge_at(fx0, fy0, fx1, fy1, px, py);
            /* inclusive on all three edges, no fill rule (raster.c:539-540) */
            active = active && (area_pos ? (e0 >= 0 && e1 >= 0 && e2 >= 0) : (e0 <= 0 && e1 <= 0 && e2 <= 0));

            const float b0 = e0 * inv_area, b1 = e1 * inv_area, b2 = e2 * inv_area;
            const int ci = y * COLOR_PITCH + x;
            float depth = 0.0f;
            if (active) {
                if (flags & RC_DEPTH_TEST) {                    /* the value is only consumed by the depth test / write */
                    const float z = b0 * z0 + b1 * z1 + b2 * z2;
                    if (flags & RC_DEPTH_RANGE_01) depth = (z + 1.0f) * 0.5f;
                    else depth = (float)((double)((z + 1.0f) * 0.5f) * (cfg->depth_far - cfg->depth_near) + cfg->depth_near);   /* raster.c:548 */
                }

                if (flags & RC_STENCIL) {                       /* raster.c:550-579 */
                    const int si = y * STENCIL_PITCH + x;
                    const uint8_t sval = sm.stencil[si];
                    const int32_t mref = (int32_t)((uint32_t)cfg->stencil_ref & cfg->stencil_mask);
                    const int32_t mval = (int32_t)((uint32_t)sval & cfg->stencil_mask);
                    const uint8_t wm = (uint8_t)(cfg->stencil_writemask & 0xFF);
                    uint32_t op;
                    if (!compare_i(cfg->stencil_func, mref, mval)) { op = cfg->stencil_fail; active = false; }
                    else if ((flags & RC_DEPTH_TEST) && !compare_f(cfg->depth_func, depth, sm.depth[ci])) { op = cfg->stencil_zfail; active = false; }
                    else op = cfg->stencil_zpass;
                    const uint8_t nv = stencil_apply(op, sval, cfg->stencil_ref);
                    sm.stencil[si] = (uint8_t)((sval & ~wm) | (nv & wm));
                } else if (flags & RC_DEPTH_TEST) {
                    if (!compare_f(cfg->depth_func, depth, sm.depth[ci])) active = false;
                }
            }
            const bool depth_write = (flags & (RC_DEPTH_TEST | RC_DEPTH_WRITE)) == (RC_DEPTH_TEST | RC_DEPTH_WRITE);

            if (defer) {
                /* nothing can discard the fragment any more: the depth write of raster.c:707-710 happens now and
                 * the fragment becomes the pixel's visible one; its colour is computed by resolve_region */
                if (active) {
                    if (depth_write) sm.depth[ci] = depth;
                    sm.vis[ci] = r;
                }
            } else if (!VIS && active) shade_now(b, sm, r, state_flags, cfg, b0, b1, b2, ci, depth, depth_write);
        }
    }
}

/* ---------------------------------------------------------------- tile load / clear / store */
template <bool VIS>
__device__ void tile_init(RasterSmem &sm, const FrameTargets &fb, const ClearOp &clr, uint32_t planes, int px0, int py0,
                          int vw, int vh)
{
    if (VIS) planes &= ~1u;         /* the colour plane is produced by the shade pass */
    /* clear rectangle relative to the tile */
    const int cx0 = max(clr.x0 - px0, 0), cy0 = max(clr.y0 - py0, 0);
    const int cx1 = min(clr.x1 - px0, vw), cy1 = min(clr.y1 - py0, vh);
    const bool clr_any = clr.mask && cx0 < cx1 && cy0 < cy1;
    const bool clr_full = clr_any && cx0 == 0 && cy0 == 0 && cx1 == vw && cy1 == vh;
    const bool vec = (vw == TILE_W) && ((fb.width & 3) == 0);

    if (planes & 1u) {
        const bool cl = clr_any && (clr.mask & G_COLOR_BUFFER_BIT);
        if (!(cl && clr_full)) {
            if (vec) {
                for (int i = threadIdx.x; i < vh * 16; i += RASTER_THREADS) {
                    int y = i >> 4, q = i & 15;
                    uint4 v = *reinterpret_cast<const uint4 *>(fb.color + (size_t)(py0 + y) * fb.width + px0 + q * 4);
                    *reinterpret_cast<uint4 *>(&sm.color[y * COLOR_PITCH + q * 4]) = v;
                }
            } else {
                for (int i = threadIdx.x; i < vh * TILE_W; i += RASTER_THREADS) {
                    int y = i >> 6, x = i & 63;
                    if (x < vw) sm.color[y * COLOR_PITCH + x] = fb.color[(size_t)(py0 + y) * fb.width + px0 + x];
                }
            }
        }
        if (cl) {
            if (!clr_full) __syncthreads();
            for (int i = threadIdx.x; i < vh * TILE_W; i += RASTER_THREADS) {
                int y = i >> 6, x = i & 63;
                if (x >= cx0 && x < cx1 && y >= cy0 && y < cy1) sm.color[y * COLOR_PITCH + x] = clr.color;
            }
        }
    }
    if (planes & 2u) {
        const bool cl = clr_any && (clr.mask & G_DEPTH_BUFFER_BIT);
        if (!(cl && clr_full)) {
            if (vec) {
                for (int i = threadIdx.x; i < vh * 16; i += RASTER_THREADS) {
                    int y = i >> 4, q = i & 15;
                    float4 v = *reinterpret_cast<const float4 *>(fb.depth + (size_t)(py0 + y) * fb.width + px0 + q * 4);
                    *reinterpret_cast<float4 *>(&sm.depth[y * COLOR_PITCH + q * 4]) = v;
                }
            } else {
                for (int i = threadIdx.x; i < vh * TILE_W; i += RASTER_THREADS) {
                    int y = i >> 6, x = i & 63;
                    if (x < vw) sm.depth[y * COLOR_PITCH + x] = fb.depth[(size_t)(py0 + y) * fb.width + px0 + x];
                }
            }
        }
        if (cl) {
            if (!clr_full) __syncthreads();
            for (int i = threadIdx.x; i < vh * TILE_W; i += RASTER_THREADS) {
                int y = i >> 6, x = i & 63;
                if (x >= cx0 && x < cx1 && y >= cy0 && y < cy1) sm.depth[y * COLOR_PITCH + x] = clr.depth;
            }
        }
    }
    if (planes & 4u) {
        const bool cl = clr_any && (clr.mask & G_STENCIL_BUFFER_BIT);
        if (!(cl && clr_full)) {
            if (vec && (fb.width & 15) == 0) {
                for (int i = threadIdx.x; i < vh * 4; i += RASTER_THREADS) {
                    int y = i >> 2, q = i & 3;
                    uint4 v = *reinterpret_cast<const uint4 *>(fb.stencil + (size_t)(py0 + y) * fb.width + px0 + q * 16);
                    *reinterpret_cast<uint4 *>(&sm.stencil[y * STENCIL_PITCH + q * 16]) = v;
                }
            } else {
                for (int i = threadIdx.x; i < vh * TILE_W; i += RASTER_THREADS) {
                    int y = i >> 6, x = i & 63;
                    if (x < vw) sm.stencil[y * STENCIL_PITCH + x] = fb.stencil[(size_t)(py0 + y) * fb.width + px0 + x];
                }
            }
        }
        if (cl) {
            if (!clr_full) __syncthreads();
            for (int i = threadIdx.x; i < vh * TILE_W; i += RASTER_THREADS) {
                int y = i >> 6, x = i & 63;
                if (x >= cx0 && x < cx1 && y >= cy0 && y < cy1) sm.stencil[y * STENCIL_PITCH + x] = (uint8_t)clr.stencil;
            }
        }
    }
}

template <bool VIS>
__device__ void tile_store(RasterSmem &sm, const FrameTargets &fb, uint32_t planes, int px0, int py0, int vw, int vh, uint32_t *vis_plane)
{
    const bool vec = (vw == TILE_W) && ((fb.width & 3) == 0);
    if (VIS) {
        planes &= ~1u;
        /* visibility entries of the tile -> HBM for the shade pass (every pixel, including "none") */
        if (vec) {
            for (int i = threadIdx.x; i < vh * 16; i += RASTER_THREADS) {
                int y = i >> 4, q = i & 15;
                *reinterpret_cast<uint4 *>(vis_plane + (size_t)(py0 + y) * fb.width + px0 + q * 4) =
                    *reinterpret_cast<const uint4 *>(&sm.vis[y * COLOR_PITCH + q * 4]);
            }
        } else {
            for (int i = threadIdx.x; i < vh * TILE_W; i += RASTER_THREADS) {
                int y = i >> 6, x = i & 63;
                if (x < vw) vis_plane[(size_t)(py0 + y) * fb.width + px0 + x] = sm.vis[y * COLOR_PITCH + x];
            }
        }
    }
    if (planes & 1u) {
        if (vec) {
            for (int i = threadIdx.x; i < vh * 16; i += RASTER_THREADS) {
                int y = i >> 4, q = i & 15;
                const uint4 v = *reinterpret_cast<const uint4 *>(&sm.color[y * COLOR_PITCH + q * 4]);
                *reinterpret_cast<uint4 *>(fb.color + (size_t)(py0 + y) * fb.width + px0 + q * 4) = v;
                if (fb.present) *reinterpret_cast<uint4 *>(fb.present + (size_t)(py0 + y) * fb.width + px0 + q * 4) = v;
            }
        } else {
            for (int i = threadIdx.x; i < vh * TILE_W; i += RASTER_THREADS) {
                int y = i >> 6, x = i & 63;
                if (x < vw) {
                    fb.color[(size_t)(py0 + y) * fb.width + px0 + x] = sm.color[y * COLOR_PITCH + x];
                    if (fb.present) fb.present[(size_t)(py0 + y) * fb.width + px0 + x] = sm.color[y * COLOR_PITCH + x];
                }
            }
        }
    }
    if (planes & 2u) {
        if (vec) {
            for (int i = threadIdx.x; i < vh * 16; i += RASTER_THREADS) {
                int y = i >> 4, q = i & 15;
                *reinterpret_cast<float4 *>(fb.depth + (size_t)(py0 + y) * fb.width + px0 + q * 4) =
                    *reinterpret_cast<const float4 *>(&sm.depth[y * COLOR_PITCH + q * 4]);
            }
        } else {
            for (int i = threadIdx.x; i < vh * TILE_W; i += RASTER_THREADS) {
                int y = i >> 6, x = i & 63;
                if (x < vw) fb.depth[(size_t)(py0 + y) * fb.width + px0 + x] = sm.depth[y * COLOR_PITCH + x];
            }
        }
    }
    if (planes & 4u) {
        if (vec && (fb.width & 15) == 0) {
            for (int i = threadIdx.x; i < vh * 4; i += RASTER_THREADS) {
                int y = i >> 2, q = i & 3;
                *reinterpret_cast<uint4 *>(fb.stencil + (size_t)(py0 + y) * fb.width + px0 + q * 16) =
                    *reinterpret_cast<const uint4 *>(&sm.stencil[y * STENCIL_PITCH + q * 16]);
            }
        } else {
            for (int i = threadIdx.x; i < vh * TILE_W; i += RASTER_THREADS) {
                int y = i >> 6, x = i & 63;
                if (x < vw) fb.stencil[(size_t)(py0 + y) * fb.width + px0 + x] = sm.stencil[y * STENCIL_PITCH + x];
            }
        }
    }
}

/* ---------------------------------------------------------------- list staging */
__device__ __forceinline__ uint32_t pack_box(uint2 bb, int px0, int py0)
{
    int x0 = max((int)(bb.x & 0xFFFFu) - px0, 0), y0 = max((int)(bb.x >> 16) - py0, 0);
    int x1 = min((int)(bb.y & 0xFFFFu) - px0, TILE_W - 1), y1 = min((int)(bb.y >> 16) - py0, TILE_H - 1);
    return (uint32_t)x0 | ((uint32_t)y0 << 8) | ((uint32_t)x1 << 16) | ((uint32_t)y1 << 24);
}

/* add one reference's estimated cost to every region its (tile-relative) box overlaps */
__device__ __forceinline__ void account_regions(RasterSmem &sm, uint32_t box)
{
    const int x0 = box & 0xFF, y0 = (box >> 8) & 0xFF, x1 = (box >> 16) & 0xFF, y1 = box >> 24;
    for (int ry = y0 / REGION_H; ry <= y1 / REGION_H; ry++)
        for (int rx = x0 / REGION_W; rx <= x1 / REGION_W; rx++) {
            const int w = min(x1, rx * REGION_W + REGION_W - 1) - max(x0, rx * REGION_W) + 1;
            const int h = min(y1, ry * REGION_H + REGION_H - 1) - max(y0, ry * REGION_H) + 1;
            atomicAdd(&sm.region_work[ry * REGIONS_X + rx], 2u + (uint32_t)(((w + 7) >> 3) * ((h + 3) >> 2)));
        }
}

/* sort the staged (key, rec, box) triples by key; n is padded to a power of two with key = ~0 */
__device__ void sort_window(RasterSmem &sm, uint32_t n)
{
    uint32_t p = 32;
    while (p < n) p <<= 1;
    for (uint32_t i = n + threadIdx.x; i < p; i += RASTER_THREADS) sm.key[i] = 0xFFFFFFFFu;
    for (uint32_t i = threadIdx.x; i < p; i += RASTER_THREADS) sm.pos[i] = (uint16_t)i;
    if (threadIdx.x == 0) sm.next_region = 0;
    if (threadIdx.x < NUM_REGIONS) {        /* longest-processing-time-first order of the regions */
        const uint32_t mine = sm.region_work[threadIdx.x];
        uint32_t rank = 0;
        for (int k = 0; k < NUM_REGIONS; k++) {
            const uint32_t other = sm.region_work[k];
            rank += (other > mine || (other == mine && k < (int)threadIdx.x)) ? 1u : 0u;
        }
        sm.region_order[rank] = threadIdx.x;
    }
    __syncthreads();
    if (threadIdx.x < NUM_REGIONS) sm.region_work[threadIdx.x] = 0;
    /* Bitonic sort of (key, original position).  Exchanges at distance >= 32 use the whole CTA and a barrier;
     * the runs of exchanges at distance 16..1 stay inside one aligned group of 32 elements, which a single warp
     * owns, so they only need warp-level synchronisation.  rec/box are permuted once at the end. */
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto exchange = [&](uint32_t i, uint32_t j, uint32_t k) {
        const uint32_t ixj = i ^ j;
        if (ixj > i) {
            const uint32_t a = sm.key[i], c = sm.key[ixj];
            const bool up = ((i & k) == 0);
            if ((a > c) == up) {
                sm.key[i] = c; sm.key[ixj] = a;
                const uint16_t t = sm.pos[i]; sm.pos[i] = sm.pos[ixj]; sm.pos[ixj] = t;
            }
        }
    };
    for (uint32_t k = 2; k <= p; k <<= 1) {
        uint32_t j = k >> 1;
        for (; j >= 32; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < p; i += RASTER_THREADS) exchange(i, j, k);
            __syncthreads();
        }
        for (uint32_t base = warp * 32; base < p; base += RASTER_THREADS) {
            for (uint32_t jj = j; jj > 0; jj >>= 1) {
                exchange(base + lane, jj, k);
                __syncwarp();
            }
        }
        __syncthreads();
    }
    uint32_t tr[LIST_WINDOW / RASTER_THREADS], tb[LIST_WINDOW / RASTER_THREADS];
#pragma unroll
    for (int q = 0; q < LIST_WINDOW / RASTER_THREADS; q++) {
        const uint32_t i = threadIdx.x + (uint32_t)q * RASTER_THREADS;
        if (i < n) { const uint32_t from = sm.pos[i]; tr[q] = sm.rec[from]; tb[q] = sm.box[from]; }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < LIST_WINDOW / RASTER_THREADS; q++) {
        const uint32_t i = threadIdx.x + (uint32_t)q * RASTER_THREADS;
        if (i < n) { sm.rec[i] = tr[q]; sm.box[i] = tb[q]; }
    }
    __syncthreads();
}

/* Rasterise the n staged (sorted) references.  The tile is cut into 16 regions of 16x16 pixels; a warp takes
 * the next unprocessed region from a shared counter and walks the whole window for it, so all fragments of a
 * pixel are produced by one warp in submission order while the regions balance the load between warps. */
template <bool VIS, bool PLAIN>
__device__ void process_window(const BatchDev &b, RasterSmem &sm, uint32_t n, int px0, int py0)
{
    const uint32_t lane = threadIdx.x & 31;
    for (;;) {
        uint32_t region = 0;
        if (lane == 0) region = atomicAdd(&sm.next_region, 1u);
        region = __shfl_sync(0xFFFFFFFFu, region, 0);
        if (region >= (uint32_t)NUM_REGIONS) break;
        region = sm.region_order[region];
        const int rx0 = (int)(region % REGIONS_X) * REGION_W, ry0 = (int)(region / REGIONS_X) * REGION_H;
        const int rx1 = rx0 + REGION_W - 1, ry1 = ry0 + REGION_H - 1;
        bool pending = false;
        for (uint32_t base = 0; base < n; base += 32) {
            uint32_t e = base + lane;
            uint32_t box = 0;
            bool hit = false;
            if (e < n) {
                box = sm.box[e];
                int x0 = box & 0xFF, y0 = (box >> 8) & 0xFF, x1 = (box >> 16) & 0xFF, y1 = box >> 24;
                hit = !(x1 < rx0 || x0 > rx1 || y1 < ry0 || y0 > ry1);
            }
            /* every lane with a hit fetches the head of its own record: up to 32 record fetches in flight at once,
             * handed to the whole warp by shuffles when the hit's turn comes */
            TriHead mine;
            if (VIS) {      /* (the general kernel is register-bound: it loads the head uniformly when the hit's turn comes) */
                if (hit) load_head(mine, b.records + sm.rec[e]);
                else { mine.row0 = make_int4(0, 0, 0, 0); mine.row1 = make_int4(0, 0, 0, 0); mine.state_flags = 0; mine.z0 = mine.z1 = mine.z2 = 0.0f; }
            }
            uint32_t mask = __ballot_sync(0xFFFFFFFFu, hit);
            while (mask) {
                int k = __ffs(mask) - 1;
                mask &= mask - 1;
                uint32_t bx = __shfl_sync(0xFFFFFFFFu, box, k);
                uint32_t r = sm.rec[base + k];
                int X0 = max((int)(bx & 0xFF), rx0), Y0 = max((int)((bx >> 8) & 0xFF), ry0);
                int X1 = min((int)((bx >> 16) & 0xFF), rx1), Y1 = min((int)(bx >> 24), ry1);
                TriHead h;
                if (VIS) h = broadcast_head(mine, k);
                else load_head(h, b.records + r);
                const uint32_t kind = (VIS || PLAIN) ? KIND_TRIANGLE : (h.state_flags & STATE_KIND_MASK) >> STATE_KIND_SHIFT;
                if (kind == KIND_TRIANGLE) raster_triangle<VIS, PLAIN>(b, sm, r, h, px0, py0, X0, Y0, X1, Y1, rx0, ry0, pending);
                else {
                    if (pending) {      /* lines and points are drawn in order on top of whatever was deferred */
                        resolve_region(b, sm, rx0, ry0, px0, py0);
                        pending = false;
                    }
                    if (kind == KIND_POINT) raster_point(b, sm, r, X0, Y0, X1, Y1);
                    else raster_line(b, sm, r, px0, py0, X0, Y0, X1, Y1);
                }
                __syncwarp();
            }
        }
        if (!VIS && !PLAIN && pending) resolve_region(b, sm, rx0, ry0, px0, py0);
    }
}

/* VIS = true : K4a, the visibility pass.  Handles the tiles all of whose records are deferrable: resolves coverage,
 *              stencil and depth in shared memory and leaves one record index per pixel for k_shade (K4b).  No colour
 *              plane, no shading code: half the registers and less shared memory than the general kernel.
 * VIS = false: the general kernel (in-order shading, blending, alpha test).  With only_flagged it handles just
 *              the tiles that reference at least one non-deferrable record. */
/* PLAIN (general kernel only): the pass has nothing but in-order filled triangles -- no deferrable record can reach a
 * general tile, no lines, no points -- so the deferred-shading bookkeeping and the kind dispatch are compiled out
 * (the fill-rate case C3: fewer live registers in the block loop). */
template <bool VIS, bool PLAIN>
__global__ void __launch_bounds__(RASTER_THREADS, VIS ? 3 : 2) k_raster(BatchDev b, FrameTargets fb, ClearOp clr, uint32_t planes, uint32_t only_flagged, uint32_t fill_mode)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RasterSmem &sm = *reinterpret_cast<RasterSmem *>(smem_raw);
    if (!lists_fit(b)) return;

    const uint32_t tile = b.tile_order ? b.tile_order[blockIdx.x] : blockIdx.x;
    const int tx = (int)(tile % (uint32_t)fb.tiles_x), ty = (int)(tile / (uint32_t)fb.tiles_x) + fb.tile_y0;
    const int px0 = tx << TILE_LOG, py0t = ty << TILE_LOG;
    /* rows of this tile owned by the band, columns inside the framebuffer */
    const int py0 = max(py0t, fb.band_y0);
    const int vw = min(TILE_W, fb.width - px0);
    const int vh = min(py0t + TILE_H, fb.band_y1) - py0;
    if (vw <= 0 || vh <= 0) return;

    const uint32_t L = b.tile_count ? b.tile_count[tile] : 0u;
    const uint32_t flagged = (b.tile_count && L) ? b.tile_flags[tile] : 0u;
    /* flags 0: the order-independent kernel (k_vis.cu) owns the tile, also when it only needs clearing; 2: sorted
     * visibility kernel; bit 0 set: general kernel */
    if (VIS ? (flagged != 2u) : (only_flagged && !(flagged & 1u))) return;
    const bool clr_here = clr.mask && clr.x0 < px0 + vw && clr.x1 > px0 && clr.y0 < py0 + vh && clr.y1 > py0;
    if (L == 0 && !clr_here) return;
    /* in-order tiles of large triangles belong to the pixel-owner kernel (k_fill.cu), which is launched over the same grid */
    if (!VIS && L && fill_owns_tile(b, fill_mode, L, flagged, b.tile_list + b.tile_offset[tile], px0, py0, vh, &sm.count)) return;

    for (int i = threadIdx.x; i < 256; i += RASTER_THREADS) sm.unorm8[i] = b.unorm8[i];
    if (threadIdx.x < NUM_REGIONS) sm.region_work[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < TILE_H * COLOR_PITCH; i += RASTER_THREADS) sm.vis[i] = VIS_NONE;
    /* shared-memory row 0 is framebuffer row py0 (the first row of the tile inside the band) */
    tile_init<VIS>(sm, fb, clr, planes, px0, py0, vw, vh);
    __syncthreads();

    if (L > 0) {
        const uint32_t *list = b.tile_list + b.tile_offset[tile];
        if (L <= (uint32_t)LIST_WINDOW) {
            for (uint32_t i = threadIdx.x; i < L; i += RASTER_THREADS) {
                uint32_t r = list[i];
                uint4 row2 = __ldg(reinterpret_cast<const uint4 *>(b.records + r) + 2);
                sm.key[i] = row2.w;
                sm.rec[i] = r;
                const uint32_t bx = pack_box(make_uint2(row2.x, row2.y), px0, py0);
                sm.box[i] = bx;
                account_regions(sm, bx);
            }
            __syncthreads();
            sort_window(sm, L);
            process_window<VIS, PLAIN>(b, sm, L, px0, py0);
        } else {
            /* Long list: take the references in windows of increasing id.  The upper id bound of a
             * window is found by bisection on the id value, counting with the whole CTA. */
            unsigned long long lo = 0;
            uint32_t done = 0;
            while (done < L) {
                unsigned long long a = lo + 1, z = 0x100000000ull;     /* answer in [a, z]: largest hi with count(lo<=id<hi) <= WINDOW */
                while (a < z) {
                    unsigned long long mid = (a + z + 1) >> 1;
                    uint32_t cnt = 0;
                    for (uint32_t i = threadIdx.x; i < L; i += RASTER_THREADS) {
                        uint32_t id = __ldg(&b.records[list[i]].id);
                        cnt += (id >= lo && id < mid) ? 1u : 0u;
                    }
                    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
                    if ((threadIdx.x & 31) == 0) sm.scratch[threadIdx.x >> 5] = cnt;
                    __syncthreads();
                    uint32_t total = 0;
                    for (int w = 0; w < RASTER_THREADS / 32; w++) total += sm.scratch[w];
                    __syncthreads();
                    if (total <= (uint32_t)LIST_WINDOW) a = mid; else z = mid - 1;
                }
                const unsigned long long hi = a;
                if (threadIdx.x == 0) sm.count = 0;
                __syncthreads();
                for (uint32_t i = threadIdx.x; i < L; i += RASTER_THREADS) {
                    uint32_t r = list[i];
                    uint4 row2 = __ldg(reinterpret_cast<const uint4 *>(b.records + r) + 2);
                    if (row2.w >= lo && row2.w < hi) {
                        uint32_t at = atomicAdd(&sm.count, 1u);
                        sm.key[at] = row2.w;
                        sm.rec[at] = r;
                        const uint32_t bx = pack_box(make_uint2(row2.x, row2.y), px0, py0);
                        sm.box[at] = bx;
                        account_regions(sm, bx);
                    }
                }
                __syncthreads();
                const uint32_t n = sm.count;
                __syncthreads();
                sort_window(sm, n);
                process_window<VIS, PLAIN>(b, sm, n, px0, py0);
                __syncthreads();
                done += n;
                lo = hi;
            }
        }
    }
    __syncthreads();
    tile_store<VIS>(sm, fb, planes, px0, py0, vw, vh, b.vis_plane);
}

void launch_raster(const BatchDev &b, const FrameTargets &fb, const ClearOp &clear, uint32_t planes,
                   const RasterPlan &plan, cudaStream_t s, cudaEvent_t ev_vis, cudaEvent_t ev_shade)
{
    const bool any_deferrable = plan.any_deferrable, any_in_order = plan.any_in_order;
    static bool configured[64] = { false };
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem_vis = offsetof(RasterSmem, color);
    if (dev >= 0 && dev < 64 && !configured[dev]) {      /* the opt-in is per device */
        cudaFuncSetAttribute(k_raster<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RasterSmem));
        cudaFuncSetAttribute(k_raster<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RasterSmem));
        cudaFuncSetAttribute(k_raster<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_vis);
        configured[dev] = true;
    }
    uint32_t tiles = (uint32_t)(fb.tiles_x * fb.tile_rows);
    if (tiles == 0 || planes == 0) { cudaEventRecord(ev_vis, s); cudaEventRecord(ev_shade, s); return; }
    /* Split path: tiles whose records are all deferrable go through K4a (visibility) + K4b (shade); tiles with an
     * in-order record go through the general kernel.  A pass without deferrable draws uses the general kernel alone
     * (it then also owns the tiles that only need clearing). */
    const bool split = any_deferrable && b.tile_count != nullptr && b.vis_plane != nullptr;
    if (split) {
        /* tiles whose records are all order-independent (and the tiles that only need clearing) */
        launch_vis_unordered(b, fb, clear, planes, plan.unordered_func ? plan.unordered_func : 1u, plan.unordered_range01, s);
        if (plan.any_ordered_vis) {
            k_raster<true, false><<<tiles, RASTER_THREADS, smem_vis, s>>>(b, fb, clear, planes, 0u, 0u);
            note_launch();
        }
        cudaEventRecord(ev_vis, s);
        if (plan.color_gate) cudaStreamWaitEvent(s, plan.color_gate, 0);       /* visibility above ran ahead; colour waits */
        if (planes & 1u) launch_shade(b, fb, clear, plan.in_order_all, plan.in_order_any, s);
        cudaEventRecord(ev_shade, s);
        if (any_in_order) {
            k_raster<false, false><<<tiles, RASTER_THREADS, sizeof(RasterSmem), s>>>(b, fb, clear, planes, 1u, plan.fill_mode);
            note_launch();
            launch_fill(b, fb, clear, planes, plan.fill_mode, plan.in_order_all, plan.in_order_any, s);
        }
    } else {
        if (plan.color_gate) cudaStreamWaitEvent(s, plan.color_gate, 0);
        cudaEventRecord(ev_vis, s);
        cudaEventRecord(ev_shade, s);
        if (plan.plain_in_order) k_raster<false, true><<<tiles, RASTER_THREADS, sizeof(RasterSmem), s>>>(b, fb, clear, planes, 0u, plan.fill_mode);
        else k_raster<false, false><<<tiles, RASTER_THREADS, sizeof(RasterSmem), s>>>(b, fb, clear, planes, 0u, plan.fill_mode);
        note_launch();
        if (b.tile_count) launch_fill(b, fb, clear, planes, plan.fill_mode, plan.in_order_all, plan.in_order_any, s);
    }
}

} // namespace mtgl_dev_impl
