/*
 * k_setup.cu -- K2: primitive assembly, frustum clipping, perspective divide, integer snapping,
 * culling and triangle set-up, one thread per assembled triangle.
 *
 * Replaces flush_triangles/quads/triangle_strip/triangle_fan/polygon/quad_strip
 * (src/raster.c:961-1017, 1199-1231), render_triangle (901-958), clip_triangle and friends
 * (src/clipping.h:27-126), perspective_divide (raster.c:729-746), ndc_to_screen (59-63),
 * should_cull (751-774) and the set-up half of rasterize_triangle_smooth (458-529).
 *
 * Work decomposition: a CTA owns a chunk of 256 consecutive input triangles, a warp a group of 32.
 * Every thread first counts how many sub-triangles of its triangle survive (0 or 1 without clipping,
 * up to 7 with), a warp scan turns the counts into submission-ordered indices inside the group, one
 * atomicAdd per warp reserves the group's slots in the record array, and the thread writes its
 * 160-byte records.  A record's id = group << 10 | index-in-group is ordered exactly like the
 * reference's sequential loop; the tile kernels sort and compare by it.
 *
 * Algorithmic bytes per input triangle: 3 x 48 B gathered (+ 3 x 32 B for per-fragment lighting),
 * 160 B written per surviving sub-triangle.
 */
#include "dev_common.cuh"
#include "dev_texture.cuh"
#include "dev_vertex.cuh"

#include <climits>

namespace mtgl_dev_impl {

void note_launch();

struct SVert {              /* vertex_t (graphics.h:438-446) minus the unused object normal */
    float x, y, z, w;
    float r, g, b, a;
    float u, v, ez;
    float epx, epy, epz, enx, eny, enz;
};

__device__ __forceinline__ float plane_dist(const SVert &v, int plane)   /* clipping.h:27-32: near far left right bottom top */
{
    switch (plane) {
    case 0: return v.z + v.w;
    case 1: return v.w - v.z;
    case 2: return v.x + v.w;
    case 3: return v.w - v.x;
    case 4: return v.y + v.w;
    default: return v.w - v.y;
    }
}

__device__ __forceinline__ bool inside_all(const SVert &v)
{
    return (v.z + v.w) >= 0 && (v.w - v.z) >= 0 && (v.x + v.w) >= 0 && (v.w - v.x) >= 0 && (v.y + v.w) >= 0 && (v.w - v.y) >= 0;
}

__device__ __forceinline__ float lerp1(float a, float b, float t) { return a + t * (b - a); }   /* graphics.h:369-371 */

__device__ SVert vertex_lerp(const SVert &a, const SVert &b, float t)   /* graphics.h:465-475 */
{
    SVert o;
    o.x = lerp1(a.x, b.x, t); o.y = lerp1(a.y, b.y, t); o.z = lerp1(a.z, b.z, t); o.w = lerp1(a.w, b.w, t);
    float s = 1.0f - t;                                     /* colour uses a*(1-t) + b*t (graphics.h:293-295) */
    o.r = a.r * s + b.r * t; o.g = a.g * s + b.g * t; o.b = a.b * s + b.b * t; o.a = a.a * s + b.a * t;
    o.u = lerp1(a.u, b.u, t); o.v = lerp1(a.v, b.v, t);
    o.ez = lerp1(a.ez, b.ez, t);
    o.epx = lerp1(a.epx, b.epx, t); o.epy = lerp1(a.epy, b.epy, t); o.epz = lerp1(a.epz, b.epz, t);
    o.enx = lerp1(a.enx, b.enx, t); o.eny = lerp1(a.eny, b.eny, t); o.enz = lerp1(a.enz, b.enz, t);
    return o;
}

__device__ __forceinline__ void snap(SVert &v, int plane)   /* clipping.h:37-47 */
{
    switch (plane) {
    case 0: v.z = -v.w; break;
    case 1: v.z = v.w; break;
    case 2: v.x = -v.w; break;
    case 3: v.x = v.w; break;
    case 4: v.y = -v.w; break;
    default: v.y = v.w; break;
    }
}

#define MAX_CLIP 12

__device__ __noinline__ int clip_plane(const SVert *in, int n, SVert *out, int plane)   /* clipping.h:50-99 */
{
    if (n == 0) return 0;
    int m = 0;
    int prev = n - 1;
    float pd = plane_dist(in[prev], plane);
    for (int i = 0; i < n; i++) {
        float cd = plane_dist(in[i], plane);
        if (pd >= 0) {
            if (cd >= 0) out[m++] = in[i];
            else {
                float den = pd - cd;
                if (fabsf(den) > 1e-10f) {
                    out[m] = vertex_lerp(in[prev], in[i], pd / den);
                    snap(out[m], plane);
                    m++;
                }
            }
        } else if (cd >= 0) {
            float den = pd - cd;
            if (fabsf(den) > 1e-10f) {
                out[m] = vertex_lerp(in[prev], in[i], pd / den);
                snap(out[m], plane);
                m++;
            }
            out[m++] = in[i];
        }
        prev = i;
        pd = cd;
    }
    return m;
}

__device__ __forceinline__ void persp_divide(SVert &v)   /* raster.c:729-746: w becomes 1/w */
{
    if (fabsf(v.w) < 1e-6f) { v.x = 0.0f; v.y = 0.0f; v.z = 0.0f; v.w = 1.0f; return; }
    float iw = 1.0f / v.w;
    v.x *= iw; v.y *= iw; v.z *= iw; v.w = iw;
}

__device__ __forceinline__ int imin3(int a, int b, int c) { return min(a, min(b, c)); }
__device__ __forceinline__ int imax3(int a, int b, int c) { return max(a, max(b, c)); }

/* ---------------------------------------------------------------- record emission */
/* where the binner's first pass (count references per tile, k_bin.cu) is fused into set-up */
struct BinOut {
    TriRecord *records;
    uint4 *bin_rows;
    uint32_t *tile_count, *tile_flags, *large_list;
    DevCounters *counters;
    int32_t tiles_x, tile_y0;
};

/* one record's contribution to the per-tile reference counts; records spanning many tiles go to the cooperative binner */
__device__ __forceinline__ void count_tiles_single(const BinOut &bo, uint32_t r, uint32_t bbox_min, uint32_t bbox_max, uint32_t state_flags)
{
    const int tx0 = (int)(bbox_min & 0xFFFFu) >> TILE_LOG, tx1 = (int)(bbox_max & 0xFFFFu) >> TILE_LOG;
    const int ty0 = ((int)(bbox_min >> 16) >> TILE_LOG) - bo.tile_y0, ty1 = ((int)(bbox_max >> 16) >> TILE_LOG) - bo.tile_y0;
    if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > LARGE_TILES) {
        bo.large_list[atomicAdd(&bo.counters->large_count, 1u)] = r;
        return;
    }
    const uint32_t tflags = tile_flag_bits(state_flags);
    for (int ty = ty0; ty <= ty1; ty++)
        for (int tx = tx0; tx <= tx1; tx++) {
            const uint32_t tile = (uint32_t)(ty * bo.tiles_x + tx);
            atomicAdd(&bo.tile_count[tile], 1u);
            if (tflags) atomicOr(&bo.tile_flags[tile], tflags);
        }
}

struct Emitter {
    bool write;             /* false: counting pass */
    uint32_t n;             /* records emitted so far by this thread */
    TriRecord *dst;         /* first slot of this thread (write pass) */
    TriEye *eye_dst;
    uint32_t id0;           /* ordered id of the first slot */
    const BinOut *bin;
};

__device__ __forceinline__ void store_record(TriRecord *dst, const TriRecord &r)
{
    const int4 *s = reinterpret_cast<const int4 *>(&r);
    int4 *d = reinterpret_cast<int4 *>(dst);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(TriRecord) / 16); k++) d[k] = s[k];
}

__device__ __forceinline__ void emit(Emitter &em, TriRecord &rec, const TriEye *eye)
{
    if (em.write) {
        rec.id = em.id0 + em.n;
        store_record(em.dst + em.n, rec);
        if (eye && em.eye_dst) em.eye_dst[em.n] = *eye;
        const uint32_t r = (uint32_t)(em.dst + em.n - em.bin->records);
        em.bin->bin_rows[r] = make_uint4(rec.bbox_min, rec.bbox_max, rec.state_flags, rec.id);
        count_tiles_single(*em.bin, r, rec.bbox_min, rec.bbox_max, rec.state_flags);
    }
    em.n++;
}

/* clamp an inclusive pixel box to framebuffer, scissor and band; false when empty */
__device__ __forceinline__ bool clamp_box(const mtgl_state *st, const FrameTargets &fb, int &minX, int &minY, int &maxX, int &maxY)
{
    if (st->caps & MTGL_CAP_SCISSOR_TEST) {
        const int32_t *sc = st->scissor;
        if (minX < sc[0]) minX = sc[0];
        if (minY < sc[1]) minY = sc[1];
        if (maxX >= sc[0] + sc[2]) maxX = sc[0] + sc[2] - 1;
        if (maxY >= sc[1] + sc[3]) maxY = sc[1] + sc[3] - 1;
    }
    if (minX < 0) minX = 0;
    if (maxX >= fb.width) maxX = fb.width - 1;
    if (minY < fb.band_y0) minY = fb.band_y0;
    if (maxY >= fb.band_y1) maxY = fb.band_y1 - 1;
    return minX <= maxX && minY <= maxY;
}

/* depth of a line / point pixel (raster.c:161, 793, 1065) */
__device__ __forceinline__ float simple_depth(const RasterCfg *cfg, float z)
{
    if (cfg->flags & RC_DEPTH_RANGE_01) return (z + 1.0f) * 0.5f;
    return (float)((double)((z + 1.0f) * 0.5f) * (cfg->depth_far - cfg->depth_near) + cfg->depth_near);
}

/* fog of lines and points: the coordinate is negated once more (raster.c:189, 809, 1093) */
__device__ __forceinline__ Color4 simple_fog(const RasterCfg *cfg, Color4 c, float eye_z)
{
    if (!(cfg->flags & RC_FOG)) return c;
    float fc = -eye_z, f = 1.0f;
    switch (cfg->fog_mode) {
    case G_LINEAR: if (cfg->fog_end != cfg->fog_start) f = (cfg->fog_end - fc) / (cfg->fog_end - cfg->fog_start); break;
    case G_EXP: f = expf(-cfg->fog_density * fc); break;
    case G_EXP2: { float d = cfg->fog_density * fc; f = expf(-d * d); break; }
    default: break;
    }
    if (f < 0.0f) f = 0.0f;
    if (f > 1.0f) f = 1.0f;
    Color4 fogc = { cfg->fog_color[0], cfg->fog_color[1], cfg->fog_color[2], cfg->fog_color[3] };
    return color_lerp_rgb(fogc, c, f);
}

/* a LINE record for draw_line_full(x0, y0, z0, x1, y1, z1, c0, c1, ez0, ez1, uv0, uv1) (raster.c:107-241) */
__device__ __noinline__ void emit_line(Emitter &em, const mtgl_state *st, const FrameTargets fb, uint32_t state_index, int32_t x0, int32_t y0, float z0,
                          int32_t x1, int32_t y1, float z1, Color4 c0, Color4 c1, float ez0, float ez1, float u0, float v0, float u1, float v1)
{
    int lw = f2i_x86(st->line_width + 0.5f);
    if (lw < 1) lw = 1;
    const int half = lw / 2;
    const long long adx = llabs((long long)x1 - x0), ady = llabs((long long)y1 - y0);
    int minX = min(x0, x1), maxX = max(x0, x1), minY = min(y0, y1), maxY = max(y0, y1);
    if (lw > 1) {               /* perpendicular replication: vertical for mostly-horizontal lines (raster.c:136-146, 219-226) */
        if (adx > ady) { minY -= half; maxY += lw - half - 1; } else { minX -= half; maxX += lw - half - 1; }
    }
    if (!clamp_box(st, fb, minX, minY, maxX, maxY)) return;
    TriRecord rec;
    rec.x0 = x0; rec.y0 = y0; rec.x1 = x1; rec.y1 = y1; rec.x2 = lw; rec.y2 = 0;
    rec.area = 0.0f; rec.inv_area = 0.0f;
    rec.bbox_min = (uint32_t)minX | ((uint32_t)minY << 16);
    rec.bbox_max = (uint32_t)maxX | ((uint32_t)maxY << 16);
    rec.state_flags = state_index | (KIND_LINE << STATE_KIND_SHIFT);
    rec.id = 0;
    rec.z0 = z0; rec.z1 = z1; rec.z2 = 0.0f; rec.lod = 0.0f;
    rec.w0 = rec.w1 = rec.w2 = 0.0f; rec.ez0 = ez0;
    rec.c0[0] = c0.r; rec.c0[1] = c0.g; rec.c0[2] = c0.b; rec.c0[3] = c0.a;
    rec.c1[0] = c1.r; rec.c1[1] = c1.g; rec.c1[2] = c1.b; rec.c1[3] = c1.a;
    rec.c2[0] = rec.c2[1] = rec.c2[2] = rec.c2[3] = 0.0f;
    rec.u0 = u0; rec.v0 = v0; rec.u1 = u1; rec.v1 = v1; rec.u2 = 0.0f; rec.v2 = 0.0f;
    rec.ez1 = ez1; rec.ez2 = 0.0f;
    emit(em, rec, nullptr);
}

/* a POINT record: a ps x ps square with constant depth and colour (raster.c:1117-1162, 777-844) */
__device__ __noinline__ void emit_point(Emitter &em, const mtgl_state *st, const FrameTargets fb, uint32_t state_index, int32_t left, int32_t top, int ps,
                           float depth, Color4 c)
{
    long long r = (long long)left + ps - 1, bt = (long long)top + ps - 1;
    int minX = left, minY = top;
    int maxX = (int)min(r, (long long)INT_MAX), maxY = (int)min(bt, (long long)INT_MAX);
    if (!clamp_box(st, fb, minX, minY, maxX, maxY)) return;
    TriRecord rec;
    rec.x0 = left; rec.y0 = top; rec.x1 = ps; rec.y1 = 0; rec.x2 = 0; rec.y2 = 0;
    rec.area = 0.0f; rec.inv_area = 0.0f;
    rec.bbox_min = (uint32_t)minX | ((uint32_t)minY << 16);
    rec.bbox_max = (uint32_t)maxX | ((uint32_t)maxY << 16);
    rec.state_flags = state_index | (KIND_POINT << STATE_KIND_SHIFT);
    rec.id = 0;
    rec.z0 = depth; rec.z1 = rec.z2 = 0.0f; rec.lod = 0.0f;
    rec.w0 = rec.w1 = rec.w2 = 0.0f; rec.ez0 = 0.0f;
    rec.c0[0] = c.r; rec.c0[1] = c.g; rec.c0[2] = c.b; rec.c0[3] = c.a;
    rec.c1[0] = rec.c1[1] = rec.c1[2] = rec.c1[3] = 0.0f;
    rec.c2[0] = rec.c2[1] = rec.c2[2] = rec.c2[3] = 0.0f;
    rec.u0 = rec.v0 = rec.u1 = rec.v1 = rec.u2 = rec.v2 = 0.0f;
    rec.ez1 = rec.ez2 = 0.0f;
    emit(em, rec, nullptr);
}

__device__ __forceinline__ void to_screen(const mtgl_state *st, float x, float y, int32_t &sx, int32_t &sy)   /* raster.c:59-63 */
{
    sx = f2i_x86((x + 1.0f) * 0.5f * (float)st->viewport[2] + (float)st->viewport[0]);
    sy = f2i_x86((1.0f - y) * 0.5f * (float)st->viewport[3] + (float)st->viewport[1]);
}

/* polygon modes GL_LINE / GL_POINT: draw_triangle_wireframe / draw_triangle_points (raster.c:847-898) */
__device__ __noinline__ void setup_outline(Emitter &em, const mtgl_state *st, const RasterCfg *cfg, const FrameTargets fb, const SVert &a,
                                           const SVert &b, const SVert &c, const int32_t *sx, const int32_t *sy, uint32_t state_index, bool as_points)
{
    const SVert *v[3] = { &a, &b, &c };
    Color4 col[3];
    for (int k = 0; k < 3; k++) {
        col[k] = { v[k]->r, v[k]->g, v[k]->b, v[k]->a };
        if ((st->caps & MTGL_CAP_LIGHTING) && st->shade_model == G_PHONG) {
            MaterialRegs mat;
            load_material(mat, &st->material_front);
            col[k] = compute_lighting(st, v[k]->epx, v[k]->epy, v[k]->epz, v[k]->enx, v[k]->eny, v[k]->enz, mat);
        }
    }
    if (as_points) {
        for (int k = 0; k < 3; k++) {           /* draw_point_at_screen, raster.c:777-844: vertex alpha, no texture */
            if ((cfg->flags & RC_ALPHA_TEST) && !compare_f(cfg->alpha_func, col[k].a, cfg->alpha_ref)) continue;
            emit_point(em, st, fb, state_index, sx[k], sy[k], 1, simple_depth(cfg, v[k]->z), simple_fog(cfg, col[k], v[k]->ez));
        }
    } else {
        for (int k = 0; k < 3; k++) {
            const int n = (k + 1) % 3;
            emit_line(em, st, fb, state_index, sx[k], sy[k], v[k]->z, sx[n], sy[n], v[n]->z, col[k], col[n], v[k]->ez, v[n]->ez,
                      v[k]->u, v[k]->v, v[n]->u, v[n]->v);
        }
    }
}

/* ---------------------------------------------------------------- filled triangles */
struct ScreenTri {
    int32_t x0, y0, x1, y1, x2, y2;     /* snapped corners */
    float area;
    uint32_t bbox_min, bbox_max;        /* x | y << 16, inclusive, clamped */
    uint32_t mode;                      /* polygon mode of the facing side */
    bool back;
};
enum { TRI_REJECT = 0, TRI_FILL = 1, TRI_OUTLINE = 2 };

/* Everything render_triangle / rasterize_triangle_smooth decide from the three NDC positions alone
 * (raster.c:916-956, 458-499): snap, cull, facing, polygon mode, bounding box, degenerate-area reject. */
__device__ __forceinline__ int screen_setup(const mtgl_state *st, const FrameTargets &fb, float ax, float ay, float bx, float by, float cx, float cy,
                                            ScreenTri &s)
{
    to_screen(st, ax, ay, s.x0, s.y0);
    to_screen(st, bx, by, s.x1, s.y1);
    to_screen(st, cx, cy, s.x2, s.y2);
    const int32_t x0 = s.x0, y0 = s.y0, x1 = s.x1, y1 = s.y1, x2 = s.x2, y2 = s.y2;

    /* signed area of the snapped triangle decides culling and facing (raster.c:923-935) */
    float sa = (float)(x1 - x0) * (float)(y2 - y0) - (float)(x2 - x0) * (float)(y1 - y0);
    if (st->caps & MTGL_CAP_CULL_FACE) {
        bool front = (st->front_face == G_CCW) ? (sa < 0) : (sa > 0);
        bool cull = (st->cull_face_mode == G_FRONT) ? front : (st->cull_face_mode == G_BACK) ? !front : true;
        if (cull) return TRI_REJECT;
    }
    s.back = (st->front_face == G_CCW) ? (sa >= 0) : (sa < 0);
    s.mode = s.back ? st->polygon_mode_back : st->polygon_mode_front;
    if (s.mode == G_POINT || s.mode == G_LINE) return TRI_OUTLINE;

    int32_t minX = imin3(x0, x1, x2), minY = imin3(y0, y1, y2), maxX = imax3(x0, x1, x2), maxY = imax3(y0, y1, y2);
    const int32_t *vp = st->viewport;
    if (minX < vp[0]) minX = vp[0];
    if (minY < vp[1]) minY = vp[1];
    if (maxX >= vp[0] + vp[2]) maxX = vp[0] + vp[2] - 1;
    if (maxY >= vp[1] + vp[3]) maxY = vp[1] + vp[3] - 1;
    if (st->caps & MTGL_CAP_SCISSOR_TEST) {
        const int32_t *sc = st->scissor;
        if (minX < sc[0]) minX = sc[0];
        if (minY < sc[1]) minY = sc[1];
        if (maxX >= sc[0] + sc[2]) maxX = sc[0] + sc[2] - 1;
        if (maxY >= sc[1] + sc[3]) maxY = sc[1] + sc[3] - 1;
    }
    if (minX > maxX || minY > maxY) return TRI_REJECT;

    /* edge_function(x0,y0,x1,y1,x2,y2) (raster.c:483, 299-302) */
    s.area = ((float)x2 - (float)x0) * ((float)y1 - (float)y0) - ((float)y2 - (float)y0) * ((float)x1 - (float)x0);
    if (fabsf(s.area) < 0.5f) return TRI_REJECT;

    /* pixels outside the framebuffer are dropped by the bounds-checked accessors (framebuffer.h:92-134);
     * rows outside this device's band belong to another GPU */
    if (minX < 0) minX = 0;
    if (maxX >= fb.width) maxX = fb.width - 1;
    if (minY < fb.band_y0) minY = fb.band_y0;
    if (maxY >= fb.band_y1) maxY = fb.band_y1 - 1;
    if (minX > maxX || minY > maxY) return TRI_REJECT;
    s.bbox_min = (uint32_t)minX | ((uint32_t)minY << 16);
    s.bbox_max = (uint32_t)maxX | ((uint32_t)maxY << 16);
    return TRI_FILL;
}

/* the 160-byte record of a filled triangle (+ the eye-space side record for per-fragment lighting) */
__device__ __forceinline__ uint4 write_fill(TriRecord *dst, TriEye *eye_dst, uint32_t id, const RasterCfg *cfg, uint32_t state_index, const ScreenTri &s,
                                            const SVert &a, const SVert &b, const SVert &c)
{
    TriRecord rec;
    rec.x0 = s.x0; rec.y0 = s.y0; rec.x1 = s.x1; rec.y1 = s.y1; rec.x2 = s.x2; rec.y2 = s.y2;
    const uint32_t cflags = cfg->flags;
    rec.state_flags = state_index | (s.back ? STATE_BACK_BIT : 0u) | ((cflags & RC_DEFER) ? STATE_DEFER_BIT : 0u) |
                      ((cflags & RC_UNORDERED) ? STATE_UNORD_BIT : 0u) |
                      (attr_bounded(a.u, a.v, a.w, b.u, b.v, b.w, c.u, c.v, c.w) ? STATE_BOUNDED_BIT : 0u);
    rec.id = id;
    rec.bbox_min = s.bbox_min;
    rec.bbox_max = s.bbox_max;
    rec.z0 = a.z; rec.z1 = b.z; rec.z2 = c.z;
    rec.w0 = a.w; rec.w1 = b.w; rec.w2 = c.w;
    rec.area = s.area;
    rec.inv_area = 1.0f / s.area;
    rec.c0[0] = a.r; rec.c0[1] = a.g; rec.c0[2] = a.b; rec.c0[3] = a.a;
    rec.c1[0] = b.r; rec.c1[1] = b.g; rec.c1[2] = b.b; rec.c1[3] = b.a;
    rec.c2[0] = c.r; rec.c2[1] = c.g; rec.c2[2] = c.b; rec.c2[3] = c.a;
    rec.u0 = a.u; rec.v0 = a.v; rec.u1 = b.u; rec.v1 = b.v; rec.u2 = c.u; rec.v2 = c.v;
    rec.ez0 = a.ez; rec.ez1 = b.ez; rec.ez2 = c.ez;

    float lod = 0.0f;                   /* one LOD per triangle from non-perspective UV deltas (raster.c:505-529) */
    if (cflags & RC_TEXTURED) {
        float screen_area = fabsf(s.area) * 0.5f;
        float tw = (float)cfg->tex_w, th = (float)cfg->tex_h;
        float du1 = (b.u - a.u) * tw, dv1 = (b.v - a.v) * th;
        float du2 = (c.u - a.u) * tw, dv2 = (c.v - a.v) * th;
        float texel_area = fabsf(du1 * dv2 - du2 * dv1) * 0.5f;
        if (screen_area > 0.0f) {
            float tpp = texel_area / screen_area;
            if (tpp > 0.0f) {
                lod = log2f(tpp) * 0.5f;
                if (lod < 0.0f) lod = 0.0f;
            }
        }
    }
    rec.lod = lod;
    store_record(dst, rec);
    if (eye_dst) {
        TriEye eye;
        eye.ep0[0] = a.epx; eye.ep0[1] = a.epy; eye.ep0[2] = a.epz; eye.ep0[3] = 0.0f;
        eye.ep1[0] = b.epx; eye.ep1[1] = b.epy; eye.ep1[2] = b.epz; eye.ep1[3] = 0.0f;
        eye.ep2[0] = c.epx; eye.ep2[1] = c.epy; eye.ep2[2] = c.epz; eye.ep2[3] = 0.0f;
        eye.en0[0] = a.enx; eye.en0[1] = a.eny; eye.en0[2] = a.enz; eye.en0[3] = 0.0f;
        eye.en1[0] = b.enx; eye.en1[1] = b.eny; eye.en1[2] = b.enz; eye.en1[3] = 0.0f;
        eye.en2[0] = c.enx; eye.en2[1] = c.eny; eye.en2[2] = c.enz; eye.en2[3] = 0.0f;
        *eye_dst = eye;
    }
    return make_uint4(rec.bbox_min, rec.bbox_max, rec.state_flags, rec.id);
}

/* One fan sub-triangle after the divide, general form (used by the out-of-line path): fill it, or turn it into its
 * outline (3 LINE records) / corners (3 POINT records).  Mirrors raster.c:916-956, 847-898 and 458-529. */
__device__ void setup_subtri(Emitter &em, const mtgl_state *st, const RasterCfg *cfg, const FrameTargets &fb,
                             const SVert &a, const SVert &b, const SVert &c, uint32_t state_index)
{
    ScreenTri s;
    const int kind = screen_setup(st, fb, a.x, a.y, b.x, b.y, c.x, c.y, s);
    if (kind == TRI_REJECT) return;
    if (kind == TRI_OUTLINE) {
        const int32_t sx[3] = { s.x0, s.x1, s.x2 }, sy[3] = { s.y0, s.y1, s.y2 };
        setup_outline(em, st, cfg, fb, a, b, c, sx, sy, state_index, s.mode == G_POINT);
        return;
    }
    if (em.write) {
        const uint4 row = write_fill(em.dst + em.n, em.eye_dst ? em.eye_dst + em.n : nullptr, em.id0 + em.n, cfg, state_index, s, a, b, c);
        const uint32_t r = (uint32_t)(em.dst + em.n - em.bin->records);
        em.bin->bin_rows[r] = row;
        count_tiles_single(*em.bin, r, row.x, row.y, row.z);
    }
    em.n++;
}

/* ---------------------------------------------------------------- line segments and points as primitives */
__device__ __forceinline__ int outcode(const SVert &v)   /* clipping.h:138-148 */
{
    int c = 0;
    if (v.x < -v.w) c |= 1; else if (v.x > v.w) c |= 2;
    if (v.y < -v.w) c |= 4; else if (v.y > v.w) c |= 8;
    if (v.z < -v.w) c |= 16; else if (v.z > v.w) c |= 32;
    return c;
}

__device__ bool clip_segment(SVert &v0, SVert &v1)   /* clip_line, clipping.h:152-229 */
{
    int c0 = outcode(v0), c1 = outcode(v1);
    for (int guard = 0; guard < 64; guard++) {      /* (the reference loops unboundedly; it converges in <= 6 rounds per end) */
        if (!(c0 | c1)) return true;
        if (c0 & c1) return false;
        const int co = c0 ? c0 : c1;
        float d0, d1;
        int plane;
        if (co & 1) { d0 = v0.x + v0.w; d1 = v1.x + v1.w; plane = 2; }
        else if (co & 2) { d0 = v0.w - v0.x; d1 = v1.w - v1.x; plane = 3; }
        else if (co & 4) { d0 = v0.y + v0.w; d1 = v1.y + v1.w; plane = 4; }
        else if (co & 8) { d0 = v0.w - v0.y; d1 = v1.w - v1.y; plane = 5; }
        else if (co & 16) { d0 = v0.z + v0.w; d1 = v1.z + v1.w; plane = 0; }
        else { d0 = v0.w - v0.z; d1 = v1.w - v1.z; plane = 1; }
        const float den = d0 - d1;
        if (fabsf(den) < 1e-10f) return false;
        SVert cl = vertex_lerp(v0, v1, d0 / den);
        snap(cl, plane);
        if (co == c0) { v0 = cl; c0 = outcode(v0); } else { v1 = cl; c1 = outcode(v1); }
    }
    return false;
}

/* draw_line_segment (raster.c:244-285) up to the call of draw_line_full */
__device__ __noinline__ void setup_segment(Emitter &em, const mtgl_state *st, const FrameTargets fb, uint32_t state_index, SVert v0, SVert v1)
{
    if (!clip_segment(v0, v1)) return;
    float z0, z1;
    if (fabsf(v0.w) >= 1e-6f) { float iw = 1.0f / v0.w; v0.x *= iw; v0.y *= iw; z0 = v0.z * iw; } else { v0.x = 0.0f; v0.y = 0.0f; z0 = 0.0f; }
    if (fabsf(v1.w) >= 1e-6f) { float iw = 1.0f / v1.w; v1.x *= iw; v1.y *= iw; z1 = v1.z * iw; } else { v1.x = 0.0f; v1.y = 0.0f; z1 = 0.0f; }
    int32_t x0, y0, x1, y1;
    to_screen(st, v0.x, v0.y, x0, y0);
    to_screen(st, v1.x, v1.y, x1, y1);
    emit_line(em, st, fb, state_index, x0, y0, z0, x1, y1, z1, { v0.r, v0.g, v0.b, v0.a }, { v1.r, v1.g, v1.b, v1.a }, v0.ez, v1.ez,
              v0.u, v0.v, v1.u, v1.v);
}

/* one vertex of flush_points (raster.c:1044-1163): everything but the per-pixel tests happens here */
__device__ __noinline__ void setup_point(Emitter &em, const float *unorm8, const mtgl_state *st, const RasterCfg *cfg, const FrameTargets fb,
                            uint32_t state_index, SVert v)
{
    if (v.x < -v.w || v.x > v.w || v.y < -v.w || v.y > v.w || v.z < -v.w || v.z > v.w || v.w <= 0.0f) return;
    const float nx = v.x / v.w, ny = v.y / v.w, nz = v.z / v.w;
    int32_t cx, cy;
    to_screen(st, nx, ny, cx, cy);
    const float depth = simple_depth(cfg, nz);
    Color4 c = { v.r, v.g, v.b, v.a };
    if (cfg->flags & RC_TEXTURED) {
        TexTaps T;
        tex_taps(T, cfg, v.u, v.v, 0.0f);               /* texture_sample(): magnification filter */
        Color4 t;
        t.a = tex_channel(T, 24, unorm8);
        if ((cfg->flags & RC_ALPHA_TEST) && !compare_f(cfg->alpha_func, t.a, cfg->alpha_ref)) return;
        t.r = tex_channel(T, 0, unorm8); t.g = tex_channel(T, 8, unorm8); t.b = tex_channel(T, 16, unorm8);
        c = { c.r * t.r, c.g * t.g, c.b * t.b, c.a * t.a };
    } else if ((cfg->flags & RC_ALPHA_TEST) && !compare_f(cfg->alpha_func, c.a, cfg->alpha_ref)) return;
    c = simple_fog(cfg, c, v.ez);
    int ps = f2i_x86(st->point_size + 0.5f);
    if (ps < 1) ps = 1;
    emit_point(em, st, fb, state_index, cx - ps / 2, cy - ps / 2, ps, depth, c);
}

__device__ __forceinline__ uint32_t find_draw_tri(const uint32_t *base, uint32_t n, uint32_t g)
{
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(base + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

/* where a thread finds its vertices; passed by value to the out-of-line path so that the kernel's parameter
 * block never has its address taken (that would copy it to local memory at entry) */
struct VertexSrc {
    const float4 *clip, *color, *tex, *epos, *enrm;
    const float *unorm8;
    int need_eye;
    /* fused draws (independent triangles): vertices are shaded here instead of being read from the streams */
    const mtgl_in_vertex *staged;
    const mtgl_state *states;
    const DevDraw *draw;
    uint32_t fused;
};

/* the vertex stage for one vertex of a fused draw (dev_vertex.cuh), as the SVert the set-up code works on */
__device__ __noinline__ SVert compute_vertex(const mtgl_in_vertex *staged, const mtgl_state *states, const DevDraw *draw, uint32_t i)
{
    VertexIn in;
    fetch_vertex(staged, states, *draw, i - draw->vbase, in);
    VertexOut v;
    shade_vertex(in, v);
    SVert o;
    o.x = v.clip.x; o.y = v.clip.y; o.z = v.clip.z; o.w = v.clip.w;
    o.r = v.color.x; o.g = v.color.y; o.b = v.color.z; o.a = v.color.w;
    o.u = v.tex.x; o.v = v.tex.y; o.ez = v.tex.z;
    o.epx = v.epos.x; o.epy = v.epos.y; o.epz = v.epos.z; o.enx = v.enrm.x; o.eny = v.enrm.y; o.enz = v.enrm.z;
    return o;
}

__device__ __forceinline__ SVert load_vertex(const VertexSrc &b, uint32_t i)
{
    if (b.fused) {
        SVert o = compute_vertex(b.staged, b.states, b.draw, i);
        if (!b.need_eye) { o.epx = o.epy = o.epz = 0.0f; o.enx = o.eny = 0.0f; o.enz = 1.0f; }
        return o;
    }
    SVert o;
    float4 p = b.clip[i], c = b.color[i], t = b.tex[i];
    o.x = p.x; o.y = p.y; o.z = p.z; o.w = p.w;
    o.r = c.x; o.g = c.y; o.b = c.z; o.a = c.w;
    o.u = t.x; o.v = t.y; o.ez = t.z;
    if (b.need_eye) {
        float4 e = b.epos[i], n = b.enrm[i];
        o.epx = e.x; o.epy = e.y; o.epz = e.z; o.enx = n.x; o.eny = n.y; o.enz = n.z;
    } else { o.epx = o.epy = o.epz = 0.0f; o.enx = o.eny = 0.0f; o.enz = 1.0f; }
    return o;
}

/* write_fill for the common path, streaming: each group of record rows is loaded from the post-transform vertex
 * arrays and stored at once, so no more than three float4 are live at a time (the generic form keeps all three
 * vertices, 51 values, in registers).  Same values, same operations. */
__device__ __forceinline__ uint4 write_fill_stream(TriRecord *__restrict__ dst, TriEye *__restrict__ eye_dst, uint32_t id, const RasterCfg *cfg,
                                                   uint32_t state_index, const ScreenTri &s, const VertexSrc &src, uint32_t i0, uint32_t i1, uint32_t i2)
{
    float4 *out = reinterpret_cast<float4 *>(dst);
    const uint32_t cflags = cfg->flags;
    uint32_t state_flags = state_index | (s.back ? STATE_BACK_BIT : 0u) | ((cflags & RC_DEFER) ? STATE_DEFER_BIT : 0u) |
                           ((cflags & RC_UNORDERED) ? STATE_UNORD_BIT : 0u);
    reinterpret_cast<int4 *>(dst)[0] = make_int4(s.x0, s.y0, s.x1, s.y1);
    reinterpret_cast<int4 *>(dst)[1] = make_int4(s.x2, s.y2, __float_as_int(s.area), __float_as_int(1.0f / s.area));
    float bu[3], bv[3];
    {   /* rows 5-7: colours are copied */
        const float4 c0 = __ldg(src.color + i0), c1 = __ldg(src.color + i1), c2 = __ldg(src.color + i2);
        out[5] = c0; out[6] = c1; out[7] = c2;
    }
    float lod = 0.0f, ez0;
    {   /* rows 8-9: texture coordinates and eye z; one LOD per triangle from non-perspective UV deltas (raster.c:505-529) */
        const float4 t0 = __ldg(src.tex + i0), t1 = __ldg(src.tex + i1), t2 = __ldg(src.tex + i2);
        if (cflags & RC_TEXTURED) {
            float screen_area = fabsf(s.area) * 0.5f;
            float tw = (float)cfg->tex_w, th = (float)cfg->tex_h;
            float du1 = (t1.x - t0.x) * tw, dv1 = (t1.y - t0.y) * th;
            float du2 = (t2.x - t0.x) * tw, dv2 = (t2.y - t0.y) * th;
            float texel_area = fabsf(du1 * dv2 - du2 * dv1) * 0.5f;
            if (screen_area > 0.0f) {
                float tpp = texel_area / screen_area;
                if (tpp > 0.0f) {
                    lod = log2f(tpp) * 0.5f;
                    if (lod < 0.0f) lod = 0.0f;
                }
            }
        }
        out[8] = make_float4(t0.x, t0.y, t1.x, t1.y);
        out[9] = make_float4(t2.x, t2.y, t1.z, t2.z);
        ez0 = t0.z;
        bu[0] = t0.x; bu[1] = t1.x; bu[2] = t2.x; bv[0] = t0.y; bv[1] = t1.y; bv[2] = t2.y;
    }
    {   /* rows 3-4: z / w after the divide (raster.c:729-746) */
        const float4 p0 = __ldg(src.clip + i0), p1 = __ldg(src.clip + i1), p2 = __ldg(src.clip + i2);
        float z0, w0, z1, w1, z2, w2;
        if (fabsf(p0.w) < 1e-6f) { z0 = 0.0f; w0 = 1.0f; } else { w0 = 1.0f / p0.w; z0 = p0.z * w0; }
        if (fabsf(p1.w) < 1e-6f) { z1 = 0.0f; w1 = 1.0f; } else { w1 = 1.0f / p1.w; z1 = p1.z * w1; }
        if (fabsf(p2.w) < 1e-6f) { z2 = 0.0f; w2 = 1.0f; } else { w2 = 1.0f / p2.w; z2 = p2.z * w2; }
        out[3] = make_float4(z0, z1, z2, lod);
        out[4] = make_float4(w0, w1, w2, ez0);
        if (attr_bounded(bu[0], bv[0], w0, bu[1], bv[1], w1, bu[2], bv[2], w2)) state_flags |= STATE_BOUNDED_BIT;
    }
    reinterpret_cast<uint4 *>(dst)[2] = make_uint4(s.bbox_min, s.bbox_max, state_flags, id);
    if (eye_dst) {
        float4 *eo = reinterpret_cast<float4 *>(eye_dst);
        float4 e0 = __ldg(src.epos + i0), e1 = __ldg(src.epos + i1), e2 = __ldg(src.epos + i2);
        e0.w = 0.0f; e1.w = 0.0f; e2.w = 0.0f;
        eo[0] = e0; eo[1] = e1; eo[2] = e2;
        float4 n0 = __ldg(src.enrm + i0), n1 = __ldg(src.enrm + i1), n2 = __ldg(src.enrm + i2);
        n0.w = 0.0f; n1.w = 0.0f; n2.w = 0.0f;
        eo[3] = n0; eo[4] = n1; eo[5] = n2;
    }
    return make_uint4(s.bbox_min, s.bbox_max, state_flags, id);
}

/* Fused path, triangle part: rows 0-4 and 8-9 of the record (geometry, z / w, texture coordinates, eye z, LOD).  z, w
 * (after the divide, raster.c:729-746) and -eye z come from the decision, which already transformed the three positions;
 * the colour rows (and the eye-space side record) are filled afterwards by the whole CTA (k_setup, phase B). */
__device__ __forceinline__ uint4 write_fused_head(TriRecord *__restrict__ dst, uint32_t id, const RasterCfg *cfg, uint32_t state_index,
                                                  const ScreenTri &s, const mtgl_state *vs, const float (&z)[3], const float (&w)[3],
                                                  const float (&nez)[3], const float (&ts)[3], const float (&tt)[3])
{
    float4 *out = reinterpret_cast<float4 *>(dst);
    const uint32_t cflags = cfg->flags;
    uint32_t state_flags = state_index | (s.back ? STATE_BACK_BIT : 0u) | ((cflags & RC_DEFER) ? STATE_DEFER_BIT : 0u) |
                           ((cflags & RC_UNORDERED) ? STATE_UNORD_BIT : 0u);
    reinterpret_cast<int4 *>(dst)[0] = make_int4(s.x0, s.y0, s.x1, s.y1);
    reinterpret_cast<int4 *>(dst)[1] = make_int4(s.x2, s.y2, __float_as_int(s.area), __float_as_int(1.0f / s.area));
    float tu[3], tv[3];
#pragma unroll
    for (int j = 0; j < 3; j++) tex_transform(vs, ts[j], tt[j], tu[j], tv[j]);
    float lod = 0.0f;       /* one LOD per triangle from non-perspective UV deltas (raster.c:505-529) */
    if (cflags & RC_TEXTURED) {
        float screen_area = fabsf(s.area) * 0.5f;
        float tw = (float)cfg->tex_w, th = (float)cfg->tex_h;
        float du1 = (tu[1] - tu[0]) * tw, dv1 = (tv[1] - tv[0]) * th;
        float du2 = (tu[2] - tu[0]) * tw, dv2 = (tv[2] - tv[0]) * th;
        float texel_area = fabsf(du1 * dv2 - du2 * dv1) * 0.5f;
        if (screen_area > 0.0f) {
            float tpp = texel_area / screen_area;
            if (tpp > 0.0f) {
                lod = log2f(tpp) * 0.5f;
                if (lod < 0.0f) lod = 0.0f;
            }
        }
    }
    if (attr_bounded(tu[0], tv[0], w[0], tu[1], tv[1], w[1], tu[2], tv[2], w[2])) state_flags |= STATE_BOUNDED_BIT;
    reinterpret_cast<uint4 *>(dst)[2] = make_uint4(s.bbox_min, s.bbox_max, state_flags, id);
    out[3] = make_float4(z[0], z[1], z[2], lod);
    out[4] = make_float4(w[0], w[1], w[2], nez[0]);
    out[8] = make_float4(tu[0], tv[0], tu[1], tv[1]);
    out[9] = make_float4(tu[2], tv[2], nez[1], nez[2]);
    return make_uint4(s.bbox_min, s.bbox_max, state_flags, id);
}

/* Everything that is not an unclipped filled triangle: clipped polygons (a fan of up to 7 sub-triangles), outlines,
 * line segments and points.  Out of line and self-contained (it reloads its vertices) so that the common path
 * stays in registers.  Returns the number of records emitted. */
__device__ __noinline__ uint32_t setup_rare(const bool write, TriRecord *const dst, TriEye *const eye_dst, const uint32_t id0, const BinOut bin, const VertexSrc src,
                                            const mtgl_state *st, const RasterCfg *cfg, const FrameTargets fb, const uint32_t state_index,
                                            const int shape, const uint32_t i0, const uint32_t i1, const uint32_t i2)
{
    Emitter em = { write, 0u, dst, eye_dst, id0, &bin };
    if (shape == 4) {
        setup_point(em, src.unorm8, st, cfg, fb, state_index, load_vertex(src, i0));
    } else if (shape == 3) {
        setup_segment(em, st, fb, state_index, load_vertex(src, i0), load_vertex(src, i1));
    } else if (shape == 5) {        /* unclipped triangle with polygon mode GL_LINE / GL_POINT */
        SVert a = load_vertex(src, i0), b = load_vertex(src, i1), c = load_vertex(src, i2);
        persp_divide(a); persp_divide(b); persp_divide(c);
        setup_subtri(em, st, cfg, fb, a, b, c, state_index);
    } else {
        SVert poly_a[MAX_CLIP], poly_b[MAX_CLIP];
        poly_a[0] = load_vertex(src, i0); poly_a[1] = load_vertex(src, i1); poly_a[2] = load_vertex(src, i2);
        int m = clip_plane(poly_a, 3, poly_b, 0);       /* clipping.h:106-126 */
        if (m) m = clip_plane(poly_b, m, poly_a, 1);
        if (m) m = clip_plane(poly_a, m, poly_b, 2);
        if (m) m = clip_plane(poly_b, m, poly_a, 3);
        if (m) m = clip_plane(poly_a, m, poly_b, 4);
        if (m) m = clip_plane(poly_b, m, poly_a, 5);
        if (m < 3) return 0u;
        for (int j = 0; j < m; j++) persp_divide(poly_a[j]);
        for (int j = 1; j + 1 < m; j++) setup_subtri(em, st, cfg, fb, poly_a[0], poly_a[j], poly_a[j + 1], state_index);
    }
    return em.n;
}

__device__ __forceinline__ bool inside_all(const float4 &v)
{
    return (v.z + v.w) >= 0 && (v.w - v.z) >= 0 && (v.x + v.w) >= 0 && (v.w - v.x) >= 0 && (v.y + v.w) >= 0 && (v.w - v.y) >= 0;
}

/* persp_divide (raster.c:729-746) on a clip-space position: x, y, z over w, and w becomes 1/w */
__device__ __forceinline__ void persp_divide4(const float4 &v, float &x, float &y, float &z, float &w)
{
    if (fabsf(v.w) < 1e-6f) { x = 0.0f; y = 0.0f; z = 0.0f; w = 1.0f; return; }
    w = 1.0f / v.w;
    x = v.x * w; y = v.y * w; z = v.z * w;
}

/*
 * The common case -- a triangle with all three vertices inside the frustum and polygon mode GL_FILL -- is decided
 * from the three clip-space positions alone (48 B); colours, texture coordinates and eye-space attributes are only
 * fetched for survivors, so culled triangles cost a third of the traffic.
 *
 * Fused draws on the fast attribute path (FastDraw) stage the chunk's 768 raw vertices (position, normal, texture
 * coordinate) in shared memory once, structure-of-arrays, every warp the 96 vertices of its own 32 triangles with all
 * loads in flight at the same time; the decision, the record heads, the vertex-cache hashing and the vertex stage all
 * read them from there, so each CTA pays one global-memory round trip for its attributes instead of one per phase.
 *
 * Record slots are reserved per WARP: a shuffle scan of the survivor counts and one atomicAdd per warp.  Storage order
 * therefore is not submission order -- the record's id is: id = group << GROUP_SHIFT | index-in-group, the group being
 * the warp's 32 consecutive input triangles, is ordered exactly like the reference's sequential loop, the tile kernels
 * sort and compare by it, and K4a turns an id back into a slot through group_base[].  Up to the vertex stage the CTA
 * meets at ONE barrier (the prologue's); the earlier form -- block-wide scan, dense slots per chunk -- had four, and
 * evaluated the three transforms twice because nothing but the screen-space result could live across them.
 */
constexpr uint32_t SETUP_VERTS = 3 * SETUP_THREADS;
constexpr uint32_t VCACHE_SLOTS = 2048;
static_assert(32 * 21 <= (1 << GROUP_SHIFT), "a group's records (32 primitives, up to 21 records each) fit its id field");

struct SetupSmem {
    /* raw attributes of local vertex lv = 3 * (triangle - first triangle of the chunk) + corner; after phase B2 the first
     * four arrays of an OWNER vertex hold its shaded colour (r, g, b, a) instead */
    float a_px[SETUP_VERTS], a_py[SETUP_VERTS], a_pz[SETUP_VERTS], a_nx[SETUP_VERTS];
    float a_ny[SETUP_VERTS], a_nz[SETUP_VERTS], a_s[SETUP_VERTS], a_t[SETUP_VERTS];
    uint32_t vcache[VCACHE_SLOTS];              /* phase B: hash slot -> first local vertex that claimed it */
    uint32_t rec_of[SETUP_THREADS];             /* record of the thread's fused survivor (eye-space rows of duplicates, need_eye) */
    uint16_t vsame[SETUP_VERTS];                /* local vertex -> the local vertex it duplicates (itself if none) */
    uint16_t shade_list[SETUP_VERTS];           /* local vertices to shade */
    uint32_t shade_n;
    FastDraw fastd;                             /* fast attribute path of the draw this chunk starts in */
    mtgl_state vstate;                          /* ... and a copy of its vertex-stage state block: matrices, lights and materials
                                                 * are read from shared memory, not through L1, by every thread and every light */
};
static_assert(sizeof(SetupSmem) <= 48 * 1024, "k_setup uses static shared memory");

__global__ void __launch_bounds__(SETUP_THREADS, 4) k_setup(BatchDev b, FrameTargets fb)
{
    __shared__ SetupSmem sm;
    FastDraw &fastd = sm.fastd;
    const uint32_t chunk = blockIdx.x;
    if (b.chunk_cull && b.chunk_cull[chunk]) return;        /* k_cull.cu: nothing of this chunk reaches the device's rows */
    const uint32_t t0 = chunk * SETUP_THREADS;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const mtgl_state *const fst = &sm.vstate;
    if (threadIdx.x == 0) sm.shade_n = 0;
    /* the fast attribute path of the draw this chunk starts in: decided per draw on the host; the descriptor and the
     * vertex-stage state block it names are copied to shared memory */
    bool any_fast = false;
    {
        const uint32_t d0 = (b.n_draws == 1) ? 0u : find_draw_tri(b.draw_tbase, b.n_draws, t0);
        const FastDraw *const gf = &b.draws[d0].fast;
        any_fast = __ldg(&gf->valid) != 0u;
        if (threadIdx.x < sizeof(FastDraw) / 4u)
            reinterpret_cast<uint32_t *>(&fastd)[threadIdx.x] = __ldg(reinterpret_cast<const uint32_t *>(gf) + threadIdx.x);
        if (any_fast) {
            const uint32_t *srcw = reinterpret_cast<const uint32_t *>(
                (const void *)__ldg(reinterpret_cast<const unsigned long long *>(&gf->st)));
            uint32_t *dstw = reinterpret_cast<uint32_t *>(&sm.vstate);
            for (uint32_t i = threadIdx.x; i < sizeof(mtgl_state) / 4u; i += SETUP_THREADS) dstw[i] = __ldg(srcw + i);
        }
    }
    for (uint32_t i = threadIdx.x; i < VCACHE_SLOTS; i += SETUP_THREADS) sm.vcache[i] = 0xFFFFFFFFu;
    __syncthreads();

    /* ---- phase A0: every warp stages the raw attributes of its own fast-path triangles ---- */
    if (any_fast) {
        const uint32_t tb = max(fastd.tri_begin, t0), te = min(min(fastd.tri_end, t0 + SETUP_THREADS), b.n_triangles);
        const uint32_t lo = max(96u * warp, 3u * (tb - t0)), hi = min(96u * warp + 96u, 3u * (te - t0));
        for (uint32_t lv = lo + lane; lv < hi; lv += 32u) {
            VertexIn in;
            fast_vertex(fastd, 3u * (t0 - fastd.tbase) + lv, in);
            sm.a_px[lv] = in.px; sm.a_py[lv] = in.py; sm.a_pz[lv] = in.pz;
            sm.a_nx[lv] = in.nx; sm.a_ny[lv] = in.ny; sm.a_nz[lv] = in.nz;
            sm.a_s[lv] = in.s; sm.a_t[lv] = in.t;
        }
        __syncwarp();
    }

    const uint32_t t = t0 + threadIdx.x;
    const bool valid = t < b.n_triangles;
    const bool fast_t = any_fast && t >= fastd.tri_begin && t < fastd.tri_end;     /* this thread's triangle is staged */

    int shape = 0;          /* 0 nothing, 1 unclipped filled triangle, 2 clipped polygon, 3 line segment, 4 point, 5 unclipped outline */
    const mtgl_state *st = nullptr;
    const RasterCfg *cfg = nullptr;
    uint32_t state_index = 0;
    uint32_t i0 = 0, i1 = 0, i2 = 0;
    uint32_t count = 0;
    ScreenTri s;
    float zd[3] = { 0.0f, 0.0f, 0.0f }, wd[3] = { 1.0f, 1.0f, 1.0f }, nez[3] = { 0.0f, 0.0f, 0.0f };   /* fused: z / w after the divide, -eye z */
    const mtgl_state *vs = nullptr;                         /* fused: the vertex-stage state of the triangle's vertices */
    VertexSrc src = { b.v_clip, b.v_color, b.v_tex, b.v_epos, b.v_enrm, b.unorm8, b.need_eye ? 1 : 0, b.staged, b.states, nullptr, 0u };
    const BinOut bin = { b.records, b.bin_rows, b.tile_count, b.tile_flags, b.large_list, b.counters, fb.tiles_x, fb.tile_y0 };

    uint32_t d = 0;
    if (valid) {
        d = (b.n_draws == 1) ? 0u : (fast_t ? fastd.draw : find_draw_tri(b.draw_tbase, b.n_draws, t));
        /* large draws start on a chunk boundary (mtgl_dev.cu): the slots between a draw's last triangle and the next
         * boundary hold nothing */
        if (t - __ldg(b.draw_tbase + d) >= b.draws[d].ntris) d = 0xFFFFFFFFu;
    }
    if (d != 0xFFFFFFFFu && valid) {
        const DevDraw &dr = b.draws[d];
        const uint32_t k = t - dr.tbase, n = dr.count;
        state_index = dr.raster_state;
        st = b.states + state_index;
        cfg = b.cfgs + state_index;
        src.draw = &dr;
        src.fused = dr.fused;
        switch (dr.mode) {                                  /* raster.c:961-1017, 1020-1044, 288-296, 1167-1231 */
        case G_POINTS: i0 = k; shape = 4; break;
        case G_LINES: i0 = 2 * k; i1 = i0 + 1; shape = 3; break;
        case G_LINE_STRIP: i0 = k; i1 = k + 1; shape = 3; break;
        case G_LINE_LOOP: if (k + 1 < n) { i0 = k; i1 = k + 1; } else { i0 = n - 1; i1 = 0; } shape = 3; break;
        case G_TRIANGLES: i0 = 3 * k; i1 = i0 + 1; i2 = i0 + 2; shape = 1; break;
        case G_QUADS: { uint32_t q = 4 * (k >> 1); i0 = q; if (k & 1) { i1 = q + 2; i2 = q + 3; } else { i1 = q + 1; i2 = q + 2; } shape = 1; break; }
        case G_TRIANGLE_STRIP: if (k & 1) { i0 = k + 1; i1 = k; } else { i0 = k; i1 = k + 1; } i2 = k + 2; shape = 1; break;
        case G_QUAD_STRIP: { uint32_t q = 2 * (k >> 1); i0 = q; if (k & 1) { i1 = q + 3; i2 = q + 2; } else { i1 = q + 1; i2 = q + 3; } shape = 1; break; }
        default: i0 = 0; i1 = k + 1; i2 = k + 2; shape = 1; break;     /* fan, polygon */
        }
        if (dr.shared_verts) {          /* the vertex stage ran once per buffer element: look the three up by index (gl_api.c:1898-1939) */
            i0 = shared_slot(dr, element_index(dr, i0)); i1 = shared_slot(dr, element_index(dr, i1)); i2 = shared_slot(dr, element_index(dr, i2));
        }
        i0 += dr.vbase; i1 += dr.vbase; i2 += dr.vbase;
        if (shape == 1) {
            float4 p0, p1, p2;
            if (src.fused) {            /* positions straight from the attribute arrays: MV, then P */
                const uint32_t l0 = i0 - dr.vbase;
                float x, y, z, ex, ey, ez, ew;
                if (fast_t) {
                    vs = fst;
                    const uint32_t lv = 3u * threadIdx.x;
                    to_eye(vs, sm.a_px[lv], sm.a_py[lv], sm.a_pz[lv], ex, ey, ez, ew); p0 = to_clip(vs, ex, ey, ez, ew); nez[0] = -ez;
                    to_eye(vs, sm.a_px[lv + 1], sm.a_py[lv + 1], sm.a_pz[lv + 1], ex, ey, ez, ew); p1 = to_clip(vs, ex, ey, ez, ew); nez[1] = -ez;
                    to_eye(vs, sm.a_px[lv + 2], sm.a_py[lv + 2], sm.a_pz[lv + 2], ex, ey, ez, ew); p2 = to_clip(vs, ex, ey, ez, ew); nez[2] = -ez;
                } else {
                    fetch_position(b.staged, b.states, dr, l0, x, y, z, vs); to_eye(vs, x, y, z, ex, ey, ez, ew); p0 = to_clip(vs, ex, ey, ez, ew); nez[0] = -ez;
                    fetch_position(b.staged, b.states, dr, l0 + 1, x, y, z, vs); to_eye(vs, x, y, z, ex, ey, ez, ew); p1 = to_clip(vs, ex, ey, ez, ew); nez[1] = -ez;
                    fetch_position(b.staged, b.states, dr, l0 + 2, x, y, z, vs); to_eye(vs, x, y, z, ex, ey, ez, ew); p2 = to_clip(vs, ex, ey, ez, ew); nez[2] = -ez;
                }
            } else { p0 = src.clip[i0]; p1 = src.clip[i1]; p2 = src.clip[i2]; }
            if (inside_all(p0) && inside_all(p1) && inside_all(p2)) {
                /* Sutherland-Hodgman returns its input unchanged when every vertex passes every plane */
                float ax, ay, bx, by, cx, cy;
                persp_divide4(p0, ax, ay, zd[0], wd[0]); persp_divide4(p1, bx, by, zd[1], wd[1]); persp_divide4(p2, cx, cy, zd[2], wd[2]);
                const int kind = screen_setup(st, fb, ax, ay, bx, by, cx, cy, s);
                if (kind == TRI_OUTLINE) shape = 5; else count = (uint32_t)kind;
            } else shape = 2;
        }
        if (shape > 1) count = setup_rare(false, nullptr, nullptr, 0u, bin, src, st, cfg, fb, state_index, shape, i0, i1, i2);
    }

    /* ---- slots: exclusive scan of the warp's survivor counts, one atomicAdd per warp ---- */
    uint32_t incl = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += n;
    }
    uint32_t group_slot0 = 0u;
    const uint32_t group = (t0 >> 5) + warp;
    if (lane == 31) {
        group_slot0 = incl ? atomicAdd(&b.counters->records, incl) : 0u;
        if (group_slot0 + incl > b.record_capacity) { atomicExch(&b.counters->overflow, 1u); group_slot0 = 0xFFFFFFFFu; }
        b.group_base[group] = group_slot0;
    }
    group_slot0 = __shfl_sync(0xFFFFFFFFu, group_slot0, 31);

    /* no early exit: the whole warp stays for the aggregated tile counting below */
    bool counted = false, fused_mine = false;
    uint32_t r = 0;
    uint4 row = make_uint4(0u, 0u, 0u, 0u);
    if (count != 0 && group_slot0 != 0xFFFFFFFFu) {
        const uint32_t slot = incl - count;                 /* index inside the group */
        r = group_slot0 + slot;
        TriRecord *const dst = b.records + r;
        TriEye *const eye_dst = b.need_eye ? b.rec_eye + r : nullptr;
        const uint32_t id0 = (group << GROUP_SHIFT) | slot;
        if (shape == 1) {
            if (src.fused) {
                float ts[3], tt[3];
                if (fast_t) {
#pragma unroll
                    for (int j = 0; j < 3; j++) { ts[j] = sm.a_s[3u * threadIdx.x + j]; tt[j] = sm.a_t[3u * threadIdx.x + j]; }
                } else {
                    const uint32_t l0 = i0 - src.draw->vbase;
#pragma unroll
                    for (int j = 0; j < 3; j++) fetch_texcoord(b.staged, *src.draw, l0 + j, ts[j], tt[j]);
                }
                row = write_fused_head(dst, id0, cfg, state_index, s, vs, zd, wd, nez, ts, tt);
                sm.rec_of[threadIdx.x] = r;
                fused_mine = true;
            } else row = write_fill_stream(dst, eye_dst, id0, cfg, state_index, s, src, i0, i1, i2);
            b.bin_rows[r] = row;
            counted = true;
        } else {
            setup_rare(true, dst, eye_dst, id0, bin, src, st, cfg, fb, state_index, shape, i0, i1, i2);   /* counts its own records */
        }
    }

    /* ---- fused first pass of the binner (k_bin.cu): references per tile.  Mesh-ordered triangles of one warp mostly
     * fall into the same tile, so single-tile records are counted with one atomic per distinct tile. ---- */
    int tx0 = 0, ty0 = 0, ntiles = 0;
    if (counted) {
        tx0 = (int)(row.x & 0xFFFFu) >> TILE_LOG; ty0 = ((int)(row.x >> 16) >> TILE_LOG) - fb.tile_y0;
        const int tx1 = (int)(row.y & 0xFFFFu) >> TILE_LOG, ty1 = ((int)(row.y >> 16) >> TILE_LOG) - fb.tile_y0;
        ntiles = (tx1 - tx0 + 1) * (ty1 - ty0 + 1);
    }
    const uint32_t single = __ballot_sync(0xFFFFFFFFu, ntiles == 1);
    if (ntiles == 1) {
        const uint32_t tile = (uint32_t)(ty0 * fb.tiles_x + tx0);
        const uint32_t tflags = tile_flag_bits(row.z);
        const uint32_t peers = __match_any_sync(single, tile);
        if ((int)lane == __ffs(peers) - 1) atomicAdd(&b.tile_count[tile], (uint32_t)__popc(peers));
        if (tflags) atomicOr(&b.tile_flags[tile], tflags);
    } else if (ntiles > 1) count_tiles_single(bin, r, row.x, row.y, row.z);

    /* ---- phase B of the fused path: the vertex stage for the survivors only -- attribute fetch, eye-space transform,
     * lighting (dev_vertex.cuh) -> colour row 5 + j of the record (and the eye-space side record).
     *
     * Independent triangles of a mesh repeat their vertices (C4: 4.3 references per distinct vertex inside a chunk of
     * 256 triangles).  The vertex stage is a pure function of (attributes, state block), so it runs once per DISTINCT
     * input of the chunk -- a post-transform vertex cache: B1 every survivor hashes the raw attributes of its three
     * vertices and claims a table slot for each; a loser compares bit for bit with the slot's owner and, if equal,
     * becomes its duplicate; B2 the owners (and hash collisions) are shaded, one thread each, and leave their colour in
     * shared memory; B3 every survivor stores its vertices' owners' colours into its record rows. ---- */
    if (!__syncthreads_or(fused_mine ? 1 : 0)) return;
    /* every triangle of the chunk is staged: its vertices share the state block and the current colour */
    const bool chunk_fast = any_fast && fastd.tri_begin <= t0 && fastd.tri_end >= min(t0 + SETUP_THREADS, b.n_triangles);
    auto vertex_in = [&](uint32_t lv, VertexIn &in, uint32_t &state_id) {
        const uint32_t tt = t0 + lv / 3u;
        if (any_fast && tt >= fastd.tri_begin && tt < fastd.tri_end) {
            in.px = sm.a_px[lv]; in.py = sm.a_py[lv]; in.pz = sm.a_pz[lv];
            in.nx = sm.a_nx[lv]; in.ny = sm.a_ny[lv]; in.nz = sm.a_nz[lv];
            in.s = sm.a_s[lv]; in.t = sm.a_t[lv];
            in.cur = { fastd.cur_color[0], fastd.cur_color[1], fastd.cur_color[2], fastd.cur_color[3] };
            in.st = fst;
            state_id = (uint32_t)(fastd.st - b.states);
            return;
        }
        const uint32_t dd = (b.n_draws == 1) ? 0u : find_draw_tri(b.draw_tbase, b.n_draws, tt);
        const DevDraw &dr = b.draws[dd];
        fetch_vertex(b.staged, b.states, dr, 3u * (tt - dr.tbase) + lv % 3u, in);
        state_id = (uint32_t)(in.st - b.states);
    };
    if (fused_mine) {                                                       /* B1 */
        for (uint32_t j = 0; j < 3u; j++) {
            const uint32_t lv = 3u * threadIdx.x + j;
            uint32_t w[13];
            if (chunk_fast) {
                w[0] = __float_as_uint(sm.a_px[lv]); w[1] = __float_as_uint(sm.a_py[lv]); w[2] = __float_as_uint(sm.a_pz[lv]);
                w[3] = __float_as_uint(sm.a_nx[lv]); w[4] = __float_as_uint(sm.a_ny[lv]); w[5] = __float_as_uint(sm.a_nz[lv]);
                w[6] = __float_as_uint(sm.a_s[lv]); w[7] = __float_as_uint(sm.a_t[lv]);
                w[8] = w[9] = w[10] = w[11] = w[12] = 0u;
            } else {
                VertexIn in;
                uint32_t sid;
                vertex_in(lv, in, sid);
                w[0] = __float_as_uint(in.px); w[1] = __float_as_uint(in.py); w[2] = __float_as_uint(in.pz); w[3] = __float_as_uint(in.nx);
                w[4] = __float_as_uint(in.ny); w[5] = __float_as_uint(in.nz); w[6] = __float_as_uint(in.s); w[7] = __float_as_uint(in.t);
                w[8] = __float_as_uint(in.cur.r); w[9] = __float_as_uint(in.cur.g); w[10] = __float_as_uint(in.cur.b);
                w[11] = __float_as_uint(in.cur.a); w[12] = sid;
            }
            uint32_t hsh = 0x811C9DC5u;
#pragma unroll
            for (int k = 0; k < 8; k++) hsh = (hsh ^ w[k]) * 0x9E3779B1u;
            if (!chunk_fast) {
#pragma unroll
                for (int k = 8; k < 13; k++) hsh = (hsh ^ w[k]) * 0x9E3779B1u;
            }
            hsh ^= hsh >> 15;
            const uint32_t owner = atomicCAS(&sm.vcache[hsh & (VCACHE_SLOTS - 1)], 0xFFFFFFFFu, lv);
            uint32_t same_as = lv;
            if (owner != 0xFFFFFFFFu) {
                bool equal;
                if (chunk_fast) {
                    equal = __float_as_uint(sm.a_px[owner]) == w[0] && __float_as_uint(sm.a_py[owner]) == w[1] &&
                            __float_as_uint(sm.a_pz[owner]) == w[2] && __float_as_uint(sm.a_nx[owner]) == w[3] &&
                            __float_as_uint(sm.a_ny[owner]) == w[4] && __float_as_uint(sm.a_nz[owner]) == w[5] &&
                            __float_as_uint(sm.a_s[owner]) == w[6] && __float_as_uint(sm.a_t[owner]) == w[7];
                } else {
                    VertexIn o;
                    uint32_t osid;
                    vertex_in(owner, o, osid);
                    equal = __float_as_uint(o.px) == w[0] && __float_as_uint(o.py) == w[1] && __float_as_uint(o.pz) == w[2] &&
                            __float_as_uint(o.nx) == w[3] && __float_as_uint(o.ny) == w[4] && __float_as_uint(o.nz) == w[5] &&
                            __float_as_uint(o.s) == w[6] && __float_as_uint(o.t) == w[7] && __float_as_uint(o.cur.r) == w[8] &&
                            __float_as_uint(o.cur.g) == w[9] && __float_as_uint(o.cur.b) == w[10] && __float_as_uint(o.cur.a) == w[11] &&
                            osid == w[12];
                }
                if (equal) same_as = owner;
            }
            sm.vsame[lv] = (uint16_t)same_as;
            if (same_as == lv) sm.shade_list[atomicAdd(&sm.shade_n, 1u)] = (uint16_t)lv;
        }
    }
    __syncthreads();
    const uint32_t ns = sm.shade_n;
    for (uint32_t i = threadIdx.x; i < ns; i += SETUP_THREADS) {            /* B2 */
        const uint32_t lv = sm.shade_list[i];
        VertexIn in;
        uint32_t sid;
        vertex_in(lv, in, sid);
        VertexOut o;
        shade_vertex(in, o);
        /* the owner's raw position / normal.x are dead from here on (only this thread read them): its colour takes their place */
        sm.a_px[lv] = o.color.x; sm.a_py[lv] = o.color.y; sm.a_pz[lv] = o.color.z; sm.a_nx[lv] = o.color.w;
        if (b.need_eye) {
            float4 *eo = reinterpret_cast<float4 *>(b.rec_eye + sm.rec_of[lv / 3u]);
            eo[lv % 3u] = o.epos;
            eo[3 + lv % 3u] = o.enrm;
        }
    }
    __syncthreads();                                        /* the owners' colours are visible to the whole CTA */
    if (fused_mine) {                                                       /* B3 */
#pragma unroll
        for (uint32_t j = 0; j < 3u; j++) {
            const uint32_t lv = 3u * threadIdx.x + j;
            const uint32_t o = sm.vsame[lv];
            reinterpret_cast<float4 *>(b.records + r)[5 + j] = make_float4(sm.a_px[o], sm.a_py[o], sm.a_pz[o], sm.a_nx[o]);
            if (b.need_eye && o != lv) {
                const uint32_t ro = sm.rec_of[o / 3u], jo = o % 3u;
                float4 *eo = reinterpret_cast<float4 *>(b.rec_eye + r);
                const float4 *es = reinterpret_cast<const float4 *>(b.rec_eye + ro);
                eo[j] = __ldcg(es + jo);
                eo[3 + j] = __ldcg(es + 3 + jo);
            }
        }
    }
}

void launch_setup(const BatchDev &b, const FrameTargets &fb, cudaStream_t s)
{
    if (b.n_triangles == 0) return;
    uint32_t chunks = (b.n_triangles + SETUP_THREADS - 1) / SETUP_THREADS;
    k_setup<<<chunks, SETUP_THREADS, 0, s>>>(b, fb);
    note_launch();
}

} // namespace mtgl_dev_impl
