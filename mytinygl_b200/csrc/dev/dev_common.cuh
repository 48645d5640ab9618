/*
 * dev_common.cuh -- device-side data layout and arithmetic helpers of the sm_100a back end.
 *
 * Parity discipline (hard part 1): this translation-unit set is compiled with
 * -fmad=false -prec-div=true -prec-sqrt=true -ftz=false, so every a*b+c below rounds twice like
 * the reference's strict build, divisions and square roots are IEEE, and denormals survive.
 * The helpers restate the reference's expressions operation for operation; each cites its source.
 */
#ifndef MTGL_DEV_COMMON_CUH
#define MTGL_DEV_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "mtgl_dev.h"

namespace mtgl_dev_impl {

/* ---- GL tokens the kernels switch on (Khronos values; GL_PHONG is the reference's gl.h:222) ---- */
enum : uint32_t {
    G_POINTS = 0, G_LINES = 1, G_LINE_LOOP = 2, G_LINE_STRIP = 3, G_TRIANGLES = 4, G_TRIANGLE_STRIP = 5,
    G_TRIANGLE_FAN = 6, G_QUADS = 7, G_QUAD_STRIP = 8, G_POLYGON = 9,
    G_NEVER = 0x0200, G_LESS = 0x0201, G_EQUAL = 0x0202, G_LEQUAL = 0x0203, G_GREATER = 0x0204,
    G_NOTEQUAL = 0x0205, G_GEQUAL = 0x0206, G_ALWAYS = 0x0207,
    G_ZERO = 0, G_ONE = 1, G_SRC_COLOR = 0x0300, G_ONE_MINUS_SRC_COLOR = 0x0301, G_SRC_ALPHA = 0x0302,
    G_ONE_MINUS_SRC_ALPHA = 0x0303, G_DST_ALPHA = 0x0304, G_ONE_MINUS_DST_ALPHA = 0x0305, G_DST_COLOR = 0x0306,
    G_ONE_MINUS_DST_COLOR = 0x0307, G_SRC_ALPHA_SATURATE = 0x0308,
    G_FRONT = 0x0404, G_BACK = 0x0405, G_FRONT_AND_BACK = 0x0408, G_CW = 0x0900, G_CCW = 0x0901,
    G_FASTEST = 0x1101,
    G_AMBIENT = 0x1200, G_DIFFUSE = 0x1201, G_SPECULAR = 0x1202, G_EMISSION = 0x1600, G_AMBIENT_AND_DIFFUSE = 0x1602,
    G_POINT = 0x1B00, G_LINE = 0x1B01, G_FILL = 0x1B02,
    G_FLAT = 0x1D00, G_SMOOTH = 0x1D01, G_PHONG = 0x1D02,
    G_KEEP = 0x1E00, G_REPLACE = 0x1E01, G_INCR = 0x1E02, G_DECR = 0x1E03, G_INVERT = 0x150A,
    G_INCR_WRAP = 0x8507, G_DECR_WRAP = 0x8508,
    G_EXP = 0x0800, G_EXP2 = 0x0801, G_LINEAR = 0x2601, G_NEAREST = 0x2600,
    G_NEAREST_MIPMAP_NEAREST = 0x2700, G_LINEAR_MIPMAP_NEAREST = 0x2701, G_NEAREST_MIPMAP_LINEAR = 0x2702,
    G_LINEAR_MIPMAP_LINEAR = 0x2703,
    G_MODULATE = 0x2100, G_DECAL = 0x2101, G_BLEND = 0x0BE2, G_ADD = 0x0104,
    G_REPEAT = 0x2901,
    G_UNSIGNED_BYTE = 0x1401, G_UNSIGNED_SHORT = 0x1403, G_UNSIGNED_INT = 0x1405,
    G_COLOR_BUFFER_BIT = 0x4000, G_DEPTH_BUFFER_BIT = 0x0100, G_STENCIL_BUFFER_BIT = 0x0400
};

/* ---- screen tiling ---- */
constexpr int TILE_W = 64;
constexpr int TILE_H = 64;
constexpr int TILE_LOG = 6;
constexpr int RASTER_THREADS = 256;           /* 8 warps; each takes one 16x16 pixel region of the tile at a time */
constexpr int REGION_W = 16;
constexpr int REGION_H = 16;
constexpr int REGIONS_X = TILE_W / REGION_W;
constexpr int NUM_REGIONS = (TILE_W / REGION_W) * (TILE_H / REGION_H);
constexpr int COLOR_PITCH = 72;               /* words per tile row in shared memory: an 8x4 pixel block */
constexpr int STENCIL_PITCH = 80;             /* (the fragment quantum) hits 32 distinct banks            */
constexpr int LIST_WINDOW = 2048;             /* triangle references sorted + staged per pass             */
constexpr int SETUP_THREADS = 256;            /* one chunk = 256 input triangles                          */
constexpr int GROUP_SHIFT = 10;               /* record id = group << 10 | index-in-group; a group is 32   */
                                              /* consecutive input triangles (one warp of k_setup, <= 21*32 records) */
constexpr int LARGE_TILES = 16;               /* records overlapping more tiles are binned cooperatively  */

/* ---- device views of objects ---- */
struct DevAttrib {
    const uint8_t *ptr;      /* element 0 (buffer base + offset), NULL when the buffer has no storage */
    uint64_t avail;          /* bytes available from ptr to the end of the buffer */
    uint32_t stride;
    uint16_t size, type;
    uint32_t enabled;
    uint32_t pad_;
};

/* Fast attribute path.  For a non-indexed array draw whose enabled arrays are 4-byte aligned floats and whose whole
 * element range lies inside the buffers (checked once, attrib_range_ok), element e of an attribute is base + e * stride:
 * no per-vertex bounds / type / alignment logic.  The values are the same loads fetch_attrib would do. */
struct FastDraw {
    const uint8_t *pos, *nrm, *tex;     /* element 0 of each array (NULL: array disabled) */
    uint32_t pos_stride, nrm_stride, tex_stride;
    uint32_t pos_size;
    int32_t first;
    uint32_t tri_begin, tri_end;        /* global indices of the draw's triangles in this pass */
    uint32_t tbase;
    uint32_t draw;                      /* index into BatchDev::draws */
    uint32_t valid;
    float cur_color[4], cur_normal[3], cur_texcoord[2];
    const mtgl_state *st;
};

struct DevDraw {
    uint32_t mode, count, raster_state, source;
    uint32_t first_staged, vertex_state;
    int32_t first;
    uint32_t index_type;
    const uint8_t *index_ptr;
    uint64_t index_avail;
    DevAttrib position, color, texcoord, normal;
    float cur_color[4];
    float cur_texcoord[2];
    float cur_normal[3];
    uint32_t vbase;          /* index of this draw's first post-transform vertex */
    uint32_t tbase;          /* index of this draw's first assembled primitive (triangle, line segment or point) */
    uint32_t ntris;          /* primitives of this draw in the pass */
    uint32_t fused;          /* independent triangles: no vertex-stage launch, k_setup shades the survivors' vertices */
    uint32_t shared_verts;   /* indexed draw whose buffer ELEMENTS 0 .. shared_verts-1 go through the vertex stage once (plus one slot of
                              * default attributes for out-of-range indices): k_setup looks a vertex up by its index.  0 = off */
    uint32_t pad3_;
    uint32_t state_max;      /* n_states - 1 of the batch: a staged vertex's state index is clamped to it (a stale index from a broken
                              * binding must not become an out-of-bounds read) */
    const float4 *bounds;    /* object-space boxes of the draw's 256-triangle chunks (k_cull.cu), or NULL */
    FastDraw fast;           /* the fast attribute path of this draw in this pass, decided on the host (dev_vertex.cuh: fast_draw_init) */
};

/* Raster-stage view of one mtgl_state: enums folded to small integers, texture resolved to
 * device pointers.  Built on the host once per state block of a batch. */
struct RasterCfg {
    uint32_t flags;                 /* RC_* */
    uint32_t depth_func, alpha_func, stencil_func;   /* 0..7 = GL_NEVER..GL_ALWAYS */
    uint32_t stencil_fail, stencil_zfail, stencil_zpass;   /* GL tokens */
    int32_t stencil_ref;
    uint32_t stencil_mask, stencil_writemask;
    uint32_t blend_src, blend_dst;  /* GL tokens */
    uint32_t color_mask;
    uint32_t tex_env_mode, fog_mode;
    uint32_t tex_min, tex_mag, tex_wrap_s, tex_wrap_t;
    int32_t tex_w, tex_h, tex_w1, tex_h1;
    const uint32_t *tex_l0, *tex_l1;
    const float4 *tex_f4;           /* the same texels as n / 255 floats, level 0 then level 1 (small textures only, else NULL): dev_fasttex.cuh */
    uint32_t fast_tex;              /* 1: the float4 copy exists and every REPEAT axis has a power-of-two size */
    uint32_t pad_;
    float alpha_ref;
    float fog_density, fog_start, fog_end;
    float fog_color[4];
    float tex_env_color[4];
    double depth_near, depth_far;
};

enum : uint32_t {
    RC_DEPTH_TEST = 1u << 0, RC_DEPTH_WRITE = 1u << 1, RC_STENCIL = 1u << 2, RC_BLEND = 1u << 3,
    RC_TEXTURED = 1u << 4, RC_ALPHA_TEST = 1u << 5, RC_FOG = 1u << 6, RC_FLAT = 1u << 7,
    RC_PHONG = 1u << 8, RC_LIGHTING = 1u << 9, RC_TWO_SIDE = 1u << 10, RC_PERSPECTIVE = 1u << 11,
    RC_DEPTH_RANGE_01 = 1u << 12,   /* depth range is exactly [0,1]: the double expression of raster.c:548 is the float one */
    RC_DEFER = 1u << 13,            /* no blending, no alpha test, full colour mask: the colour work can be deferred */
    RC_UNORDERED = 1u << 14         /* RC_DEFER + depth test and write, no stencil, the batch's common LESS/LEQUAL/GREATER/GEQUAL
                                     * function: the surviving fragment of a pixel does not depend on submission order (k_vis.cu) */
};

/* TriRecord.state_flags: state block index | bounded << 26 | unordered << 27 | record kind << 28 | deferrable << 30 | back-facing << 31 */
constexpr uint32_t STATE_INDEX_MASK = 0x03FFFFFFu;
constexpr uint32_t STATE_BOUNDED_BIT = 1u << 26;        /* filled triangle whose texture coordinates and 1/w are finite and bounded (attr_bounded) */
constexpr uint32_t STATE_UNORD_BIT = 1u << 27;
constexpr uint32_t STATE_KIND_SHIFT = 28;
constexpr uint32_t STATE_KIND_MASK = 3u << STATE_KIND_SHIFT;
constexpr uint32_t KIND_TRIANGLE = 0u, KIND_LINE = 1u, KIND_POINT = 2u;
constexpr uint32_t STATE_DEFER_BIT = 1u << 30;
constexpr uint32_t STATE_BACK_BIT = 1u << 31;

/* |u|, |v| <= 2^20 and 2^-40 <= 1/w <= 2^40 at all three vertices: no NaN or infinity can then arise in the interpolated
 * texture coordinates, which is what the staged / float4 texture path (dev_fasttex.cuh) relies on */
__device__ __forceinline__ bool attr_bounded(float u0, float v0, float w0, float u1, float v1, float w1, float u2, float v2, float w2)
{
    const float lim = 1048576.0f, wlo = 9.094947e-13f, whi = 1.0995116e12f;
    return fabsf(u0) <= lim && fabsf(v0) <= lim && fabsf(u1) <= lim && fabsf(v1) <= lim && fabsf(u2) <= lim && fabsf(v2) <= lim &&
           w0 >= wlo && w0 <= whi && w1 >= wlo && w1 <= whi && w2 >= wlo && w2 <= whi;
}

/* what a record contributes to the flags of every tile it is binned into */
__device__ __forceinline__ uint32_t tile_flag_bits(uint32_t state_flags)
{
    return ((state_flags & STATE_DEFER_BIT) ? 0u : 1u) | ((state_flags & STATE_UNORD_BIT) ? 0u : 2u) |
           ((state_flags & STATE_KIND_MASK) ? 4u : 0u);
}

/* One set-up primitive: 10 x 16 B.  For a sub-triangle the fields mean what their names say.  A LINE record
 * (draw_line_full, raster.c:107-241) stores its end points in (x0,y0)-(x1,y1), the line width in x2, NDC z in z0/z1,
 * end colours in c0/c1, texture coordinates in (u0,v0)/(u1,v1) and eye z in ez0/ez1.  A POINT record (flush_points
 * 1020-1164, draw_point_at_screen 777-844) stores the top-left corner of its square in (x0,y0), the size in x1, its
 * final depth in z0 and its final colour (after texture, alpha test and fog, all per vertex) in c0.
 * One set-up sub-triangle: 10 x 16 B.  Row 2 (clamped bounding box, state, ordered id) is all the
 * binner and the tile kernel's list builder read. */
struct __align__(16) TriRecord {
    int32_t x0, y0, x1, y1;                 /* row 0: integer-snapped vertices (raster.c:59-63) */
    int32_t x2, y2; float area, inv_area;   /* row 1: signed doubled area (raster.c:483) and its reciprocal */
    uint32_t bbox_min, bbox_max;            /* row 2: inclusive bbox after viewport/scissor/framebuffer/band clamps, x | y << 16 */
    uint32_t state_flags, id;               /*        state index | deferrable << 30 | back-facing << 31 ; submission-ordered id */
    float z0, z1, z2, lod;                  /* row 3 */
    float w0, w1, w2, ez0;                  /* row 4: 1/w per vertex */
    float c0[4], c1[4], c2[4];              /* rows 5-7 */
    float u0, v0, u1, v1;                   /* row 8 */
    float u2, v2, ez1, ez2;                 /* row 9 */
};
static_assert(sizeof(TriRecord) == 160, "TriRecord layout");

/* eye-space position + normal per vertex, only materialised for GL_PHONG / two-sided lighting */
struct __align__(16) TriEye {
    float ep0[4], ep1[4], ep2[4];
    float en0[4], en1[4], en2[4];
};

struct DevCounters {
    unsigned int records;        /* set-up sub-triangles written */
    unsigned int large_count;    /* records handed to the cooperative binner */
    unsigned int tile_refs;      /* total (record, tile) references = list length */
    unsigned int triangles_in;
    unsigned int overflow;
    unsigned int culled_chunks;  /* 256-triangle chunks dropped by the culling pass (k_cull.cu) */
    unsigned int scan_ticket;    /* CTAs of k_bin_scan that have finished counting: the last one scans */
    unsigned int pad_[1];
};

struct Color4 { float r, g, b, a; };

/* ---------------------------------------------------------------- scalar helpers */

/* x86 cvttss2si semantics: the reference's (int32_t) casts of out-of-range / NaN floats yield INT_MIN */
__device__ __forceinline__ int32_t f2i_x86(float f)
{
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return (int32_t)0x80000000;
    return __float2int_rz(f);
}

__device__ __forceinline__ float sat01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }   /* keeps NaN like the C ternaries */

__device__ __forceinline__ Color4 color_clamp(Color4 c) { return { sat01(c.r), sat01(c.g), sat01(c.b), sat01(c.a) }; }

/* color_to_rgba32 (graphics.h:337-348): clamp, truncate, pack */
__device__ __forceinline__ uint32_t color_pack(Color4 c)
{
    /* __saturatef maps NaN to 0 where the reference's ternaries keep it, but NaN * 255 packs to 0 too (cvttss2si: low
     * byte of 0x80000000), so the bytes are identical for every input */
    uint32_t r = __float2uint_rz(__saturatef(c.r) * 255.0f);
    uint32_t g = __float2uint_rz(__saturatef(c.g) * 255.0f);
    uint32_t b = __float2uint_rz(__saturatef(c.b) * 255.0f);
    uint32_t a = __float2uint_rz(__saturatef(c.a) * 255.0f);
    return (a << 24) | (b << 16) | (g << 8) | r;
}

/* color_from_rgba32 (graphics.h:350-357).  n / 255.0f for n in 0..255 through a 256-entry table of
 * the correctly rounded quotients (filled with the IEEE division itself, so bit-identical). */
__device__ __forceinline__ Color4 color_unpack(uint32_t p, const float *unorm8)
{
    return { unorm8[p & 0xFFu], unorm8[(p >> 8) & 0xFFu], unorm8[(p >> 16) & 0xFFu], unorm8[p >> 24] };
}

__device__ __forceinline__ Color4 color_lerp(Color4 a, Color4 b, float t)   /* graphics.h:293-295: a*(1-t) + b*t */
{
    float s = 1.0f - t;
    return { a.r * s + b.r * t, a.g * s + b.g * t, a.b * s + b.b * t, a.a * s + b.a * t };
}

__device__ __forceinline__ Color4 color_lerp_rgb(Color4 a, Color4 b, float t)   /* graphics.h:298-305 */
{
    return { a.r + (b.r - a.r) * t, a.g + (b.g - a.g) * t, a.b + (b.b - a.b) * t, a.a };
}

__device__ __forceinline__ bool compare_f(uint32_t func, float a, float b)   /* raster.c:344-357, 391-404; func = token - GL_NEVER */
{
    switch (func) {
    case 0: return false;
    case 1: return a < b;
    case 2: return a == b;
    case 3: return a <= b;
    case 4: return a > b;
    case 5: return a != b;
    case 6: return a >= b;
    default: return true;
    }
}

__device__ __forceinline__ Color4 blend_factor(uint32_t f, Color4 s, Color4 d)   /* raster.c:360-379 == drawpixels_blend_factor, gl_api.c:1266-1284 */
{
    switch (f) {
    case G_ZERO: return { 0.0f, 0.0f, 0.0f, 0.0f };
    case G_SRC_COLOR: return s;
    case G_ONE_MINUS_SRC_COLOR: return { 1 - s.r, 1 - s.g, 1 - s.b, 1 - s.a };
    case G_DST_COLOR: return d;
    case G_ONE_MINUS_DST_COLOR: return { 1 - d.r, 1 - d.g, 1 - d.b, 1 - d.a };
    case G_SRC_ALPHA: return { s.a, s.a, s.a, s.a };
    case G_ONE_MINUS_SRC_ALPHA: return { 1 - s.a, 1 - s.a, 1 - s.a, 1 - s.a };
    case G_DST_ALPHA: return { d.a, d.a, d.a, d.a };
    case G_ONE_MINUS_DST_ALPHA: return { 1 - d.a, 1 - d.a, 1 - d.a, 1 - d.a };
    case G_SRC_ALPHA_SATURATE: { float k = (s.a < (1 - d.a)) ? s.a : (1 - d.a); return { k, k, k, 1.0f }; }
    default: return { 1.0f, 1.0f, 1.0f, 1.0f };     /* GL_ONE and the accepted-but-unimplemented GL_CONSTANT_* */
    }
}

__device__ __forceinline__ bool compare_i(uint32_t func, int32_t a, int32_t b)   /* raster.c:407-422 */
{
    switch (func) {
    case 0: return false;
    case 1: return a < b;
    case 2: return a == b;
    case 3: return a <= b;
    case 4: return a > b;
    case 5: return a != b;
    case 6: return a >= b;
    default: return true;
    }
}

__device__ __forceinline__ void normalize3(float &x, float &y, float &z)   /* vec3_normalize, graphics.h:54-60 */
{
    float len = sqrtf(x * x + y * y + z * z);
    if (len > 0.0f) {
        float s = 1.0f / len;
        x *= s; y *= s; z *= s;
    }
}

/* compute_lighting (lighting.h:53-142).  'mat' may carry COLOR_MATERIAL overrides. */
struct MaterialRegs {
    float ambient[4], diffuse[4], specular[4], emission[4];
    float shininess;
};

__device__ __forceinline__ void load_material(MaterialRegs &m, const mtgl_material *s)
{
#pragma unroll
    for (int k = 0; k < 4; k++) {
        m.ambient[k] = s->ambient[k]; m.diffuse[k] = s->diffuse[k];
        m.specular[k] = s->specular[k]; m.emission[k] = s->emission[k];
    }
    m.shininess = s->shininess;
}

/* powf(x, y) for x > 0 as the lighting equations use it (lighting.h:103, 132).  glibc's powf is correctly rounded in
 * all but a vanishing fraction of cases; for the integral exponents that shininess and spot exponents usually are,
 * binary powering in double (<= 13 roundings of 2^-53, then one rounding to float) has the same property and costs a
 * dozen FP64 multiplies instead of CUDA's ~100-instruction powf, which is kept for every other exponent. */
__device__ __forceinline__ int integral_exponent(float y)
{
    return (y >= 1.0f && y <= 128.0f && y == truncf(y)) ? (int)y : 0;
}
__device__ __forceinline__ float pow_pos(float x, float y, int iy)
{
    if (iy == 0) return powf(x, y);
    double b = (double)x;
    if ((iy & (iy - 1)) == 0) {         /* 1, 2, 4 ... 128: squarings only (the exponent is uniform across the warp) */
        for (int n = iy; n > 1; n >>= 1) b *= b;
        return (float)b;
    }
    double r = 1.0;
    for (int n = iy; n; n >>= 1) {
        if (n & 1) r *= b;
        b *= b;
    }
    return (float)r;
}

__device__ __forceinline__ Color4 lighting_body(const mtgl_state *st, float px, float py, float pz,
                                                float nx, float ny, float nz, const MaterialRegs &mat)
{
    Color4 res;
    res.r = mat.emission[0] + mat.ambient[0] * st->light_model_ambient[0];
    res.g = mat.emission[1] + mat.ambient[1] * st->light_model_ambient[1];
    res.b = mat.emission[2] + mat.ambient[2] * st->light_model_ambient[2];
    res.a = mat.diffuse[3];
    if (st->caps & MTGL_CAP_NORMALIZE) normalize3(nx, ny, nz);
    const uint32_t local_viewer = st->light_model_local_viewer;
    const int ishine = integral_exponent(mat.shininess);

    for (int i = 0; i < MTGL_MAX_LIGHTS; i++) {
        const mtgl_light *l = &st->lights[i];
        if (!l->enabled) continue;
        float Lx, Ly, Lz, att = 1.0f;
        if (l->position[3] == 0.0f) {
            Lx = l->dir_unit[0]; Ly = l->dir_unit[1]; Lz = l->dir_unit[2];
        } else {
            float tx = l->position[0] - px, ty = l->position[1] - py, tz = l->position[2] - pz;
            float dist = sqrtf(tx * tx + ty * ty + tz * tz);
            if (dist < 1e-6f) dist = 1e-6f;
            float inv = 1.0f / dist;
            Lx = tx * inv; Ly = ty * inv; Lz = tz * inv;
            float den = l->att_constant + l->att_linear * dist + l->att_quadratic * dist * dist;
            if (den < 1e-6f) den = 1e-6f;
            att = 1.0f / den;
            if (l->spot_cutoff < 180.0f) {
                float cos_angle = -(Lx * l->spot_dir_unit[0] + Ly * l->spot_dir_unit[1] + Lz * l->spot_dir_unit[2]);
                if (cos_angle < l->cos_cutoff) att = 0.0f;
                else att *= (cos_angle > 0.0f) ? pow_pos(cos_angle, l->spot_exponent, integral_exponent(l->spot_exponent)) : powf(cos_angle, l->spot_exponent);
            }
        }
        if (att <= 0.0f) continue;

        res.r = res.r + (mat.ambient[0] * l->ambient[0]) * att;
        res.g = res.g + (mat.ambient[1] * l->ambient[1]) * att;
        res.b = res.b + (mat.ambient[2] * l->ambient[2]) * att;
        res.a = res.a + (mat.ambient[3] * l->ambient[3]) * att;

        float NdotL = nx * Lx + ny * Ly + nz * Lz;
        if (NdotL > 0.0f) {
            float k = att * NdotL;
            res.r = res.r + (mat.diffuse[0] * l->diffuse[0]) * k;
            res.g = res.g + (mat.diffuse[1] * l->diffuse[1]) * k;
            res.b = res.b + (mat.diffuse[2] * l->diffuse[2]) * k;
            res.a = res.a + (mat.diffuse[3] * l->diffuse[3]) * k;
            if (mat.shininess > 0.0f) {
                float Vx = 0.0f, Vy = 0.0f, Vz = 1.0f;
                if (local_viewer) {
                    Vx = px * -1.0f; Vy = py * -1.0f; Vz = pz * -1.0f;
                    normalize3(Vx, Vy, Vz);
                }
                float Hx = Lx + Vx, Hy = Ly + Vy, Hz = Lz + Vz;
                normalize3(Hx, Hy, Hz);
                float NdotH = nx * Hx + ny * Hy + nz * Hz;
                if (NdotH > 0.0f) {
                    float spec = pow_pos(NdotH, mat.shininess, ishine) * att;
                    res.r = res.r + (mat.specular[0] * l->specular[0]) * spec;
                    res.g = res.g + (mat.specular[1] * l->specular[1]) * spec;
                    res.b = res.b + (mat.specular[2] * l->specular[2]) * spec;
                    res.a = res.a + (mat.specular[3] * l->specular[3]) * spec;
                }
            }
        }
    }
    return color_clamp(res);
}

/* out-of-line copy for the register-bound raster / set-up kernels */
static __device__ __noinline__ Color4 compute_lighting(const mtgl_state *st, float px, float py, float pz,
                                                float nx, float ny, float nz, const MaterialRegs &mat)
{
    return lighting_body(st, px, py, pz, nx, ny, nz, mat);
}

/* ---------------------------------------------------------------- launch wrappers (defined in the .cu files) */
struct FrameTargets {
    uint32_t *color; float *depth; uint8_t *stencil;
    uint32_t *present;              /* peer GPU's colour plane (NVLink-mapped) that colour stores are mirrored into, or NULL */
    int32_t width, height;
    int32_t band_y0, band_y1;       /* rows owned by this device */
    int32_t tiles_x;                /* tiles per row */
    int32_t tile_y0, tile_rows;     /* first tile row and number of tile rows covering the band */
};

struct ClearOp {
    uint32_t mask;
    int32_t x0, y0, x1, y1;
    uint32_t color; float depth; uint32_t stencil;
};

struct BatchDev {
    const mtgl_state *states;
    const RasterCfg *cfgs;
    const mtgl_in_vertex *staged;
    const DevDraw *draws;
    const uint32_t *draw_vbase;     /* n_draws + 1 */
    const uint32_t *draw_tbase;     /* n_draws + 1 */
    uint32_t n_draws, n_vertices, n_triangles;
    uint32_t n_states;              /* entries of states[] / cfgs[] */
    uint32_t need_eye;
    uint32_t n_unfused_draws;       /* draws whose vertices go through k_vertex */
    /* post-transform vertices */
    float4 *v_clip, *v_color, *v_tex, *v_epos, *v_enrm;
    /* set-up output */
    TriRecord *records; TriEye *rec_eye; uint32_t record_capacity;
    uint32_t *group_base;           /* first record slot of every group of 32 input triangles (one warp of k_setup) */
    uint8_t *chunk_cull;            /* 1 = the chunk cannot produce a record on this device (k_cull.cu); NULL = no culling pass */
    uint32_t *large_list;
    uint4 *bin_rows;                /* copy of every record's row 2 (bbox_min, bbox_max, state_flags, id): all the binner reads */
    DevCounters *counters;
    DevCounters *host_counters;     /* pinned host copy, device-mapped (unified addressing): k_bin_scan publishes the counters there */
    /* binning */
    uint32_t *tile_count, *tile_offset, *tile_cursor;
    uint32_t *tile_flags;           /* bit 0: the tile references a record whose colour work cannot be deferred (general kernel);
                                     * bit 1: it references a record outside the unordered class (sorted visibility kernel);
                                     * bit 2: it references a line or a point (tile_flag_bits) */
    uint32_t *tile_order;           /* launch order of the tile kernels: blockIdx -> tile, heaviest lists first (k_bin_scan); NULL = identity */
    uint32_t *tile_list; uint32_t list_capacity;
    uint32_t guard;                 /* 1: the list buffer was sized by guess -- fill and raster kernels must check lists_fit() */
    uint32_t *vis_plane;            /* visibility buffer in HBM (record index per pixel) between K4a and K4b */
    const float *unorm8;
};

/* optimistic batches (mtgl_dev.cu): when the scanned reference count does not fit the list buffer, or set-up ran out of
 * record slots, the fill and raster kernels must not touch anything; the host re-queues them with a larger buffer */
__device__ __forceinline__ bool lists_fit(const BatchDev &b)
{
    return !b.guard || (b.counters->tile_refs <= b.list_capacity && b.counters->overflow == 0u);
}

void launch_vertex_stage(const BatchDev &b, cudaStream_t s);
void launch_setup(const BatchDev &b, const FrameTargets &fb, cudaStream_t s);
void launch_chunk_bounds(const uint8_t *pos, uint32_t stride, uint32_t size, int32_t first, uint32_t nverts, float4 *out, cudaStream_t s);
void launch_chunk_cull(const BatchDev &b, const FrameTargets &fb, cudaStream_t s);
void launch_bin_scan(const BatchDev &b, const FrameTargets &fb, cudaStream_t s);
void launch_bin_fill(const BatchDev &b, const FrameTargets &fb, cudaStream_t s);
/* which raster kernels a pass needs, decided on the host from the raster states its draws use */
struct RasterPlan {
    bool any_deferrable;        /* some draw has a deferrable state (K4a + K4b) */
    bool any_ordered_vis;       /* ... that is not in the unordered class (sorted K4a) */
    bool any_in_order;          /* some draw needs in-order shading (general kernel) */
    bool plain_in_order;        /* ... and the pass holds nothing but in-order filled triangles */
    uint32_t unordered_func;    /* depth function (0..7) of the unordered class, 0 when the class is empty */
    bool unordered_range01;     /* every unordered state has depth range [0,1] */
    uint32_t fill_mode;         /* FILL_* (dev_fill.cuh): may the pixel-owner kernel (k_fill.cu) take in-order tiles of large triangles */
    uint32_t in_order_all, in_order_any;    /* AND / OR of the RasterCfg flags of the pass's in-order states */
    cudaEvent_t color_gate;     /* the kernels that write the colour plane wait for this event first (MTGL_PRESENT_COPY: the previous
                                 * frame's band is still being copied out of the plane), or NULL */
};
/* ev_vis / ev_shade are recorded after the visibility kernels and after the shade kernel (stage timing) */
void launch_raster(const BatchDev &b, const FrameTargets &fb, const ClearOp &clear, uint32_t plane_rw_mask,
                   const RasterPlan &plan, cudaStream_t s, cudaEvent_t ev_vis, cudaEvent_t ev_shade);
void launch_fill(const BatchDev &b, const FrameTargets &fb, const ClearOp &clear, uint32_t planes, uint32_t fill_mode, uint32_t all_on, uint32_t any_on,
                 cudaStream_t s);
void launch_shade(const BatchDev &b, const FrameTargets &fb, const ClearOp &clear, uint32_t all_on, uint32_t any_on, cudaStream_t s);
void launch_vis_unordered(const BatchDev &b, const FrameTargets &fb, const ClearOp &clear, uint32_t planes, uint32_t depth_func,
                          bool all_range01, cudaStream_t s);
void launch_draw_pixels(const ::mtgl_pixel_rect &rect, const uint8_t *src, const FrameTargets &fb, const float *unorm8, cudaStream_t s);
void launch_read_pixels(const FrameTargets &fb, int32_t x, int32_t y, int32_t w, int32_t h, uint32_t bpp, uint8_t *dst, cudaStream_t s);
void launch_frame_barrier(unsigned long long *counter, unsigned long long target, unsigned long long timeout_ns, uint32_t *timed_out, cudaStream_t s);
void launch_upload(const void *host_mapped, void *dst, size_t bytes, void *zero, size_t zero_bytes, cudaStream_t s);
void launch_mip1(const uint32_t *l0, int w, int h, uint32_t *l1, const float *unorm8, cudaStream_t s);
void launch_tex_f4(const uint32_t *l0, int n0, const uint32_t *l1, int n1, float4 *out, const float *unorm8, cudaStream_t s);
void launch_fill_unorm8(float *table, cudaStream_t s);
uint64_t kernel_launch_count();
/* A tile grid well below one wave of 8-warp CTAs (148 SMs x 4 = 592): the band of a multi-GPU frame (measured: the
 * latency shapes win at 240-300 tiles, lose at 540).  The tile kernels
 * then run in their latency shapes -- more warps per tile (k_vis<512>), sub-tile CTAs (k_shade<4>). */
bool small_grid(uint32_t tiles, uint32_t limit = 400u);      /* tiles <= limit, or what MTGL_GRID_SHAPE=small|large forces (tests cover both shapes at any size) */

} // namespace mtgl_dev_impl

#endif
