/*
 * mtgl_oracle.c -- TEST INFRASTRUCTURE ONLY.  Scalar CPU restatement of the reference renderer's
 * hot path (zbufferoverflow/MyTinyGL, src/gl_api.c emit_vertex, src/lighting.h, src/clipping.h,
 * src/raster.c, src/textures.c sampling half) behind the same C ABI as the CUDA back end
 * (include/mtgl_dev.h).  It exists so that tests can check the device path on arbitrary state
 * blocks; it is never linked into the product library and the product never calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py renders every scene of
 * tests/scenes/scenes.c through (front end + this file) and through the unmodified reference
 * compiled from /root/reference (oracle/_ref/libref_strict.so) and requires bit-identical colour,
 * depth and stencil planes; the resulting plane hashes are committed under tests/golden/.
 *
 * Must be compiled with IEEE semantics:  -O2 -fno-fast-math -ffp-contract=off  (the canonical
 * "strict" build of the reference).  Every function cites the reference
 * lines it restates.  libm calls (powf, expf, log2f) are glibc's, like the reference's.
 */
#include "mtgl_dev.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* GL tokens used by the hot path (values: Khronos registry; GL_PHONG is the reference's, gl.h:222) */
enum {
    T_POINTS = 0, T_LINES = 1, T_LINE_LOOP = 2, T_LINE_STRIP = 3, T_TRIANGLES = 4, T_TRIANGLE_STRIP = 5,
    T_TRIANGLE_FAN = 6, T_QUADS = 7, T_QUAD_STRIP = 8, T_POLYGON = 9,
    T_NEVER = 0x0200, T_LESS, T_EQUAL, T_LEQUAL, T_GREATER, T_NOTEQUAL, T_GEQUAL, T_ALWAYS,
    T_ZERO = 0, T_ONE = 1, T_SRC_COLOR = 0x0300, T_ONE_MINUS_SRC_COLOR, T_SRC_ALPHA, T_ONE_MINUS_SRC_ALPHA,
    T_DST_ALPHA, T_ONE_MINUS_DST_ALPHA, T_DST_COLOR, T_ONE_MINUS_DST_COLOR, T_SRC_ALPHA_SATURATE,
    T_FRONT = 0x0404, T_BACK = 0x0405, T_FRONT_AND_BACK = 0x0408,
    T_CW = 0x0900, T_CCW = 0x0901,
    T_FASTEST = 0x1101,
    T_AMBIENT = 0x1200, T_DIFFUSE = 0x1201, T_SPECULAR = 0x1202, T_EMISSION = 0x1600, T_AMBIENT_AND_DIFFUSE = 0x1602,
    T_POINT = 0x1B00, T_LINE = 0x1B01, T_FILL = 0x1B02,
    T_FLAT = 0x1D00, T_SMOOTH = 0x1D01, T_PHONG = 0x1D02,
    T_KEEP = 0x1E00, T_REPLACE = 0x1E01, T_INCR = 0x1E02, T_DECR = 0x1E03, T_INVERT = 0x150A,
    T_INCR_WRAP = 0x8507, T_DECR_WRAP = 0x8508,
    T_EXP = 0x0800, T_EXP2 = 0x0801, T_LINEAR_FOG = 0x2601,
    T_MODULATE = 0x2100, T_DECAL = 0x2101, T_BLEND_ENV = 0x0BE2, T_ADD = 0x0104,
    T_NEAREST = 0x2600, T_LINEAR = 0x2601, T_NEAREST_MIPMAP_NEAREST = 0x2700, T_LINEAR_MIPMAP_NEAREST = 0x2701,
    T_NEAREST_MIPMAP_LINEAR = 0x2702, T_LINEAR_MIPMAP_LINEAR = 0x2703,
    T_REPEAT = 0x2901,
    T_UNSIGNED_BYTE = 0x1401, T_UNSIGNED_SHORT = 0x1403, T_UNSIGNED_INT = 0x1405,
    T_COLOR_BUFFER_BIT = 0x4000, T_DEPTH_BUFFER_BIT = 0x0100, T_STENCIL_BUFFER_BIT = 0x0400
};

typedef struct { float r, g, b, a; } col4;

/* post-vertex-stage record: vertex_t (graphics.h:438-446) minus the object-space normal,
 * which nothing downstream reads */
typedef struct {
    float pos[4];
    col4 color;
    float uv[2];
    float eye_z;
    float eye_pos[3];
    float eye_nrm[3];
} overt;

typedef struct {
    int32_t w, h;
    uint32_t *px;
    int32_t w1, h1;
    uint32_t *px1;   /* NULL when the level cannot exist (w < 2 or h < 2, textures.c:317) */
} otex;

typedef struct {
    uint8_t *data;
    uint64_t size;
} obuf;

#define O_MAX_OBJECTS 257
#define O_MAX_BUFFERS MTGL_MAX_BUFFER_IDS

struct mtgl_dev {
    int32_t width, height;
    int32_t band_y0, band_y1;
    uint32_t *color;
    float *depth;
    uint8_t *stencil;
    otex tex[O_MAX_OBJECTS];
    obuf buf[O_MAX_BUFFERS];
    mtgl_dev_stats stats;
    uint64_t frag_covered, frag_tested, frag_shaded;   /* cumulative: inside test / past stencil+depth / colour write */
    double t_mark[2];
    char err[128];
};

/* ------------------------------------------------------------------------------------------ */
/* colour helpers (graphics.h:293-366)                                                        */
/* ------------------------------------------------------------------------------------------ */
static float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

static col4 col_clamp(col4 c) /* graphics.h:359-366 */
{
    col4 o = { c.r < 0 ? 0 : (c.r > 1 ? 1 : c.r), c.g < 0 ? 0 : (c.g > 1 ? 1 : c.g),
               c.b < 0 ? 0 : (c.b > 1 ? 1 : c.b), c.a < 0 ? 0 : (c.a > 1 ? 1 : c.a) };
    return o;
}

static uint32_t col_pack(col4 c) /* graphics.h:337-348: clamp, truncate, pack a<<24|b<<16|g<<8|r */
{
    uint8_t r = (uint8_t)(clamp01(c.r) * 255.0f);
    uint8_t g = (uint8_t)(clamp01(c.g) * 255.0f);
    uint8_t b = (uint8_t)(clamp01(c.b) * 255.0f);
    uint8_t a = (uint8_t)(clamp01(c.a) * 255.0f);
    return ((uint32_t)a << 24) | ((uint32_t)b << 16) | ((uint32_t)g << 8) | r;
}

static col4 col_unpack(uint32_t p) /* graphics.h:350-357: true division by 255 */
{
    col4 c = { (p & 0xFF) / 255.0f, ((p >> 8) & 0xFF) / 255.0f, ((p >> 16) & 0xFF) / 255.0f,
               ((p >> 24) & 0xFF) / 255.0f };
    return c;
}

static col4 col_lerp(col4 a, col4 b, float t) /* graphics.h:293-295: a*(1-t) + b*t */
{
    float s = 1.0f - t;
    col4 o = { a.r * s + b.r * t, a.g * s + b.g * t, a.b * s + b.b * t, a.a * s + b.a * t };
    return o;
}

static col4 col_lerp_rgb(col4 a, col4 b, float t) /* graphics.h:298-305: keeps a's alpha */
{
    col4 o = { a.r + (b.r - a.r) * t, a.g + (b.g - a.g) * t, a.b + (b.b - a.b) * t, a.a };
    return o;
}

static col4 col_from(const float *p) { col4 c = { p[0], p[1], p[2], p[3] }; return c; }

/* ------------------------------------------------------------------------------------------ */
/* vertex stage                                                                               */
/* ------------------------------------------------------------------------------------------ */
static void mat_vec4(const float *m, float x, float y, float z, float w, float *o) /* graphics.h:131-138 */
{
    o[0] = m[0] * x + m[4] * y + m[8] * z + m[12] * w;
    o[1] = m[1] * x + m[5] * y + m[9] * z + m[13] * w;
    o[2] = m[2] * x + m[6] * y + m[10] * z + m[14] * w;
    o[3] = m[3] * x + m[7] * y + m[11] * z + m[15] * w;
}

static void v3_normalize(float *v) /* graphics.h:54-60 */
{
    float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (len > 0.0f) {
        float s = 1.0f / len;
        v[0] *= s; v[1] *= s; v[2] *= s;
    }
}

/* compute_lighting (lighting.h:53-142).  dir_unit / spot_dir_unit / cos_cutoff are the
 * state-only sub-expressions of lines 77, 101 and 103, precomputed by the front end. */
static col4 light_vertex(const mtgl_state *st, const float *eye_pos, const float *eye_nrm, const mtgl_material *mat)
{
    col4 res;
    res.r = mat->emission[0] + mat->ambient[0] * st->light_model_ambient[0];
    res.g = mat->emission[1] + mat->ambient[1] * st->light_model_ambient[1];
    res.b = mat->emission[2] + mat->ambient[2] * st->light_model_ambient[2];
    res.a = mat->diffuse[3];

    float N[3] = { eye_nrm[0], eye_nrm[1], eye_nrm[2] };
    if (st->caps & MTGL_CAP_NORMALIZE) v3_normalize(N);

    for (int i = 0; i < MTGL_MAX_LIGHTS; i++) {
        const mtgl_light *l = &st->lights[i];
        if (!l->enabled) continue;
        float L[3];
        float att = 1.0f;
        if (l->position[3] == 0.0f) {
            L[0] = l->dir_unit[0]; L[1] = l->dir_unit[1]; L[2] = l->dir_unit[2];
        } else {
            float tl[3] = { l->position[0] - eye_pos[0], l->position[1] - eye_pos[1], l->position[2] - eye_pos[2] };
            float dist = sqrtf(tl[0] * tl[0] + tl[1] * tl[1] + tl[2] * tl[2]);
            if (dist < 1e-6f) dist = 1e-6f;
            float inv = 1.0f / dist;
            L[0] = tl[0] * inv; L[1] = tl[1] * inv; L[2] = tl[2] * inv;
            float den = l->att_constant + l->att_linear * dist + l->att_quadratic * dist * dist;
            if (den < 1e-6f) den = 1e-6f;
            att = 1.0f / den;
            if (l->spot_cutoff < 180.0f) {
                float cos_angle = -(L[0] * l->spot_dir_unit[0] + L[1] * l->spot_dir_unit[1] + L[2] * l->spot_dir_unit[2]);
                if (cos_angle < l->cos_cutoff) att = 0.0f;
                else att *= powf(cos_angle, l->spot_exponent);
            }
        }
        if (att <= 0.0f) continue;

        res.r = res.r + (mat->ambient[0] * l->ambient[0]) * att;
        res.g = res.g + (mat->ambient[1] * l->ambient[1]) * att;
        res.b = res.b + (mat->ambient[2] * l->ambient[2]) * att;
        res.a = res.a + (mat->ambient[3] * l->ambient[3]) * att;

        float NdotL = N[0] * L[0] + N[1] * L[1] + N[2] * L[2];
        if (NdotL > 0.0f) {
            float k = att * NdotL;
            res.r = res.r + (mat->diffuse[0] * l->diffuse[0]) * k;
            res.g = res.g + (mat->diffuse[1] * l->diffuse[1]) * k;
            res.b = res.b + (mat->diffuse[2] * l->diffuse[2]) * k;
            res.a = res.a + (mat->diffuse[3] * l->diffuse[3]) * k;
            if (mat->shininess > 0.0f) {
                float V[3] = { 0.0f, 0.0f, 1.0f };
                if (st->light_model_local_viewer) {
                    V[0] = eye_pos[0] * -1.0f; V[1] = eye_pos[1] * -1.0f; V[2] = eye_pos[2] * -1.0f;
                    v3_normalize(V);
                }
                float H[3] = { L[0] + V[0], L[1] + V[1], L[2] + V[2] };
                v3_normalize(H);
                float NdotH = N[0] * H[0] + N[1] * H[1] + N[2] * H[2];
                if (NdotH > 0.0f) {
                    float spec = powf(NdotH, mat->shininess) * att;
                    res.r = res.r + (mat->specular[0] * l->specular[0]) * spec;
                    res.g = res.g + (mat->specular[1] * l->specular[1]) * spec;
                    res.b = res.b + (mat->specular[2] * l->specular[2]) * spec;
                    res.a = res.a + (mat->specular[3] * l->specular[3]) * spec;
                }
            }
        }
    }
    return col_clamp(res);
}

/* COLOR_MATERIAL override of gl_api.c:285-312 applied to a private copy of the materials */
static void apply_color_material(const mtgl_state *st, col4 cur, mtgl_material *front, mtgl_material *back)
{
    col4 c = col_clamp(cur);
    float v[4] = { c.r, c.g, c.b, c.a };
    uint32_t mode = st->color_material_mode, face = st->color_material_face;
    mtgl_material *m[2] = { (face == T_FRONT || face == T_FRONT_AND_BACK) ? front : NULL,
                            (face == T_BACK || face == T_FRONT_AND_BACK) ? back : NULL };
    for (int k = 0; k < 2; k++) {
        if (!m[k]) continue;
        if (mode == T_AMBIENT || mode == T_AMBIENT_AND_DIFFUSE) memcpy(m[k]->ambient, v, sizeof v);
        if (mode == T_DIFFUSE || mode == T_AMBIENT_AND_DIFFUSE) memcpy(m[k]->diffuse, v, sizeof v);
        if (mode == T_SPECULAR) memcpy(m[k]->specular, v, sizeof v);
        if (mode == T_EMISSION) memcpy(m[k]->emission, v, sizeof v);
    }
}

/* emit_vertex (gl_api.c:263-348) */
static void vertex_stage(const mtgl_state *st, const float *p3, col4 cur_color, const float *uv, const float *nrm,
                         overt *out)
{
    float eye[4];
    mat_vec4(st->modelview, p3[0], p3[1], p3[2], 1.0f, eye);
    out->eye_z = -eye[2];
    out->eye_pos[0] = eye[0]; out->eye_pos[1] = eye[1]; out->eye_pos[2] = eye[2];

    const float *nm = st->normal; /* 3 columns, stride 4; fourth matrix column is zero (graphics.h:247-257) */
    float en[3];
    en[0] = nm[0] * nrm[0] + nm[4] * nrm[1] + nm[8] * nrm[2] + 0.0f * 0.0f;
    en[1] = nm[1] * nrm[0] + nm[5] * nrm[1] + nm[9] * nrm[2] + 0.0f * 0.0f;
    en[2] = nm[2] * nrm[0] + nm[6] * nrm[1] + nm[10] * nrm[2] + 0.0f * 0.0f;
    v3_normalize(en);
    out->eye_nrm[0] = en[0]; out->eye_nrm[1] = en[1]; out->eye_nrm[2] = en[2];

    col4 vc = cur_color;
    if (st->caps & MTGL_CAP_LIGHTING) {
        mtgl_material front = st->material_front, back = st->material_back;
        if (st->caps & MTGL_CAP_COLOR_MATERIAL) apply_color_material(st, cur_color, &front, &back);
        if (st->shade_model != T_PHONG) vc = light_vertex(st, out->eye_pos, out->eye_nrm, &front);
    }
    out->color = vc;

    mat_vec4(st->projection, eye[0], eye[1], eye[2], eye[3], out->pos); /* raster.c:48-56 */

    const float *tm = st->texture;
    float t4[4];
    t4[0] = tm[0] * uv[0] + tm[4] * uv[1] + tm[8] * 0.0f + tm[12] * 1.0f;
    t4[1] = tm[1] * uv[0] + tm[5] * uv[1] + tm[9] * 0.0f + tm[13] * 1.0f;
    t4[3] = tm[3] * uv[0] + tm[7] * uv[1] + tm[11] * 0.0f + tm[15] * 1.0f;
    if (t4[3] != 0.0f && t4[3] != 1.0f) {
        out->uv[0] = t4[0] / t4[3];
        out->uv[1] = t4[1] / t4[3];
    } else {
        out->uv[0] = t4[0];
        out->uv[1] = t4[1];
    }
}

/* glColor4f sanitising applied to colour-array elements (gl_api.c:708-722) */
static col4 sanitize_color(const float *c)
{
    float r = c[0], g = c[1], b = c[2], a = c[3];
    if (isnan(r) || isinf(r)) r = 0.0f;
    if (isnan(g) || isinf(g)) g = 0.0f;
    if (isnan(b) || isinf(b)) b = 0.0f;
    if (isnan(a) || isinf(a)) a = 1.0f;
    col4 o = { clamp01(r), clamp01(g), clamp01(b), clamp01(a) };
    return o;
}

/* get_array_element (gl_api.c:1758-1797) against a buffer-object mirror */
static void fetch_attrib(const mtgl_dev *dev, const mtgl_attrib *a, int32_t index, float *out, int want)
{
    const obuf *b = (a->buffer < O_MAX_BUFFERS) ? &dev->buf[a->buffer] : NULL;
    if (!b || !b->data || index < 0) {
        for (int i = 0; i < want; i++) out[i] = (i < 3) ? 0.0f : 1.0f;
        return;
    }
    const uint8_t *p = b->data + a->offset + (uint64_t)index * a->stride;
    for (int i = 0; i < (int)a->size && i < want; i++) {
        if (a->type == MTGL_TYPE_F32) { float f; memcpy(&f, p + 4 * i, 4); out[i] = f; }
        else out[i] = p[i] / 255.0f;
    }
    for (int i = a->size; i < want; i++) out[i] = (i == 3) ? 1.0f : 0.0f;
}

/* ------------------------------------------------------------------------------------------ */
/* texture sampling (textures.c:272-557)                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const uint32_t *px; int32_t w, h;
    const uint32_t *px1; int32_t w1, h1;
    uint32_t min_f, mag_f, wrap_s, wrap_t;
} sampler;

static uint32_t texel_wrapped(const uint32_t *px, int32_t w, int32_t h, uint32_t ws, uint32_t wt, int32_t x, int32_t y)
{ /* textures.c:272-291 and 357-376 */
    if (ws == T_REPEAT) x = ((x % w) + w) % w; else { if (x < 0) x = 0; else if (x >= w) x = w - 1; }
    if (wt == T_REPEAT) y = ((y % h) + h) % h; else { if (y < 0) y = 0; else if (y >= h) y = h - 1; }
    return px[y * w + x];
}

static uint32_t bilinear(uint32_t c00, uint32_t c10, uint32_t c01, uint32_t c11, float fx, float fy)
{ /* textures.c:294-307: result truncated back to RGBA8 */
    col4 top = col_lerp(col_unpack(c00), col_unpack(c10), fx);
    col4 bot = col_lerp(col_unpack(c01), col_unpack(c11), fx);
    return col_pack(col_lerp(top, bot, fy));
}

static uint32_t sample_level(const uint32_t *px, int32_t w, int32_t h, uint32_t ws, uint32_t wt, float u, float v, int linear)
{ /* textures.c:379-410 (level 0), 421-450 (level 1), 524-556 */
    float tx = u * w - 0.5f;
    float ty = v * h - 0.5f;
    if (linear) {
        int32_t x0 = (int32_t)floorf(tx), y0 = (int32_t)floorf(ty);
        float fx = tx - x0, fy = ty - y0;
        return bilinear(texel_wrapped(px, w, h, ws, wt, x0, y0), texel_wrapped(px, w, h, ws, wt, x0 + 1, y0),
                        texel_wrapped(px, w, h, ws, wt, x0, y0 + 1), texel_wrapped(px, w, h, ws, wt, x0 + 1, y0 + 1),
                        fx, fy);
    }
    int32_t x = (int32_t)floorf(tx + 0.5f), y = (int32_t)floorf(ty + 0.5f);
    if (x < 0) x = 0;
    if (x >= w) x = w - 1;
    if (y < 0) y = 0;
    if (y >= h) y = h - 1;
    return px[y * w + x];
}

static uint32_t sample_mip1(const sampler *s, float u, float v, uint32_t filter)
{ /* textures.c:413-451; a level that cannot be generated samples as opaque white (416-419) */
    if (!s->px1) return 0xFFFFFFFFu;
    int linear = (filter == T_LINEAR || filter == T_LINEAR_MIPMAP_NEAREST || filter == T_LINEAR_MIPMAP_LINEAR);
    return sample_level(s->px1, s->w1, s->h1, s->wrap_s, s->wrap_t, u, v, linear);
}

static uint32_t sample_lod(const sampler *s, float u, float v, float lod)
{ /* texture_sample_lod, textures.c:457-557 */
    if (s->wrap_s == T_REPEAT) { u = u - (float)(int)u; if (u < 0) u += 1.0f; }
    else { if (u < 0.0f) u = 0.0f; if (u > 1.0f) u = 1.0f; }
    if (s->wrap_t == T_REPEAT) { v = v - (float)(int)v; if (v < 0) v += 1.0f; }
    else { if (v < 0.0f) v = 0.0f; if (v > 1.0f) v = 1.0f; }

    uint32_t filter = (lod > 0.0f) ? s->min_f : s->mag_f;
    if (filter == T_NEAREST_MIPMAP_NEAREST || filter == T_LINEAR_MIPMAP_NEAREST) {
        if (lod >= 0.5f) return sample_mip1(s, u, v, filter);
        filter = (filter == T_NEAREST_MIPMAP_NEAREST) ? T_NEAREST : T_LINEAR;
    } else if (filter == T_NEAREST_MIPMAP_LINEAR || filter == T_LINEAR_MIPMAP_LINEAR) {
        if (lod > 0.0f) {
            float cl = (lod > 1.0f) ? 1.0f : lod;
            int base_linear = (filter != T_NEAREST_MIPMAP_LINEAR);
            uint32_t c0 = sample_level(s->px, s->w, s->h, s->wrap_s, s->wrap_t, u, v, base_linear);
            uint32_t c1 = sample_mip1(s, u, v, filter);
            return col_pack(col_lerp(col_unpack(c0), col_unpack(c1), cl));
        }
        filter = (filter == T_NEAREST_MIPMAP_LINEAR) ? T_NEAREST : T_LINEAR;
    }
    return sample_level(s->px, s->w, s->h, s->wrap_s, s->wrap_t, u, v, filter == T_LINEAR);
}

/* ------------------------------------------------------------------------------------------ */
/* per-fragment tests (raster.c:344-448)                                                      */
/* ------------------------------------------------------------------------------------------ */
static int cmp_f(uint32_t func, float a, float b)
{
    switch (func) {
    case T_NEVER: return 0;
    case T_LESS: return a < b;
    case T_EQUAL: return a == b;
    case T_LEQUAL: return a <= b;
    case T_GREATER: return a > b;
    case T_NOTEQUAL: return a != b;
    case T_GEQUAL: return a >= b;
    default: return 1;
    }
}

static int cmp_i(uint32_t func, int32_t a, int32_t b)
{
    switch (func) {
    case T_NEVER: return 0;
    case T_LESS: return a < b;
    case T_EQUAL: return a == b;
    case T_LEQUAL: return a <= b;
    case T_GREATER: return a > b;
    case T_NOTEQUAL: return a != b;
    case T_GEQUAL: return a >= b;
    default: return 1;
    }
}

static uint8_t stencil_apply(uint32_t op, uint8_t v, int32_t ref) /* raster.c:425-438 */
{
    switch (op) {
    case T_KEEP: return v;
    case T_ZERO: return 0;
    case T_REPLACE: return (uint8_t)(ref & 0xFF);
    case T_INCR: return v < 255 ? (uint8_t)(v + 1) : 255;
    case T_INCR_WRAP: return (uint8_t)(v + 1);
    case T_DECR: return v > 0 ? (uint8_t)(v - 1) : 0;
    case T_DECR_WRAP: return (uint8_t)(v - 1);
    case T_INVERT: return (uint8_t)~v;
    default: return v;
    }
}

static col4 blend_factor(uint32_t f, col4 s, col4 d) /* raster.c:360-379 */
{
    col4 one = { 1, 1, 1, 1 };
    switch (f) {
    case T_ZERO: { col4 z = { 0, 0, 0, 0 }; return z; }
    case T_ONE: return one;
    case T_SRC_COLOR: return s;
    case T_ONE_MINUS_SRC_COLOR: { col4 o = { 1 - s.r, 1 - s.g, 1 - s.b, 1 - s.a }; return o; }
    case T_DST_COLOR: return d;
    case T_ONE_MINUS_DST_COLOR: { col4 o = { 1 - d.r, 1 - d.g, 1 - d.b, 1 - d.a }; return o; }
    case T_SRC_ALPHA: { col4 o = { s.a, s.a, s.a, s.a }; return o; }
    case T_ONE_MINUS_SRC_ALPHA: { col4 o = { 1 - s.a, 1 - s.a, 1 - s.a, 1 - s.a }; return o; }
    case T_DST_ALPHA: { col4 o = { d.a, d.a, d.a, d.a }; return o; }
    case T_ONE_MINUS_DST_ALPHA: { col4 o = { 1 - d.a, 1 - d.a, 1 - d.a, 1 - d.a }; return o; }
    case T_SRC_ALPHA_SATURATE: { float k = (s.a < (1 - d.a)) ? s.a : (1 - d.a); col4 o = { k, k, k, 1 }; return o; }
    default: return one;
    }
}

static void put_color_masked(mtgl_dev *dev, const mtgl_state *st, int32_t x, int32_t y, col4 c) /* raster.c:20-45 */
{
    uint32_t *p = &dev->color[(size_t)y * dev->width + x];
    uint32_t m = st->color_mask & 0xF;
    if (m == 0xF) { *p = col_pack(c); return; }
    if (m == 0) return;
    col4 d = col_unpack(*p);
    if (m & 1) d.r = c.r;
    if (m & 2) d.g = c.g;
    if (m & 4) d.b = c.b;
    if (m & 8) d.a = c.a;
    *p = col_pack(d);
}

static void put_stencil_masked(mtgl_dev *dev, const mtgl_state *st, size_t idx, uint8_t nv) /* raster.c:441-448 */
{
    uint8_t m = (uint8_t)(st->stencil_writemask & 0xFF);
    dev->stencil[idx] = (uint8_t)((dev->stencil[idx] & ~m) | (nv & m));
}

static float fog_factor(const mtgl_state *st, float c) /* raster.c:677-701 */
{
    float f;
    switch (st->fog_mode) {
    case T_LINEAR_FOG: f = (st->fog_end != st->fog_start) ? (st->fog_end - c) / (st->fog_end - st->fog_start) : 1.0f; break;
    case T_EXP: f = expf(-st->fog_density * c); break;
    case T_EXP2: { float d = st->fog_density * c; f = expf(-d * d); break; }
    default: f = 1.0f; break;
    }
    if (f < 0.0f) f = 0.0f;
    if (f > 1.0f) f = 1.0f;
    return f;
}

/* ------------------------------------------------------------------------------------------ */
/* triangle rasterisation: rasterize_triangle_smooth (raster.c:451-725)                       */
/* ------------------------------------------------------------------------------------------ */
static float edge_fn(float ax, float ay, float bx, float by, float px, float py) /* raster.c:299-302 */
{
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

static int imin3(int a, int b, int c) { int m = a < b ? a : b; return m < c ? m : c; }
static int imax3(int a, int b, int c) { int m = a > b ? a : b; return m > c ? m : c; }

static void raster_triangle(mtgl_dev *dev, const mtgl_state *st, const overt *v0, const overt *v1, const overt *v2,
                            const int32_t *sx, const int32_t *sy, int back_facing)
{
    int32_t x0 = sx[0], y0 = sy[0], x1 = sx[1], y1 = sy[1], x2 = sx[2], y2 = sy[2];
    int32_t minX = imin3(x0, x1, x2), minY = imin3(y0, y1, y2);
    int32_t maxX = imax3(x0, x1, x2), maxY = imax3(y0, y1, y2);
    const int32_t *vp = st->viewport, *sc = st->scissor;
    if (minX < vp[0]) minX = vp[0];
    if (minY < vp[1]) minY = vp[1];
    if (maxX >= vp[0] + vp[2]) maxX = vp[0] + vp[2] - 1;
    if (maxY >= vp[1] + vp[3]) maxY = vp[1] + vp[3] - 1;
    if (st->caps & MTGL_CAP_SCISSOR_TEST) {
        if (minX < sc[0]) minX = sc[0];
        if (minY < sc[1]) minY = sc[1];
        if (maxX >= sc[0] + sc[2]) maxX = sc[0] + sc[2] - 1;
        if (maxY >= sc[1] + sc[3]) maxY = sc[1] + sc[3] - 1;
    }
    if (minX > maxX || minY > maxY) return;

    float area = edge_fn((float)x0, (float)y0, (float)x1, (float)y1, (float)x2, (float)y2);
    if (fabsf(area) < 0.5f) return;
    float inv_area = 1.0f / area;

    int depth_on = (st->caps & MTGL_CAP_DEPTH_TEST) != 0;
    int stencil_on = (st->caps & MTGL_CAP_STENCIL_TEST) != 0;
    int persp = (st->perspective_hint != T_FASTEST);

    sampler smp;
    int textured = 0;
    if ((st->caps & MTGL_CAP_TEXTURE_2D) && st->texture_id != 0 && st->texture_id < O_MAX_OBJECTS &&
        dev->tex[st->texture_id].px) {
        const otex *t = &dev->tex[st->texture_id];
        smp.px = t->px; smp.w = t->w; smp.h = t->h;
        smp.px1 = t->px1; smp.w1 = t->w1; smp.h1 = t->h1;
        smp.min_f = st->tex_min_filter; smp.mag_f = st->tex_mag_filter;
        smp.wrap_s = st->tex_wrap_s; smp.wrap_t = st->tex_wrap_t;
        textured = 1;
    }

    float z0 = v0->pos[2], z1 = v1->pos[2], z2 = v2->pos[2];
    float w0 = v0->pos[3], w1 = v1->pos[3], w2 = v2->pos[3]; /* 1/w after perspective_divide */
    float u0w = v0->uv[0] * w0, v0w = v0->uv[1] * w0;
    float u1w = v1->uv[0] * w1, v1w = v1->uv[1] * w1;
    float u2w = v2->uv[0] * w2, v2w = v2->uv[1] * w2;

    float lod = 0.0f; /* raster.c:505-529: one LOD per triangle */
    if (textured) {
        float screen_area = fabsf(area) * 0.5f;
        float du1 = (v1->uv[0] - v0->uv[0]) * smp.w, dv1 = (v1->uv[1] - v0->uv[1]) * smp.h;
        float du2 = (v2->uv[0] - v0->uv[0]) * smp.w, dv2 = (v2->uv[1] - v0->uv[1]) * smp.h;
        float texel_area = fabsf(du1 * dv2 - du2 * dv1) * 0.5f;
        if (screen_area > 0.0f) {
            float tpp = texel_area / screen_area;
            if (tpp > 0.0f) {
                lod = log2f(tpp) * 0.5f;
                if (lod < 0.0f) lod = 0.0f;
            }
        }
    }

    if (minX < 0) minX = 0;             /* framebuffer accessors are bounds checked (framebuffer.h:92-134); */
    if (minY < 0) minY = 0;             /* out-of-range pixels can neither change nor read state that matters */
    if (maxX >= dev->width) maxX = dev->width - 1;
    if (maxY >= dev->height) maxY = dev->height - 1;
    if (minY < dev->band_y0) minY = dev->band_y0;
    if (maxY >= dev->band_y1) maxY = dev->band_y1 - 1;

    for (int32_t y = minY; y <= maxY; y++) {
        for (int32_t x = minX; x <= maxX; x++) {
            float e0 = edge_fn((float)x1, (float)y1, (float)x2, (float)y2, (float)x, (float)y);
            float e1 = edge_fn((float)x2, (float)y2, (float)x0, (float)y0, (float)x, (float)y);
            float e2 = edge_fn((float)x0, (float)y0, (float)x1, (float)y1, (float)x, (float)y);
            if (!((area > 0 && e0 >= 0 && e1 >= 0 && e2 >= 0) || (area < 0 && e0 <= 0 && e1 <= 0 && e2 <= 0))) continue;

            dev->frag_covered++;
            float b0 = e0 * inv_area, b1 = e1 * inv_area, b2 = e2 * inv_area;
            float z = b0 * z0 + b1 * z1 + b2 * z2;
            float depth = (float)((z + 1.0f) * 0.5f * (st->depth_far - st->depth_near) + st->depth_near);
            size_t idx = (size_t)y * dev->width + x;

            uint8_t sval = 0;
            if (stencil_on) {
                sval = dev->stencil[idx];
                int32_t mref = (int32_t)((uint32_t)st->stencil_ref & st->stencil_mask);
                int32_t mval = (int32_t)((uint32_t)sval & st->stencil_mask);
                if (!cmp_i(st->stencil_func, mref, mval)) {
                    put_stencil_masked(dev, st, idx, stencil_apply(st->stencil_fail, sval, st->stencil_ref));
                    continue;
                }
            }
            if (depth_on) {
                if (!cmp_f(st->depth_func, depth, dev->depth[idx])) {
                    if (stencil_on) put_stencil_masked(dev, st, idx, stencil_apply(st->stencil_zfail, sval, st->stencil_ref));
                    continue;
                }
            }
            if (stencil_on) put_stencil_masked(dev, st, idx, stencil_apply(st->stencil_zpass, sval, st->stencil_ref));
            dev->frag_tested++;

            col4 c;
            if (st->shade_model == T_FLAT) c = v2->color;
            else {
                c.r = v0->color.r * b0 + v1->color.r * b1 + v2->color.r * b2;
                c.g = v0->color.g * b0 + v1->color.g * b1 + v2->color.g * b2;
                c.b = v0->color.b * b0 + v1->color.b * b1 + v2->color.b * b2;
                c.a = v0->color.a * b0 + v1->color.a * b1 + v2->color.a * b2;
            }

            if (st->caps & MTGL_CAP_LIGHTING) { /* raster.c:592-615 */
                int phong = (st->shade_model == T_PHONG);
                int flip = back_facing && st->light_model_two_side;
                if (phong || flip) {
                    float ep[3], en[3];
                    for (int k = 0; k < 3; k++) {
                        ep[k] = v0->eye_pos[k] * b0 + v1->eye_pos[k] * b1 + v2->eye_pos[k] * b2;
                        en[k] = v0->eye_nrm[k] * b0 + v1->eye_nrm[k] * b1 + v2->eye_nrm[k] * b2;
                    }
                    const mtgl_material *mat = &st->material_front;
                    if (flip) {
                        en[0] *= -1.0f; en[1] *= -1.0f; en[2] *= -1.0f;
                        mat = &st->material_back;
                    }
                    c = light_vertex(st, ep, en, mat);
                }
            }

            if (textured) {
                float u, v;
                if (persp) {
                    float uw = b0 * u0w + b1 * u1w + b2 * u2w;
                    float vw = b0 * v0w + b1 * v1w + b2 * v2w;
                    float ow = b0 * w0 + b1 * w1 + b2 * w2;
                    float w = 1.0f / ow;
                    u = uw * w;
                    v = vw * w;
                } else {
                    u = b0 * v0->uv[0] + b1 * v1->uv[0] + b2 * v2->uv[0];
                    v = b0 * v0->uv[1] + b1 * v1->uv[1] + b2 * v2->uv[1];
                }
                col4 t = col_unpack(sample_lod(&smp, u, v, lod));
                if ((st->caps & MTGL_CAP_ALPHA_TEST) && !cmp_f(st->alpha_func, t.a, st->alpha_ref)) continue;
                switch (st->tex_env_mode) { /* raster.c:646-668 */
                case T_REPLACE: c = t; break;
                case T_DECAL: c = col_lerp_rgb(c, t, t.a); break;
                case T_BLEND_ENV: {
                    col4 e = col_from(st->tex_env_color);
                    col4 o = { c.r * (1.0f - t.r) + e.r * t.r, c.g * (1.0f - t.g) + e.g * t.g,
                               c.b * (1.0f - t.b) + e.b * t.b, c.a * t.a };
                    c = o;
                    break;
                }
                case T_ADD: { col4 o = { c.r + t.r, c.g + t.g, c.b + t.b, c.a * t.a }; c = o; break; }
                default: { col4 o = { c.r * t.r, c.g * t.g, c.b * t.b, c.a * t.a }; c = o; break; }
                }
            }

            if (st->caps & MTGL_CAP_FOG) { /* raster.c:672-705: result alpha is the FOG colour's alpha */
                float fc = b0 * v0->eye_z + b1 * v1->eye_z + b2 * v2->eye_z;
                c = col_lerp_rgb(col_from(st->fog_color), c, fog_factor(st, fc));
            }

            if (depth_on && st->depth_mask) dev->depth[idx] = depth;

            if (st->caps & MTGL_CAP_BLEND) { /* raster.c:382-387 */
                col4 d = col_unpack(dev->color[idx]);
                col4 sf = blend_factor(st->blend_src, c, d), df = blend_factor(st->blend_dst, c, d);
                col4 o = { c.r * sf.r + d.r * df.r, c.g * sf.g + d.g * df.g, c.b * sf.b + d.b * df.b, c.a * sf.a + d.a * df.a };
                c = col_clamp(o);
            }
            c = col_clamp(c);
            dev->frag_shaded++;
            put_color_masked(dev, st, x, y, c);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* clip, divide, fan, cull: render_triangle (raster.c:901-958), clipping.h:50-126              */
/* ------------------------------------------------------------------------------------------ */
static float lerp1(float a, float b, float t) { return a + t * (b - a); } /* graphics.h:369-371 */

static overt vert_lerp(const overt *a, const overt *b, float t) /* vertex_lerp, graphics.h:465-475 */
{
    overt o;
    for (int k = 0; k < 4; k++) o.pos[k] = lerp1(a->pos[k], b->pos[k], t);
    o.color = col_lerp(a->color, b->color, t);
    o.uv[0] = lerp1(a->uv[0], b->uv[0], t);
    o.uv[1] = lerp1(a->uv[1], b->uv[1], t);
    o.eye_z = lerp1(a->eye_z, b->eye_z, t);
    for (int k = 0; k < 3; k++) {
        o.eye_pos[k] = lerp1(a->eye_pos[k], b->eye_pos[k], t);
        o.eye_nrm[k] = lerp1(a->eye_nrm[k], b->eye_nrm[k], t);
    }
    return o;
}

static float plane_dist(const float *p, int plane) /* clipping.h:27-32; order near,far,left,right,bottom,top */
{
    switch (plane) {
    case 0: return p[2] + p[3];
    case 1: return p[3] - p[2];
    case 2: return p[0] + p[3];
    case 3: return p[3] - p[0];
    case 4: return p[1] + p[3];
    default: return p[3] - p[1];
    }
}

static void plane_snap(float *p, int plane) /* clipping.h:37-47 */
{
    switch (plane) {
    case 0: p[2] = -p[3]; break;
    case 1: p[2] = p[3]; break;
    case 2: p[0] = -p[3]; break;
    case 3: p[0] = p[3]; break;
    case 4: p[1] = -p[3]; break;
    default: p[1] = p[3]; break;
    }
}

#define O_MAX_CLIP 12

static int clip_plane(const overt *in, int n, overt *out, int plane) /* clipping.h:50-99 */
{
    if (n == 0) return 0;
    int m = 0;
    const overt *prev = &in[n - 1];
    float pd = plane_dist(prev->pos, plane);
    for (int i = 0; i < n; i++) {
        const overt *cur = &in[i];
        float cd = plane_dist(cur->pos, plane);
        if (pd >= 0) {
            if (cd >= 0) out[m++] = *cur;
            else {
                float den = pd - cd;
                if (fabsf(den) > 1e-10f) {
                    out[m] = vert_lerp(prev, cur, pd / den);
                    plane_snap(out[m].pos, plane);
                    m++;
                }
            }
        } else if (cd >= 0) {
            float den = pd - cd;
            if (fabsf(den) > 1e-10f) {
                out[m] = vert_lerp(prev, cur, pd / den);
                plane_snap(out[m].pos, plane);
                m++;
            }
            out[m++] = *cur;
        }
        prev = cur;
        pd = cd;
    }
    return m;
}

static void to_screen(const mtgl_state *st, float x, float y, int32_t *sx, int32_t *sy) /* raster.c:59-63 */
{
    *sx = (int32_t)((x + 1.0f) * 0.5f * st->viewport[2] + st->viewport[0]);
    *sy = (int32_t)((1.0f - y) * 0.5f * st->viewport[3] + st->viewport[1]);
}

/* ------------------------------------------------------------------------------------------ */
/* points and lines (raster.c:65-296, 776-898, 1019-1196; clipping.h:128-229)                  */
/* ------------------------------------------------------------------------------------------ */
static sampler bound_sampler(const mtgl_dev *dev, const mtgl_state *st, int *textured)
{
    sampler smp;
    memset(&smp, 0, sizeof smp);
    *textured = 0;
    if ((st->caps & MTGL_CAP_TEXTURE_2D) && st->texture_id != 0 && st->texture_id < O_MAX_OBJECTS && dev->tex[st->texture_id].px) {
        const otex *t = &dev->tex[st->texture_id];
        smp.px = t->px; smp.w = t->w; smp.h = t->h; smp.px1 = t->px1; smp.w1 = t->w1; smp.h1 = t->h1;
        smp.min_f = st->tex_min_filter; smp.mag_f = st->tex_mag_filter; smp.wrap_s = st->tex_wrap_s; smp.wrap_t = st->tex_wrap_t;
        *textured = 1;
    }
    return smp;
}

/* depth test, blend, depth write, masked colour write of one line / point pixel: write_line_pixel (raster.c:66-104)
 * and the pixel loops of flush_points (1123-1161) / draw_point_at_screen (795-843).  No stencil on these paths. */
static void put_simple_pixel(mtgl_dev *dev, const mtgl_state *st, int32_t px, int32_t py, float depth, col4 c)
{
    if (px < 0 || px >= dev->width || py < 0 || py >= dev->height) return;
    if (py < dev->band_y0 || py >= dev->band_y1) return;
    if (st->caps & MTGL_CAP_SCISSOR_TEST) {
        if (px < st->scissor[0] || px >= st->scissor[0] + st->scissor[2] || py < st->scissor[1] || py >= st->scissor[1] + st->scissor[3]) return;
    }
    size_t idx = (size_t)py * dev->width + px;
    int depth_on = (st->caps & MTGL_CAP_DEPTH_TEST) != 0;
    if (depth_on && !cmp_f(st->depth_func, depth, dev->depth[idx])) return;
    dev->frag_tested++;
    if (st->caps & MTGL_CAP_BLEND) {
        col4 d = col_unpack(dev->color[idx]);
        col4 sf = blend_factor(st->blend_src, c, d), df = blend_factor(st->blend_dst, c, d);
        col4 o = { c.r * sf.r + d.r * df.r, c.g * sf.g + d.g * df.g, c.b * sf.b + d.b * df.b, c.a * sf.a + d.a * df.a };
        c = col_clamp(o);
    }
    if (depth_on && st->depth_mask) dev->depth[idx] = depth;
    dev->frag_shaded++;
    put_color_masked(dev, st, px, py, c);
}

static float simple_depth(const mtgl_state *st, float z) /* raster.c:161, 793, 1065 */
{
    return (float)((z + 1.0f) * 0.5f * (st->depth_far - st->depth_near) + st->depth_near);
}

/* fog for lines and points: the coordinate is NEGATED once more (raster.c:189, 809, 1093) */
static col4 simple_fog(const mtgl_state *st, col4 c, float eye_z)
{
    if (!(st->caps & MTGL_CAP_FOG)) return c;
    float fc = -eye_z, f = 1.0f;
    switch (st->fog_mode) {
    case T_LINEAR_FOG: if (st->fog_end != st->fog_start) f = (st->fog_end - fc) / (st->fog_end - st->fog_start); break;
    case T_EXP: f = expf(-st->fog_density * fc); break;
    case T_EXP2: { float d = st->fog_density * fc; f = expf(-d * d); break; }
    default: break;
    }
    if (f < 0.0f) f = 0.0f;
    if (f > 1.0f) f = 1.0f;
    return col_lerp_rgb(col_from(st->fog_color), c, f);
}

/* draw_line_full (raster.c:107-241).  One deliberate deviation: when the alpha test rejects a pixel the reference
 * executes `continue` inside `for (;;)` and never advances (an infinite loop); here the pixel is
 * skipped and the walk goes on. */
static void draw_line(mtgl_dev *dev, const mtgl_state *st, int32_t x0, int32_t y0, float z0, int32_t x1, int32_t y1, float z1,
                      col4 c0, col4 c1, float ez0, float ez1, const float *uv0, const float *uv1)
{
    int32_t dx = x1 - x0, dy = y1 - y0;
    int32_t adx = dx < 0 ? -dx : dx, ady = dy < 0 ? -dy : dy;
    int32_t sx = dx < 0 ? -1 : 1, sy = dy < 0 ? -1 : 1;
    int32_t err = adx - ady;
    int32_t total = (adx > ady) ? adx : ady;
    if (total == 0) total = 1;
    int32_t step = 0;
    int lw = (int)(st->line_width + 0.5f);
    if (lw < 1) lw = 1;
    int half = lw / 2;
    int ex = (adx > ady) ? 0 : 1, ey = (adx > ady) ? 1 : 0;
    int textured;
    sampler smp = bound_sampler(dev, st, &textured);
    int32_t cx = x0, cy = y0;
    for (;;) {
        float t = (float)step / (float)total;
        float z = z0 + t * (z1 - z0);
        float depth = simple_depth(st, z);
        col4 c = col_lerp(c0, c1, t);
        int keep = 1;
        if (textured) {
            float u = uv0[0] + t * (uv1[0] - uv0[0]), v = uv0[1] + t * (uv1[1] - uv0[1]);
            col4 tc = col_unpack(sample_lod(&smp, u, v, 0.0f));
            if ((st->caps & MTGL_CAP_ALPHA_TEST) && !cmp_f(st->alpha_func, tc.a, st->alpha_ref)) keep = 0;
            else { col4 o = { c.r * tc.r, c.g * tc.g, c.b * tc.b, c.a * tc.a }; c = o; }
        } else if ((st->caps & MTGL_CAP_ALPHA_TEST) && !cmp_f(st->alpha_func, c.a, st->alpha_ref)) keep = 0;
        if (keep) {
            c = simple_fog(st, c, ez0 + t * (ez1 - ez0));
            dev->frag_covered += (uint64_t)lw;
            if (lw == 1) put_simple_pixel(dev, st, cx, cy, depth, c);
            else for (int w = -half; w < lw - half; w++) put_simple_pixel(dev, st, cx + w * ex, cy + w * ey, depth, c);
        }
        if (cx == x1 && cy == y1) break;
        int32_t e2 = err * 2;
        if (e2 > -ady) { err -= ady; cx += sx; }
        if (e2 < adx) { err += adx; cy += sy; }
        step++;
    }
}

static int outcode(const float *p) /* clipping.h:138-148 */
{
    int c = 0;
    if (p[0] < -p[3]) c |= 1; else if (p[0] > p[3]) c |= 2;
    if (p[1] < -p[3]) c |= 4; else if (p[1] > p[3]) c |= 8;
    if (p[2] < -p[3]) c |= 16; else if (p[2] > p[3]) c |= 32;
    return c;
}

static int clip_segment(overt *v0, overt *v1) /* clip_line, clipping.h:152-229 (Cohen-Sutherland with snap) */
{
    int c0 = outcode(v0->pos), c1 = outcode(v1->pos);
    for (;;) {
        if (!(c0 | c1)) return 1;
        if (c0 & c1) return 0;
        int co = c0 ? c0 : c1;
        const float *p0 = v0->pos, *p1 = v1->pos;
        float d0, d1;
        int plane;
        if (co & 1) { d0 = p0[0] + p0[3]; d1 = p1[0] + p1[3]; plane = 2; }
        else if (co & 2) { d0 = p0[3] - p0[0]; d1 = p1[3] - p1[0]; plane = 3; }
        else if (co & 4) { d0 = p0[1] + p0[3]; d1 = p1[1] + p1[3]; plane = 4; }
        else if (co & 8) { d0 = p0[3] - p0[1]; d1 = p1[3] - p1[1]; plane = 5; }
        else if (co & 16) { d0 = p0[2] + p0[3]; d1 = p1[2] + p1[3]; plane = 0; }
        else { d0 = p0[3] - p0[2]; d1 = p1[3] - p1[2]; plane = 1; }
        float den = d0 - d1;
        if (fabsf(den) < 1e-10f) return 0;
        overt cl = vert_lerp(v0, v1, d0 / den);
        plane_snap(cl.pos, plane);
        if (co == c0) { *v0 = cl; c0 = outcode(v0->pos); }
        else { *v1 = cl; c1 = outcode(v1->pos); }
    }
}

static void line_segment(mtgl_dev *dev, const mtgl_state *st, const overt *a, const overt *b) /* draw_line_segment, raster.c:244-285 */
{
    overt v0 = *a, v1 = *b;
    if (!clip_segment(&v0, &v1)) return;
    float z0, z1;
    if (fabsf(v0.pos[3]) >= 1e-6f) { float iw = 1.0f / v0.pos[3]; v0.pos[0] *= iw; v0.pos[1] *= iw; z0 = v0.pos[2] * iw; }
    else { v0.pos[0] = 0.0f; v0.pos[1] = 0.0f; z0 = 0.0f; }
    if (fabsf(v1.pos[3]) >= 1e-6f) { float iw = 1.0f / v1.pos[3]; v1.pos[0] *= iw; v1.pos[1] *= iw; z1 = v1.pos[2] * iw; }
    else { v1.pos[0] = 0.0f; v1.pos[1] = 0.0f; z1 = 0.0f; }
    int32_t x0, y0, x1, y1;
    to_screen(st, v0.pos[0], v0.pos[1], &x0, &y0);
    to_screen(st, v1.pos[0], v1.pos[1], &x1, &y1);
    draw_line(dev, st, x0, y0, z0, x1, y1, z1, v0.color, v1.color, v0.eye_z, v1.eye_z, v0.uv, v1.uv);
}

static void points(mtgl_dev *dev, const mtgl_state *st, const overt *v, uint32_t n) /* flush_points, raster.c:1020-1164 */
{
    int ps = (int)(st->point_size + 0.5f);
    if (ps < 1) ps = 1;
    int half = ps / 2;
    int textured;
    sampler smp = bound_sampler(dev, st, &textured);
    for (uint32_t i = 0; i < n; i++) {
        const float *p = v[i].pos;
        if (p[0] < -p[3] || p[0] > p[3] || p[1] < -p[3] || p[1] > p[3] || p[2] < -p[3] || p[2] > p[3] || p[3] <= 0.0f) continue;
        float nx = p[0] / p[3], ny = p[1] / p[3], nz = p[2] / p[3];
        int32_t cx, cy;
        to_screen(st, nx, ny, &cx, &cy);
        float depth = simple_depth(st, nz);
        col4 c = v[i].color;
        if (textured) {
            col4 tc = col_unpack(sample_lod(&smp, v[i].uv[0], v[i].uv[1], 0.0f));
            if ((st->caps & MTGL_CAP_ALPHA_TEST) && !cmp_f(st->alpha_func, tc.a, st->alpha_ref)) continue;
            col4 o = { c.r * tc.r, c.g * tc.g, c.b * tc.b, c.a * tc.a };
            c = o;
        } else if ((st->caps & MTGL_CAP_ALPHA_TEST) && !cmp_f(st->alpha_func, c.a, st->alpha_ref)) continue;
        c = simple_fog(st, c, v[i].eye_z);
        dev->frag_covered += (uint64_t)ps * ps;
        for (int py = cy - half; py < cy - half + ps; py++)
            for (int px = cx - half; px < cx - half + ps; px++) put_simple_pixel(dev, st, px, py, depth, c);
    }
}

/* polygon modes GL_LINE / GL_POINT (draw_triangle_wireframe / draw_triangle_points, raster.c:847-898) */
static void triangle_outline(mtgl_dev *dev, const mtgl_state *st, const overt *c0, const overt *c1, const overt *c2,
                             const int32_t *sx, const int32_t *sy, int as_points)
{
    col4 col[3] = { c0->color, c1->color, c2->color };
    const overt *cv[3] = { c0, c1, c2 };
    if ((st->caps & MTGL_CAP_LIGHTING) && st->shade_model == T_PHONG)
        for (int k = 0; k < 3; k++) col[k] = light_vertex(st, cv[k]->eye_pos, cv[k]->eye_nrm, &st->material_front);
    if (as_points) {
        for (int k = 0; k < 3; k++) {   /* draw_point_at_screen, raster.c:777-844: depth test before the alpha test, vertex alpha */
            if ((st->caps & MTGL_CAP_ALPHA_TEST) && !cmp_f(st->alpha_func, col[k].a, st->alpha_ref)) {
                /* the reference tests depth first, but neither test has side effects before the writes */
                continue;
            }
            dev->frag_covered++;
            put_simple_pixel(dev, st, sx[k], sy[k], simple_depth(st, cv[k]->pos[2]), simple_fog(st, col[k], cv[k]->eye_z));
        }
        return;
    }
    for (int k = 0; k < 3; k++) {
        int n = (k + 1) % 3;
        draw_line(dev, st, sx[k], sy[k], cv[k]->pos[2], sx[n], sy[n], cv[n]->pos[2], col[k], col[n], cv[k]->eye_z, cv[n]->eye_z,
                  cv[k]->uv, cv[n]->uv);
    }
}

static void render_triangle(mtgl_dev *dev, const mtgl_state *st, const overt *a, const overt *b, const overt *c)
{
    overt t1[O_MAX_CLIP], t2[O_MAX_CLIP];
    overt tri[3] = { *a, *b, *c };
    int n = clip_plane(tri, 3, t1, 0);          /* clipping.h:106-126 */
    if (n) n = clip_plane(t1, n, t2, 1);
    if (n) n = clip_plane(t2, n, t1, 2);
    if (n) n = clip_plane(t1, n, t2, 3);
    if (n) n = clip_plane(t2, n, t1, 4);
    if (n) n = clip_plane(t1, n, t2, 5);
    if (n < 3) return;
    overt *cl = t2;
    dev->stats.triangles_in++;

    for (int j = 0; j < n; j++) { /* perspective_divide, raster.c:729-746 */
        float *p = cl[j].pos;
        if (fabsf(p[3]) < 1e-6f) { p[0] = 0.0f; p[1] = 0.0f; p[2] = 0.0f; p[3] = 1.0f; continue; }
        float iw = 1.0f / p[3];
        p[0] *= iw; p[1] *= iw; p[2] *= iw; p[3] = iw;
    }

    for (int j = 1; j + 1 < n; j++) {
        int32_t sx[3], sy[3];
        to_screen(st, cl[0].pos[0], cl[0].pos[1], &sx[0], &sy[0]);
        to_screen(st, cl[j].pos[0], cl[j].pos[1], &sx[1], &sy[1]);
        to_screen(st, cl[j + 1].pos[0], cl[j + 1].pos[1], &sx[2], &sy[2]);
        float sa = (float)(sx[1] - sx[0]) * (float)(sy[2] - sy[0]) - (float)(sx[2] - sx[0]) * (float)(sy[1] - sy[0]);
        if (st->caps & MTGL_CAP_CULL_FACE) { /* should_cull, raster.c:751-774 */
            int front = (st->front_face == T_CCW) ? (sa < 0) : (sa > 0);
            int cull = (st->cull_face_mode == T_FRONT) ? front : (st->cull_face_mode == T_BACK) ? !front : 1;
            if (cull) continue;
        }
        int back = (st->front_face == T_CCW) ? (sa >= 0) : (sa < 0);
        uint32_t pm = back ? st->polygon_mode_back : st->polygon_mode_front;
        if (pm == T_POINT || pm == T_LINE) { triangle_outline(dev, st, &cl[0], &cl[j], &cl[j + 1], sx, sy, pm == T_POINT); continue; }
        dev->stats.triangles_setup++;
        raster_triangle(dev, st, &cl[0], &cl[j], &cl[j + 1], sx, sy, back);
    }
}

/* primitive assembly: flush_* (raster.c:961-1017, 1199-1231) */
static void assemble(mtgl_dev *dev, const mtgl_state *st, uint32_t mode, const overt *v, uint32_t n)
{
    switch (mode) {
    case T_TRIANGLES:
        for (uint32_t i = 0; i + 2 < n; i += 3) render_triangle(dev, st, &v[i], &v[i + 1], &v[i + 2]);
        break;
    case T_QUADS:
        for (uint32_t i = 0; i + 3 < n; i += 4) {
            render_triangle(dev, st, &v[i], &v[i + 1], &v[i + 2]);
            render_triangle(dev, st, &v[i], &v[i + 2], &v[i + 3]);
        }
        break;
    case T_TRIANGLE_STRIP:
        for (uint32_t i = 0; i + 2 < n; i++) {
            if (i % 2 == 0) render_triangle(dev, st, &v[i], &v[i + 1], &v[i + 2]);
            else render_triangle(dev, st, &v[i + 1], &v[i], &v[i + 2]);
        }
        break;
    case T_TRIANGLE_FAN:
    case T_POLYGON:
        for (uint32_t i = 1; i + 1 < n; i++) render_triangle(dev, st, &v[0], &v[i], &v[i + 1]);
        break;
    case T_QUAD_STRIP:
        if (n < 4) break;
        for (uint32_t i = 0; i + 3 < n; i += 2) {
            render_triangle(dev, st, &v[i], &v[i + 1], &v[i + 3]);
            render_triangle(dev, st, &v[i], &v[i + 3], &v[i + 2]);
        }
        break;
    case T_POINTS:
        points(dev, st, v, n);
        break;
    case T_LINES:                               /* raster.c:288-296 */
        for (uint32_t i = 0; i + 1 < n; i += 2) line_segment(dev, st, &v[i], &v[i + 1]);
        break;
    case T_LINE_STRIP:                          /* raster.c:1167-1177 */
    case T_LINE_LOOP:                           /* raster.c:1180-1196 */
        if (n < 2) break;
        for (uint32_t i = 0; i + 1 < n; i++) line_segment(dev, st, &v[i], &v[i + 1]);
        if (mode == T_LINE_LOOP) line_segment(dev, st, &v[n - 1], &v[0]);
        break;
    default:
        break;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* C ABI                                                                                      */
/* ------------------------------------------------------------------------------------------ */
int mtgl_dev_abi_version(void) { return MTGL_DEV_ABI_VERSION; }

int mtgl_dev_create(int32_t width, int32_t height, int32_t device, mtgl_dev **out)
{
    (void)device;
    if (!out || width <= 0 || height <= 0 || width > 16384 || height > 16384) return MTGL_E_INVALID;
    mtgl_dev *d = (mtgl_dev *)calloc(1, sizeof *d);
    if (!d) return MTGL_E_OOM;
    size_t n = (size_t)width * height;
    d->width = width; d->height = height; d->band_y0 = 0; d->band_y1 = height;
    d->color = (uint32_t *)malloc(n * 4);
    d->depth = (float *)malloc(n * 4);
    d->stencil = (uint8_t *)malloc(n);
    if (!d->color || !d->depth || !d->stencil) { mtgl_dev_destroy(d); return MTGL_E_OOM; }
    *out = d;
    return MTGL_OK;
}

void mtgl_dev_destroy(mtgl_dev *d)
{
    if (!d) return;
    for (int i = 0; i < O_MAX_OBJECTS; i++) { free(d->tex[i].px); free(d->tex[i].px1); }
    for (unsigned i = 0; i < O_MAX_BUFFERS; i++) free(d->buf[i].data);
    free(d->color); free(d->depth); free(d->stencil);
    free(d);
}

int mtgl_dev_set_band(mtgl_dev *d, int32_t y0, int32_t y1)
{
    if (!d || y0 < 0 || y1 > d->height || y0 > y1) return MTGL_E_INVALID;
    d->band_y0 = y0; d->band_y1 = y1;
    return MTGL_OK;
}

int mtgl_dev_buffer_data(mtgl_dev *d, uint32_t id, uint64_t size, const void *data)
{
    if (!d || id == 0 || id >= O_MAX_BUFFERS) return MTGL_E_INVALID;
    free(d->buf[id].data);
    d->buf[id].data = NULL; d->buf[id].size = 0;
    if (size == 0) return MTGL_OK;
    d->buf[id].data = (uint8_t *)malloc(size);
    if (!d->buf[id].data) return MTGL_E_OOM;
    d->buf[id].size = size;
    if (data) memcpy(d->buf[id].data, data, size);
    return MTGL_OK;
}

int mtgl_dev_buffer_pointer(mtgl_dev *d, uint32_t id, void **ptr, uint64_t *size)
{
    if (!d || id == 0 || id >= O_MAX_BUFFERS) return MTGL_E_INVALID;
    if (ptr) *ptr = d->buf[id].data;
    if (size) *size = d->buf[id].size;
    return MTGL_OK;
}

int mtgl_dev_buffer_orphan(mtgl_dev *d, uint32_t id, uint64_t size, void **ptr)     /* nothing is in flight here: same storage, resized */
{
    if (!d || id == 0 || id >= O_MAX_BUFFERS || size == 0 || !ptr) return MTGL_E_INVALID;
    if (d->buf[id].size != size) { int rc = mtgl_dev_buffer_data(d, id, size, NULL); if (rc) return rc; }
    *ptr = d->buf[id].data;
    return MTGL_OK;
}

/* the pipelined forms (include/mtgl_dev.h) are the synchronous ones here: the oracle executes every call at once */
int mtgl_dev_buffer_data_pinned(mtgl_dev *d, uint32_t id, uint64_t size, const void *data) { return mtgl_dev_buffer_data(d, id, size, data); }
int mtgl_dev_read_color_async(mtgl_dev *d, int32_t y0, int32_t y1, uint32_t *color) { return mtgl_dev_read_framebuffer(d, y0, y1, color, NULL, NULL); }

int mtgl_dev_buffer_sub_data(mtgl_dev *d, uint32_t id, uint64_t offset, uint64_t size, const void *data)
{
    if (!d || id == 0 || id >= O_MAX_BUFFERS || !d->buf[id].data || !data) return MTGL_E_INVALID;
    if (offset + size > d->buf[id].size) return MTGL_E_INVALID;
    memcpy(d->buf[id].data + offset, data, size);
    return MTGL_OK;
}

int mtgl_dev_buffer_delete(mtgl_dev *d, uint32_t id)
{
    if (!d || id == 0 || id >= O_MAX_BUFFERS) return MTGL_E_INVALID;
    free(d->buf[id].data);
    d->buf[id].data = NULL; d->buf[id].size = 0;
    return MTGL_OK;
}

int mtgl_dev_buffer_read(mtgl_dev *d, uint32_t id, uint64_t offset, uint64_t size, void *out)
{
    if (!d || id == 0 || id >= O_MAX_BUFFERS || !d->buf[id].data || !out) return MTGL_E_INVALID;
    if (offset + size > d->buf[id].size) return MTGL_E_INVALID;
    memcpy(out, d->buf[id].data + offset, size);
    return MTGL_OK;
}

int mtgl_dev_texture_image(mtgl_dev *d, uint32_t id, int32_t w, int32_t h, const uint32_t *rgba8)
{
    if (!d || id == 0 || id >= O_MAX_OBJECTS || w <= 0 || h <= 0 || w > 2048 || h > 2048 || !rgba8) return MTGL_E_INVALID;
    otex *t = &d->tex[id];
    free(t->px); free(t->px1);
    memset(t, 0, sizeof *t);
    t->px = (uint32_t *)malloc((size_t)w * h * 4);
    if (!t->px) return MTGL_E_OOM;
    memcpy(t->px, rgba8, (size_t)w * h * 4);
    t->w = w; t->h = h;
    if (w >= 2 && h >= 2) { /* texture_generate_mip1, textures.c:311-354 */
        t->w1 = w / 2; t->h1 = h / 2;
        t->px1 = (uint32_t *)malloc((size_t)t->w1 * t->h1 * 4);
        if (!t->px1) return MTGL_E_OOM;
        for (int32_t y = 0; y < t->h1; y++)
            for (int32_t x = 0; x < t->w1; x++) {
                col4 a = col_unpack(t->px[(2 * y) * w + 2 * x]);
                col4 b = col_unpack(t->px[(2 * y) * w + 2 * x + 1]);
                col4 c = col_unpack(t->px[(2 * y + 1) * w + 2 * x]);
                col4 e = col_unpack(t->px[(2 * y + 1) * w + 2 * x + 1]);
                col4 s = { ((a.r + b.r) + c.r) + e.r, ((a.g + b.g) + c.g) + e.g, ((a.b + b.b) + c.b) + e.b,
                           ((a.a + b.a) + c.a) + e.a };
                s.r *= 0.25f; s.g *= 0.25f; s.b *= 0.25f; s.a *= 0.25f;
                t->px1[y * t->w1 + x] = col_pack(s);
            }
    }
    return MTGL_OK;
}

int mtgl_dev_texture_delete(mtgl_dev *d, uint32_t id)
{
    if (!d || id == 0 || id >= O_MAX_OBJECTS) return MTGL_E_INVALID;
    free(d->tex[id].px); free(d->tex[id].px1);
    memset(&d->tex[id], 0, sizeof d->tex[id]);
    return MTGL_OK;
}

static void do_clear(mtgl_dev *d, const mtgl_batch *b) /* glClear, gl_api.c:409-457 */
{
    int32_t x0 = b->clear_rect[0], y0 = b->clear_rect[1], x1 = b->clear_rect[2], y1 = b->clear_rect[3];
    if (y0 < d->band_y0) y0 = d->band_y0;
    if (y1 > d->band_y1) y1 = d->band_y1;
    for (int32_t y = y0; y < y1; y++)
        for (int32_t x = x0; x < x1; x++) {
            size_t i = (size_t)y * d->width + x;
            if (b->clear_mask & T_COLOR_BUFFER_BIT) d->color[i] = b->clear_color;
            if (b->clear_mask & T_DEPTH_BUFFER_BIT) d->depth[i] = b->clear_depth;
            if (b->clear_mask & T_STENCIL_BUFFER_BIT) d->stencil[i] = (uint8_t)b->clear_stencil;
        }
}

int mtgl_dev_submit(mtgl_dev *d, const mtgl_batch *b)
{
    if (!d || !b) return MTGL_E_INVALID;
    memset(&d->stats, 0, sizeof d->stats);
    if (b->clear_mask) do_clear(d, b);
    for (uint32_t di = 0; di < b->n_draws; di++) {
        const mtgl_draw *dr = &b->draws[di];
        if (dr->raster_state >= b->n_states || dr->count == 0) continue;
        overt *v = (overt *)malloc(sizeof(overt) * dr->count);
        if (!v) return MTGL_E_OOM;
        for (uint32_t i = 0; i < dr->count; i++) {
            if (dr->source == MTGL_SRC_STAGED) {
                const mtgl_in_vertex *iv = &b->vertices[dr->first_staged + i];
                vertex_stage(&b->states[iv->state], iv->position, col_from(iv->color), iv->texcoord, iv->normal, &v[i]);
            } else { /* glDrawArrays / glDrawElements attribute fetch, gl_api.c:1799-1941 */
                int32_t idx;
                if (dr->index_type) {
                    const uint8_t *ib = dr->index_buffer ? d->buf[dr->index_buffer].data : (const uint8_t *)b->blob;
                    ib += dr->index_offset;
                    uint32_t u;
                    if (dr->index_type == T_UNSIGNED_SHORT) { uint16_t s; memcpy(&s, ib + 2 * (size_t)i, 2); u = s; }
                    else if (dr->index_type == T_UNSIGNED_INT) memcpy(&u, ib + 4 * (size_t)i, 4);
                    else u = ib[i];
                    idx = (int32_t)u;
                } else idx = dr->first + (int32_t)i;
                float p[4], c[4], t[2], nn[3];
                fetch_attrib(d, &dr->position, idx, p, 4);
                if (dr->position.size == 2) p[2] = 0.0f;
                col4 cc = col_from(dr->cur_color);
                if (dr->color.enabled) { fetch_attrib(d, &dr->color, idx, c, 4); cc = sanitize_color(c); }
                t[0] = dr->cur_texcoord[0]; t[1] = dr->cur_texcoord[1];
                if (dr->texcoord.enabled) fetch_attrib(d, &dr->texcoord, idx, t, 2);
                nn[0] = dr->cur_normal[0]; nn[1] = dr->cur_normal[1]; nn[2] = dr->cur_normal[2];
                if (dr->normal.enabled) fetch_attrib(d, &dr->normal, idx, nn, 3);
                vertex_stage(&b->states[dr->vertex_state], p, cc, t, nn, &v[i]);
            }
        }
        d->stats.vertices += dr->count;
        assemble(d, &b->states[dr->raster_state], dr->mode, v, dr->count);
        free(v);
    }
    return MTGL_OK;
}

int mtgl_dev_finish(mtgl_dev *d) { return d ? MTGL_OK : MTGL_E_INVALID; }

int mtgl_dev_read_framebuffer(mtgl_dev *d, int32_t y0, int32_t y1, uint32_t *color, float *depth, uint8_t *stencil)
{
    if (!d || y0 < 0 || y1 > d->height || y0 > y1) return MTGL_E_INVALID;
    size_t o = (size_t)y0 * d->width, n = (size_t)(y1 - y0) * d->width;
    if (color) memcpy(color + o, d->color + o, n * 4);
    if (depth) memcpy(depth + o, d->depth + o, n * 4);
    if (stencil) memcpy(stencil + o, d->stencil + o, n);
    return MTGL_OK;
}

int mtgl_dev_write_framebuffer(mtgl_dev *d, int32_t y0, int32_t y1, const uint32_t *color, const float *depth,
                               const uint8_t *stencil)
{
    if (!d || y0 < 0 || y1 > d->height || y0 > y1) return MTGL_E_INVALID;
    size_t o = (size_t)y0 * d->width, n = (size_t)(y1 - y0) * d->width;
    if (color) memcpy(d->color + o, color + o, n * 4);
    if (depth) memcpy(d->depth + o, depth + o, n * 4);
    if (stencil) memcpy(d->stencil + o, stencil + o, n);
    return MTGL_OK;
}

int mtgl_dev_plane_pointers(mtgl_dev *d, void **color, void **depth, void **stencil)
{
    if (!d) return MTGL_E_INVALID;
    if (color) *color = d->color;
    if (depth) *depth = d->depth;
    if (stencil) *stencil = d->stencil;
    return MTGL_OK;
}

int mtgl_dev_get_stats(mtgl_dev *d, mtgl_dev_stats *out)
{
    if (!d || !out) return MTGL_E_INVALID;
    *out = d->stats;
    return MTGL_OK;
}

#include <time.h>
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* glDrawPixels, gl_api.c:1286-1373: per pixel -- row / column bounds, format expansion, alpha test on a / 255, depth
 * test of depth 0 against the stored depth, blend with the 8-bit destination, depth write, truncating pack */
int mtgl_dev_draw_pixels(mtgl_dev *d, const mtgl_pixel_rect *rc, const void *pixels)
{
    if (!d || !rc || !pixels) return MTGL_E_INVALID;
    const uint8_t *src = (const uint8_t *)pixels;
    const int alpha_on = (rc->caps & MTGL_CAP_ALPHA_TEST) != 0, depth_on = (rc->caps & MTGL_CAP_DEPTH_TEST) != 0;
    const int blend_on = (rc->caps & MTGL_CAP_BLEND) != 0;
    const uint32_t afunc = rc->alpha_func, dfunc = rc->depth_func;          /* GL tokens; unknown ones pass (1233-1264) */
    const float pixel_depth = 0.0f;                                         /* gl_api.c:1297 */
    for (int32_t row = 0; row < rc->height; row++) {
        const int32_t fy = d->height - 1 - (rc->y + row);                   /* 1304-1306 */
        if (fy < 0 || fy >= d->height || fy < d->band_y0 || fy >= d->band_y1) continue;
        for (int32_t col = 0; col < rc->width; col++) {
            const int32_t dx = rc->x + col;
            if (dx < 0 || dx >= d->width) continue;
            uint8_t r, g, b, a = 255;
            const size_t i = (size_t)row * (size_t)rc->width + (size_t)col;
            if (rc->format == 0x1908) { r = src[i * 4]; g = src[i * 4 + 1]; b = src[i * 4 + 2]; a = src[i * 4 + 3]; }
            else if (rc->format == 0x1907) { r = src[i * 3]; g = src[i * 3 + 1]; b = src[i * 3 + 2]; }
            else if (rc->format == 0x1909) { r = g = b = src[i]; }
            else if (rc->format == 0x190A) { r = g = b = src[i * 2]; a = src[i * 2 + 1]; }
            else continue;
            if (alpha_on && !cmp_f(afunc, (float)a / 255.0f, rc->alpha_ref)) continue;
            const size_t at = (size_t)fy * (size_t)d->width + (size_t)dx;
            if (depth_on && !cmp_f(dfunc, pixel_depth, d->depth[at])) continue;
            col4 s = { (float)r / 255.0f, (float)g / 255.0f, (float)b / 255.0f, (float)a / 255.0f };
            if (blend_on) {                                                 /* 1359-1365 */
                const col4 dc = col_unpack(d->color[at]);
                const col4 sf = blend_factor(rc->blend_src, s, dc), df = blend_factor(rc->blend_dst, s, dc);
                const col4 o = { s.r * sf.r + dc.r * df.r, s.g * sf.g + dc.g * df.g, s.b * sf.b + dc.b * df.b, s.a * sf.a + dc.a * df.a };
                s = col_clamp(o);
            }
            if (depth_on && rc->depth_mask) d->depth[at] = pixel_depth;
            d->color[at] = col_pack(s);
        }
    }
    return MTGL_OK;
}

/* glReadPixels, gl_api.c:1180-1230 (GL_RGBA / GL_RGB; anything else leaves the destination untouched) */
int mtgl_dev_read_pixels(mtgl_dev *d, int32_t x, int32_t y, int32_t width, int32_t height, uint32_t format, void *out)
{
    if (!d || !out) return MTGL_E_INVALID;
    const int bpp = (format == 0x1908) ? 4 : (format == 0x1907 ? 3 : 0);
    uint8_t *dst = (uint8_t *)out;
    if (!bpp) return MTGL_OK;
    for (int32_t row = 0; row < height; row++) {
        const int32_t fy = d->height - 1 - (y + row);
        if (fy < 0 || fy >= d->height) { memset(dst + (size_t)row * (size_t)width * (size_t)bpp, 0, (size_t)width * (size_t)bpp); continue; }
        for (int32_t col = 0; col < width; col++) {
            const int32_t sx = x + col;
            uint8_t px[4] = { 0, 0, 0, 255 };
            if (sx >= 0 && sx < d->width) {
                const uint32_t p = d->color[(size_t)fy * (size_t)d->width + (size_t)sx];
                px[0] = (uint8_t)(p & 0xFF); px[1] = (uint8_t)((p >> 8) & 0xFF); px[2] = (uint8_t)((p >> 16) & 0xFF); px[3] = (uint8_t)(p >> 24);
            }
            memcpy(dst + ((size_t)row * (size_t)width + (size_t)col) * (size_t)bpp, px, (size_t)bpp);
        }
    }
    return MTGL_OK;
}

void *mtgl_dev_stream(mtgl_dev *d) { (void)d; return NULL; }
int mtgl_dev_set_present_mode(mtgl_dev *d, int mode) { (void)mode; return d ? MTGL_OK : MTGL_E_INVALID; }
int mtgl_dev_frame_barrier(mtgl_dev *d, uint32_t participants) { (void)d; return participants == 1 ? MTGL_OK : MTGL_E_INVALID; }   /* no device, no stream */

int mtgl_dev_timer_mark(mtgl_dev *d, int which)
{
    if (!d || which < 0 || which > 1) return MTGL_E_INVALID;
    d->t_mark[which] = now_s();
    return MTGL_OK;
}

int mtgl_dev_timer_elapsed_ms(mtgl_dev *d, float *ms)
{
    if (!d || !ms) return MTGL_E_INVALID;
    *ms = (float)((d->t_mark[1] - d->t_mark[0]) * 1e3);
    return MTGL_OK;
}

/* Oracle-only: exact fragment counters -- fragments passing the inclusive inside
 * test, fragments past the stencil + depth tests, fragments reaching the colour write.  Cumulative. */
void mtgl_oracle_fragment_counts(mtgl_dev *d, uint64_t out[3])
{
    out[0] = d->frag_covered; out[1] = d->frag_tested; out[2] = d->frag_shaded;
}

const char *mtgl_dev_last_error(mtgl_dev *d) { return d ? d->err : "no device"; }
