/*
 * gl_front.cpp -- context management, draw batching and framebuffer access of the gl* front end.
 *
 * Behavioural mirror of the reference's draw front end (src/gl_api.c: gl_create_context 81-236,
 * emit_vertex 263-348, glClear 409-457, glBegin/glEnd 614-662, glVertex* 664-686,
 * glReadPixels/glDrawPixels/glRasterPos 1180-1424, glDrawArrays/glDrawElements 1745-1941) --
 * except that nothing is rasterised here: vertices, state snapshots and draw records are queued
 * and handed to the sm_100a back end through include/mtgl_dev.h at synchronisation points.
 *
 * Compile with IEEE semantics (-ffp-contract=off, no -ffast-math): the few floating-point
 * expressions evaluated on the host (matrix stack, normal matrix, light directions) must round
 * exactly like the reference's strict build.
 */
#include "front_internal.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

namespace mtgl {

thread_local GLState *g_ctx = nullptr;
static thread_local int g_device_ordinal = -1;

void set_error(GLState *c, GLenum e)
{
    if (c && c->error == GL_NO_ERROR) c->error = e;
}

/* ---------------------------------------------------------------- math (graphics.h) */
void mat_identity(float *m)
{
    for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
}

void mat_mul(const float *a, const float *b, float *out)
{
    float r[16];
    for (int col = 0; col < 4; col++)
        for (int row = 0; row < 4; row++)
            r[col * 4 + row] = a[row] * b[col * 4] + a[4 + row] * b[col * 4 + 1] + a[8 + row] * b[col * 4 + 2] +
                               a[12 + row] * b[col * 4 + 3];
    std::memcpy(out, r, sizeof r);
}

void mat_vec(const float *m, const float *v, float *out)
{
    float r[4];
    for (int row = 0; row < 4; row++)
        r[row] = m[row] * v[0] + m[4 + row] * v[1] + m[8 + row] * v[2] + m[12 + row] * v[3];
    std::memcpy(out, r, sizeof r);
}

float *current_matrix(GLState *c)
{
    switch (c->matrix_mode) {
    case GL_PROJECTION: return c->projection[c->projection_depth];
    case GL_TEXTURE: return c->texture[c->texture_depth];
    default: return c->modelview[c->modelview_depth];
    }
}

GLint *current_depth(GLState *c)
{
    switch (c->matrix_mode) {
    case GL_PROJECTION: return &c->projection_depth;
    case GL_TEXTURE: return &c->texture_depth;
    default: return &c->modelview_depth;
    }
}

static float sat(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

uint32_t pack_rgba(Rgba c)
{
    uint8_t r = (uint8_t)(sat(c.r) * 255.0f), g = (uint8_t)(sat(c.g) * 255.0f);
    uint8_t b = (uint8_t)(sat(c.b) * 255.0f), a = (uint8_t)(sat(c.a) * 255.0f);
    return ((uint32_t)a << 24) | ((uint32_t)b << 16) | ((uint32_t)g << 8) | r;
}

static void unit3(const float *v, float *o) /* vec3_normalize, graphics.h:54-60 */
{
    float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
    if (len > 0.0f) {
        float s = 1.0f / len;
        o[0] = v[0] * s; o[1] = v[1] * s; o[2] = v[2] * s;
    }
}

/* inverse-transpose of the upper 3x3 (mat4_normal_matrix, graphics.h:229-260); identity when singular */
static void normal_matrix(const float *m, float *out12)
{
    const float m00 = m[0], m10 = m[1], m20 = m[2];
    const float m01 = m[4], m11 = m[5], m21 = m[6];
    const float m02 = m[8], m12 = m[9], m22 = m[10];
    float det = m00 * (m11 * m22 - m21 * m12) - m10 * (m01 * m22 - m21 * m02) + m20 * (m01 * m12 - m11 * m02);
    std::memset(out12, 0, 12 * sizeof(float));
    if (fabsf(det) < 1e-10f) {
        out12[0] = 1.0f; out12[5] = 1.0f; out12[10] = 1.0f;
        return;
    }
    float k = 1.0f / det;
    out12[0] = (m11 * m22 - m21 * m12) * k;
    out12[1] = (m01 * m22 - m21 * m02) * -k;
    out12[2] = (m01 * m12 - m11 * m02) * k;
    out12[4] = (m10 * m22 - m20 * m12) * -k;
    out12[5] = (m00 * m22 - m20 * m02) * k;
    out12[6] = (m00 * m12 - m10 * m02) * -k;
    out12[8] = (m10 * m21 - m20 * m11) * k;
    out12[9] = (m00 * m21 - m20 * m01) * -k;
    out12[10] = (m00 * m11 - m10 * m01) * k;
}

/* ---------------------------------------------------------------- object lookup */
Texture *get_texture(GLState *c, GLuint id)
{
    if (id == 0 || id > c->textures.size()) return nullptr;
    Texture *t = &c->textures[id - 1];
    return t->allocated ? t : nullptr;
}

Buffer *get_buffer(GLState *c, GLuint id)
{
    if (id == 0 || id > c->buffers.size()) return nullptr;
    Buffer *b = &c->buffers[id - 1];
    return b->allocated ? b : nullptr;
}

const uint8_t *buffer_host_data(GLState *c, GLuint id)
{
    Buffer *b = get_buffer(c, id);
    if (!b || !b->has_data) return nullptr;
    if (!b->host_valid) {
        b->data.resize((size_t)b->size);
        if (mtgl_dev_buffer_read(c->dev, id, 0, b->size, b->data.data()) != MTGL_OK) return nullptr;
        b->host_valid = true;
    }
    return b->data.data();
}

bool buffer_read(GLState *c, GLuint id, uint64_t offset, uint64_t n, void *out)
{
    Buffer *b = get_buffer(c, id);
    if (!b || !b->has_data || offset + n > b->size) return false;
    if (b->host_valid) { std::memcpy(out, b->data.data() + offset, (size_t)n); return true; }
    if (n <= sizeof(Buffer::Peek::raw) && !b->peeks_frozen) {
        for (const Buffer::Peek &p : b->peeks)
            if (p.off == offset && p.n == n) { std::memcpy(out, p.raw, (size_t)n); return true; }
    }
    if (mtgl_dev_buffer_read(c->dev, id, offset, n, out) != MTGL_OK) return false;      /* waits for the queued frames */
    if (n <= sizeof(Buffer::Peek::raw) && !b->peeks_frozen) {
        if (b->peeks.size() >= 16) b->peeks.erase(b->peeks.begin());
        Buffer::Peek p; p.off = offset; p.n = (uint32_t)n;
        std::memcpy(p.raw, out, (size_t)n);
        b->peeks.push_back(p);
    }
    return true;
}

DisplayList *get_list(GLState *c, GLuint id)
{
    if (id == 0 || id > c->lists.size()) return nullptr;
    DisplayList *l = &c->lists[id - 1];
    return l->allocated ? l : nullptr;
}

/* ---------------------------------------------------------------- state snapshot */
static void copy4(float *d, const Rgba &s) { d[0] = s.r; d[1] = s.g; d[2] = s.b; d[3] = s.a; }

static void snapshot_material(mtgl_material *d, const Material &s)
{
    copy4(d->ambient, s.ambient); copy4(d->diffuse, s.diffuse);
    copy4(d->specular, s.specular); copy4(d->emission, s.emission);
    d->shininess = s.shininess;
}

static void snapshot_state(const GLState *c, mtgl_state *s)
{
    std::memset(s, 0, sizeof *s);
    std::memcpy(s->modelview, c->modelview[c->modelview_depth], 64);
    std::memcpy(s->projection, c->projection[c->projection_depth], 64);
    std::memcpy(s->texture, c->texture[c->texture_depth], 64);
    normal_matrix(s->modelview, s->normal);
    for (int i = 0; i < kMaxLights; i++) {
        const Light &l = c->lights[i];
        mtgl_light *d = &s->lights[i];
        copy4(d->ambient, l.ambient); copy4(d->diffuse, l.diffuse); copy4(d->specular, l.specular);
        std::memcpy(d->position, l.position, 16);
        unit3(l.position, d->dir_unit);
        unit3(l.spot_direction, d->spot_dir_unit);
        d->spot_exponent = l.spot_exponent;
        d->spot_cutoff = l.spot_cutoff;
        d->cos_cutoff = cosf(l.spot_cutoff * 3.14159265f / 180.0f);
        d->att_constant = l.att_constant; d->att_linear = l.att_linear; d->att_quadratic = l.att_quadratic;
        d->enabled = l.enabled ? 1u : 0u;
    }
    snapshot_material(&s->material_front, c->material_front);
    snapshot_material(&s->material_back, c->material_back);
    copy4(s->light_model_ambient, c->light_model_ambient);
    s->caps = c->caps;
    s->light_model_local_viewer = c->light_model_local_viewer ? 1u : 0u;
    s->light_model_two_side = c->light_model_two_side ? 1u : 0u;
    s->color_material_face = c->color_material_face;
    s->color_material_mode = c->color_material_mode;
    s->shade_model = c->shade_model;
    s->viewport[0] = c->viewport_x; s->viewport[1] = c->viewport_y;
    s->viewport[2] = c->viewport_w; s->viewport[3] = c->viewport_h;
    s->scissor[0] = c->scissor_x; s->scissor[1] = c->scissor_y;
    s->scissor[2] = c->scissor_w; s->scissor[3] = c->scissor_h;
    s->cull_face_mode = c->cull_face_mode; s->front_face = c->front_face;
    s->polygon_mode_front = c->polygon_mode_front; s->polygon_mode_back = c->polygon_mode_back;

    /* the rasteriser looks the texture up at glEnd (raster.c:495-498) */
    const Texture *t = nullptr;
    if ((c->caps & MTGL_CAP_TEXTURE_2D) && c->bound_texture_2d != 0)
        t = get_texture(const_cast<GLState *>(c), c->bound_texture_2d);
    if (t && !t->pixels.empty()) {
        s->texture_id = c->bound_texture_2d;
        s->tex_min_filter = t->min_filter; s->tex_mag_filter = t->mag_filter;
        s->tex_wrap_s = t->wrap_s; s->tex_wrap_t = t->wrap_t;
    }
    s->tex_env_mode = c->tex_env_mode;
    copy4(s->tex_env_color, c->tex_env_color);
    s->perspective_hint = c->perspective_hint;
    s->alpha_func = c->alpha_func; s->alpha_ref = c->alpha_ref;
    s->stencil_func = c->stencil_func; s->stencil_ref = c->stencil_ref; s->stencil_mask = c->stencil_mask;
    s->stencil_fail = c->stencil_fail; s->stencil_zfail = c->stencil_zfail; s->stencil_zpass = c->stencil_zpass;
    s->stencil_writemask = c->stencil_writemask;
    s->depth_func = c->depth_func; s->depth_mask = c->depth_mask ? 1u : 0u;
    s->blend_src = c->blend_src; s->blend_dst = c->blend_dst;
    s->color_mask = (c->color_mask[0] ? 1u : 0u) | (c->color_mask[1] ? 2u : 0u) | (c->color_mask[2] ? 4u : 0u) |
                    (c->color_mask[3] ? 8u : 0u);
    s->fog_mode = c->fog_mode; s->fog_density = c->fog_density; s->fog_start = c->fog_start; s->fog_end = c->fog_end;
    copy4(s->fog_color, c->fog_color);
    s->line_width = c->line_width; s->point_size = c->point_size;
    s->depth_near = c->depth_near; s->depth_far = c->depth_far;
}

void mark_state_dirty(GLState *c) { c->vstate_dirty = true; c->vstate_full_dirty = true; }
void mark_matrix_dirty(GLState *c) { c->vstate_dirty = true; }

/* index of a state block equal to the live state, appending one when the last block differs */
static uint32_t current_state_block(GLState *c)
{
    if (!c->vstate_dirty && !c->states.empty()) return c->vstate_index;
    /* A render loop changes matrices far more often than anything else (glPushMatrix / glTranslatef / draw /
     * glPopMatrix per object): then the previous snapshot is patched -- three matrices and the normal matrix -- instead
     * of being rebuilt (eight lights with two normalisations and a cosf each, materials, ~60 scalar fields). */
    mtgl_state &s = c->last_snapshot;
    if (c->vstate_full_dirty || !c->have_snapshot) {
        snapshot_state(c, &s);
        c->have_snapshot = true;
        c->vstate_full_dirty = false;
    } else {
        std::memcpy(s.modelview, c->modelview[c->modelview_depth], 64);
        std::memcpy(s.projection, c->projection[c->projection_depth], 64);
        std::memcpy(s.texture, c->texture[c->texture_depth], 64);
        normal_matrix(s.modelview, s.normal);
    }
    if (c->states.empty() || std::memcmp(&c->states.back(), &s, sizeof s) != 0) c->states.push_back(s);
    c->vstate_index = (uint32_t)c->states.size() - 1;
    c->vstate_dirty = false;
    return c->vstate_index;
}

/* ---------------------------------------------------------------- batching */
void flush_batch(GLState *c)
{
    if (c->draws.empty() && c->pending_clear_mask == 0) return;
    mtgl_batch b;
    std::memset(&b, 0, sizeof b);
    b.states = c->states.data(); b.n_states = (uint32_t)c->states.size();
    b.vertices = c->staged.data(); b.n_vertices = c->prim_first;
    b.draws = c->draws.data(); b.n_draws = (uint32_t)c->draws.size();
    b.blob = c->blob.data(); b.blob_size = c->blob.size();
    b.clear_mask = c->pending_clear_mask;
    std::memcpy(b.clear_rect, c->pending_clear_rect, sizeof b.clear_rect);
    b.clear_color = c->pending_clear_color;
    b.clear_depth = c->pending_clear_depth;
    b.clear_stencil = c->pending_clear_stencil;
    int rc = mtgl_dev_submit(c->dev, &b);
    if (rc != MTGL_OK) set_error(c, rc == MTGL_E_OOM ? GL_OUT_OF_MEMORY : GL_INVALID_OPERATION);
    c->draws.clear();
    c->blob.clear();
    c->pending_clear_mask = 0;
    /* vertices of a primitive still being assembled stay queued */
    c->staged.erase(c->staged.begin(), c->staged.begin() + c->prim_first);
    c->prim_first = 0;
    if (c->staged.empty()) {
        c->states.clear();
        c->vstate_dirty = true;
    }
}

void sync_device(GLState *c)
{
    flush_batch(c);
    if (mtgl_dev_finish(c->dev) != MTGL_OK) set_error(c, GL_INVALID_OPERATION);
}

constexpr size_t kMaxStagedVertices = 1u << 22;
constexpr size_t kMaxBatchDraws = 1u << 16;

static void apply_color_material(GLState *c, const Rgba &color) /* gl_api.c:285-312 */
{
    if (!(c->caps & MTGL_CAP_LIGHTING) || !(c->caps & MTGL_CAP_COLOR_MATERIAL)) return;
    Rgba k = rgba(color.r < 0 ? 0 : (color.r > 1 ? 1 : color.r), color.g < 0 ? 0 : (color.g > 1 ? 1 : color.g),
                  color.b < 0 ? 0 : (color.b > 1 ? 1 : color.b), color.a < 0 ? 0 : (color.a > 1 ? 1 : color.a));
    GLenum mode = c->color_material_mode, face = c->color_material_face;
    Material *m[2] = { (face == GL_FRONT || face == GL_FRONT_AND_BACK) ? &c->material_front : nullptr,
                       (face == GL_BACK || face == GL_FRONT_AND_BACK) ? &c->material_back : nullptr };
    for (Material *mat : m) {
        if (!mat) continue;
        if (mode == GL_AMBIENT || mode == GL_AMBIENT_AND_DIFFUSE) mat->ambient = k;
        if (mode == GL_DIFFUSE || mode == GL_AMBIENT_AND_DIFFUSE) mat->diffuse = k;
        if (mode == GL_SPECULAR) mat->specular = k;
        if (mode == GL_EMISSION) mat->emission = k;
    }
    /* The device re-applies this override per vertex from the vertex colour, so the queued
     * vertex-state block stays valid; the raster-state snapshot at glEnd must see the new
     * material (raster.c:599-613 reads ctx->material_* live). */
    c->material_touched = true;
}

void emit_vertex(GLState *c, float x, float y, float z)
{
    uint32_t st = current_state_block(c);
    mtgl_in_vertex v;
    v.position[0] = x; v.position[1] = y; v.position[2] = z;
    v.color[0] = c->current_color.r; v.color[1] = c->current_color.g;
    v.color[2] = c->current_color.b; v.color[3] = c->current_color.a;
    v.texcoord[0] = c->current_texcoord[0]; v.texcoord[1] = c->current_texcoord[1];
    v.normal[0] = c->current_normal[0]; v.normal[1] = c->current_normal[1]; v.normal[2] = c->current_normal[2];
    v.state = st;
    c->staged.push_back(v);
    apply_color_material(c, c->current_color);
}

static uint32_t raster_state_block(GLState *c)
{
    if (c->material_touched) {
        mark_state_dirty(c);
        c->material_touched = false;
    }
    return current_state_block(c);
}

void end_primitive(GLState *c) /* the flush_* dispatch of glEnd, gl_api.c:648-661 */
{
    uint32_t n = (uint32_t)c->staged.size() - c->prim_first;
    if (n > 0) {
        mtgl_draw d;
        std::memset(&d, 0, sizeof d);
        d.mode = c->primitive_mode;
        d.count = n;
        d.raster_state = raster_state_block(c);
        d.source = MTGL_SRC_STAGED;
        d.first_staged = c->prim_first;
        c->draws.push_back(d);
    }
    c->prim_first = (uint32_t)c->staged.size();
    if (c->staged.size() > kMaxStagedVertices || c->draws.size() > kMaxBatchDraws) flush_batch(c);
}

} // namespace mtgl

using namespace mtgl;

/* ================================================================ context API */
extern "C" {

void mtgl_set_device(int ordinal) { g_device_ordinal = ordinal; }

GLState *gl_create_context(int32_t width, int32_t height)
{
    GLState *c = new (std::nothrow) GLState();
    if (!c) return nullptr;
    c->dev = nullptr;
    if (mtgl_dev_create(width, height, g_device_ordinal, &c->dev) != MTGL_OK) {
        delete c;
        return nullptr;
    }
    c->fb_width = width; c->fb_height = height;
    c->mirror.width = width; c->mirror.height = height;
    c->mirror.color = nullptr; c->mirror.depth = nullptr; c->mirror.stencil = nullptr;

    /* defaults: gl_api.c:92-233 */
    c->clear_color = rgba(0, 0, 0, 1);
    c->clear_depth = 1.0;
    c->stencil_clear = 0;
    c->viewport_x = 0; c->viewport_y = 0; c->viewport_w = width; c->viewport_h = height;
    c->current_color = rgba(1, 1, 1, 1);
    c->current_texcoord[0] = 0; c->current_texcoord[1] = 0;
    c->current_normal[0] = 0; c->current_normal[1] = 0; c->current_normal[2] = 1;
    c->matrix_mode = GL_MODELVIEW;
    c->modelview_depth = c->projection_depth = c->texture_depth = 0;
    for (int i = 0; i < kMatrixStackDepth; i++) {
        mat_identity(c->modelview[i]); mat_identity(c->projection[i]); mat_identity(c->texture[i]);
    }
    c->primitive_mode = 0;
    c->inside_begin_end = false;
    c->caps = 0;
    c->blend_src = GL_ONE; c->blend_dst = GL_ZERO;
    c->cull_face_mode = GL_BACK; c->front_face = GL_CCW;
    c->depth_func = GL_LESS; c->depth_mask = GL_TRUE;
    c->alpha_func = GL_ALWAYS; c->alpha_ref = 0.0f;
    c->scissor_x = 0; c->scissor_y = 0; c->scissor_w = width; c->scissor_h = height;
    c->stencil_func = GL_ALWAYS; c->stencil_ref = 0; c->stencil_mask = 0xFFFFFFFFu;
    c->stencil_fail = c->stencil_zfail = c->stencil_zpass = GL_KEEP;
    c->stencil_writemask = 0xFFFFFFFFu;
    c->depth_near = 0.0; c->depth_far = 1.0;
    c->color_mask[0] = c->color_mask[1] = c->color_mask[2] = c->color_mask[3] = GL_TRUE;
    c->line_width = 1.0f; c->point_size = 1.0f;
    c->polygon_mode_front = c->polygon_mode_back = GL_FILL;
    c->bound_texture_2d = 0;
    c->tex_env_mode = GL_MODULATE;
    c->tex_env_color = rgba(0, 0, 0, 0);
    c->perspective_hint = GL_DONT_CARE;
    c->raster_pos_x = 0; c->raster_pos_y = 0; c->raster_pos_valid = GL_TRUE;
    c->fog_mode = GL_EXP; c->fog_density = 1.0f; c->fog_start = 0.0f; c->fog_end = 1.0f;
    c->fog_color = rgba(0, 0, 0, 0);
    for (int i = 0; i < kMaxLights; i++) { /* light_init, lighting.h:18-32 */
        Light &l = c->lights[i];
        l.ambient = rgba(0, 0, 0, 1);
        l.diffuse = (i == 0) ? rgba(1, 1, 1, 1) : rgba(0, 0, 0, 1);
        l.specular = l.diffuse;
        l.position[0] = 0; l.position[1] = 0; l.position[2] = 1; l.position[3] = 0;
        l.spot_direction[0] = 0; l.spot_direction[1] = 0; l.spot_direction[2] = -1;
        l.spot_exponent = 0; l.spot_cutoff = 180;
        l.att_constant = 1; l.att_linear = 0; l.att_quadratic = 0;
        l.enabled = GL_FALSE;
    }
    c->material_front.ambient = rgba(0.2f, 0.2f, 0.2f, 1); /* material_init, lighting.h:35-42 */
    c->material_front.diffuse = rgba(0.8f, 0.8f, 0.8f, 1);
    c->material_front.specular = rgba(0, 0, 0, 1);
    c->material_front.emission = rgba(0, 0, 0, 1);
    c->material_front.shininess = 0;
    c->material_back = c->material_front;
    c->light_model_ambient = rgba(0.2f, 0.2f, 0.2f, 1.0f);
    c->light_model_local_viewer = GL_FALSE; c->light_model_two_side = GL_FALSE;
    c->color_material_face = GL_FRONT_AND_BACK; c->color_material_mode = GL_AMBIENT_AND_DIFFUSE;
    c->shade_model = GL_SMOOTH;
    c->bound_array_buffer = 0; c->bound_element_buffer = 0;
    c->client_state = 0;
    std::memset(&c->vertex_pointer, 0, sizeof(ArrayPointer));
    std::memset(&c->color_pointer, 0, sizeof(ArrayPointer));
    std::memset(&c->texcoord_pointer, 0, sizeof(ArrayPointer));
    std::memset(&c->normal_pointer, 0, sizeof(ArrayPointer));
    c->list_base = 0; c->list_index = 0; c->list_mode = 0; c->list_call_depth = 0;
    c->error = GL_NO_ERROR;

    c->prim_first = 0;
    c->vstate_dirty = true;
    c->vstate_full_dirty = true;
    c->have_snapshot = false;
    c->vstate_index = 0;
    c->material_touched = false;
    c->pending_clear_mask = 0;
    return c;
}

void gl_destroy_context(GLState *c)
{
    if (!c) return;
    if (g_ctx == c) g_ctx = nullptr;
    if (c->dev) {
        mtgl_dev_finish(c->dev);
        mtgl_dev_destroy(c->dev);
    }
    delete c;
}

void gl_make_current(GLState *c) { g_ctx = c; }
GLState *gl_get_current_context(void) { return g_ctx; }
struct mtgl_dev *mtgl_context_device(GLState *c) { return c ? c->dev : nullptr; }
uint64_t mtgl_context_list_runs_drawn(GLState *c) { return c ? c->list_runs_drawn : 0; }

/* Device address of a buffer object's storage for applications that fill it behind the API (a sharded upload completed
 * by an NCCL all-gather): from here on the front end cannot know the contents, so what it remembers of them is dropped
 * and the bytes it needs (the last element of an array draw, gl_api.c:1826-1842) are read from the device when asked for. */
int mtgl_context_buffer_pointer(GLState *c, unsigned id, void **ptr, uint64_t *size)
{
    if (!c) return MTGL_E_INVALID;
    Buffer *b = get_buffer(c, id);
    if (!b) return MTGL_E_INVALID;
    flush_batch(c);
    b->peeks.clear();
    b->peeks_frozen = true;
    b->host_valid = false;
    b->data.clear();
    return mtgl_dev_buffer_pointer(c->dev, id, ptr, size);
}

int mtgl_context_buffer_orphan(GLState *c, unsigned id, const void *contents, void **ptr, uint64_t *size)
{
    if (!c || !ptr) return MTGL_E_INVALID;
    Buffer *b = get_buffer(c, id);
    if (!b || !b->has_data || b->size == 0) return MTGL_E_INVALID;
    flush_batch(c);                             /* queued draws read the storage the name has now */
    if (contents) {                             /* what the host has looked at before, from the contents to come (as glBufferData) */
        for (Buffer::Peek &p : b->peeks) std::memcpy(p.raw, (const uint8_t *)contents + p.off, p.n);
        b->peeks_frozen = false;
        if (b->size <= kHostMirrorLimit) {
            b->data.assign((const uint8_t *)contents, (const uint8_t *)contents + b->size);
            b->host_valid = true;
        }
    } else {
        b->peeks.clear();
        b->peeks_frozen = true;
        b->host_valid = false;
        b->data.clear();
    }
    if (size) *size = b->size;
    return mtgl_dev_buffer_orphan(c->dev, id, b->size, ptr);
}

const mtgl_framebuffer *mtgl_map_framebuffer(GLState *c, unsigned planes)
{
    if (!c) return nullptr;
    sync_device(c);
    size_t n = (size_t)c->fb_width * c->fb_height;
    uint32_t *col = nullptr; float *dep = nullptr; uint8_t *sten = nullptr;
    if (planes & MTGL_PLANE_COLOR) { c->mirror_color.resize(n); col = c->mirror_color.data(); }
    if (planes & MTGL_PLANE_DEPTH) { c->mirror_depth.resize(n); dep = c->mirror_depth.data(); }
    if (planes & MTGL_PLANE_STENCIL) { c->mirror_stencil.resize(n); sten = c->mirror_stencil.data(); }
    if (mtgl_dev_read_framebuffer(c->dev, 0, c->fb_height, col, dep, sten) != MTGL_OK) {
        set_error(c, GL_INVALID_OPERATION);
        return nullptr;
    }
    c->mirror.color = c->mirror_color.empty() ? nullptr : c->mirror_color.data();
    c->mirror.depth = c->mirror_depth.empty() ? nullptr : c->mirror_depth.data();
    c->mirror.stencil = c->mirror_stencil.empty() ? nullptr : c->mirror_stencil.data();
    return &c->mirror;
}

/* include/mtgl_context.h: queue the frame, then a copy of its colour rows into page-locked host memory; glFinish() waits */
void mtglReadColorAsync(int32_t y0, int32_t y1, uint32_t *pinned_color)
{
    MTGL_CTX();
    if (!pinned_color) { set_error(c, GL_INVALID_VALUE); return; }
    flush_batch(c);
    if (mtgl_dev_read_color_async(c->dev, y0, y1, pinned_color) != MTGL_OK) set_error(c, GL_INVALID_OPERATION);
}

/* ================================================================ clears and synchronisation */
void glClear(GLbitfield mask) /* gl_api.c:409-457 */
{
    MTGL_CTX();
    int32_t x0 = 0, y0 = 0, x1 = c->fb_width, y1 = c->fb_height;
    if (c->caps & MTGL_CAP_SCISSOR_TEST) {
        x0 = c->scissor_x; y0 = c->scissor_y;
        x1 = c->scissor_x + c->scissor_w; y1 = c->scissor_y + c->scissor_h;
        if (x0 < 0) x0 = 0;
        if (y0 < 0) y0 = 0;
        if (x1 > c->fb_width) x1 = c->fb_width;
        if (y1 > c->fb_height) y1 = c->fb_height;
    }
    mask &= (GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT | GL_STENCIL_BUFFER_BIT);
    if (mask == 0 || x0 >= x1 || y0 >= y1) return;
    int32_t rect[4] = { x0, y0, x1, y1 };
    /* a clear can only ride in front of a batch: anything already queued must run first */
    if (!c->draws.empty() ||
        (c->pending_clear_mask != 0 && std::memcmp(rect, c->pending_clear_rect, sizeof rect) != 0))
        flush_batch(c);
    std::memcpy(c->pending_clear_rect, rect, sizeof rect);
    c->pending_clear_mask |= mask;
    if (mask & GL_COLOR_BUFFER_BIT) c->pending_clear_color = pack_rgba(c->clear_color);
    if (mask & GL_DEPTH_BUFFER_BIT) c->pending_clear_depth = (float)c->clear_depth;
    if (mask & GL_STENCIL_BUFFER_BIT) c->pending_clear_stencil = (uint32_t)(c->stencil_clear & 0xFF);
}

void glFlush(void)
{
    MTGL_CTX();
    flush_batch(c);
}

void glFinish(void)
{
    MTGL_CTX();
    sync_device(c);
}

/* ================================================================ immediate mode */
void glBegin(GLenum mode) /* gl_api.c:614-633 */
{
    MTGL_CTX();
    ListCmd cmd; cmd.op = OP_BEGIN; cmd.e[0] = mode;
    if (record(c, cmd)) return;
    if (c->inside_begin_end) { set_error(c, GL_INVALID_OPERATION); return; }
    if (mode > GL_POLYGON) { set_error(c, GL_INVALID_ENUM); return; }
    c->primitive_mode = mode;
    c->inside_begin_end = true;
}

void glEnd(void) /* gl_api.c:635-662 */
{
    MTGL_CTX();
    ListCmd cmd; cmd.op = OP_END;
    if (record(c, cmd)) return;
    if (!c->inside_begin_end) { set_error(c, GL_INVALID_OPERATION); return; }
    c->inside_begin_end = false;
    end_primitive(c);
}

void glVertex3f(GLfloat x, GLfloat y, GLfloat z)
{
    MTGL_CTX();
    ListCmd cmd; cmd.op = OP_VERTEX; cmd.f[0] = x; cmd.f[1] = y; cmd.f[2] = z;
    if (record(c, cmd)) return;
    emit_vertex(c, x, y, z);
}

void glVertex2f(GLfloat x, GLfloat y)
{
    MTGL_CTX();
    ListCmd cmd; cmd.op = OP_VERTEX; cmd.f[0] = x; cmd.f[1] = y; cmd.f[2] = 0.0f;
    if (record(c, cmd)) return;
    emit_vertex(c, x, y, 0.0f);
}

void glVertex2i(GLint x, GLint y) { glVertex2f((float)x, (float)y); }
void glVertex3i(GLint x, GLint y, GLint z) { glVertex3f((float)x, (float)y, (float)z); }

static bool finite_f(float f) { return !std::isnan(f) && !std::isinf(f); }

void glColor3f(GLfloat r, GLfloat g, GLfloat b) /* gl_api.c:694-706 */
{
    MTGL_CTX();
    if (!finite_f(r)) r = 0.0f;
    if (!finite_f(g)) g = 0.0f;
    if (!finite_f(b)) b = 0.0f;
    r = sat(r); g = sat(g); b = sat(b);
    ListCmd cmd; cmd.op = OP_COLOR; cmd.f[0] = r; cmd.f[1] = g; cmd.f[2] = b; cmd.f[3] = 1.0f;
    if (record(c, cmd)) return;
    c->current_color = rgba(r, g, b, 1.0f);
}

void glColor4f(GLfloat r, GLfloat g, GLfloat b, GLfloat a) /* gl_api.c:708-722 */
{
    MTGL_CTX();
    if (!finite_f(r)) r = 0.0f;
    if (!finite_f(g)) g = 0.0f;
    if (!finite_f(b)) b = 0.0f;
    if (!finite_f(a)) a = 1.0f;
    r = sat(r); g = sat(g); b = sat(b); a = sat(a);
    ListCmd cmd; cmd.op = OP_COLOR; cmd.f[0] = r; cmd.f[1] = g; cmd.f[2] = b; cmd.f[3] = a;
    if (record(c, cmd)) return;
    c->current_color = rgba(r, g, b, a);
}

void glColor3ub(GLubyte r, GLubyte g, GLubyte b)
{
    MTGL_CTX();
    float rf = r / 255.0f, gf = g / 255.0f, bf = b / 255.0f;
    ListCmd cmd; cmd.op = OP_COLOR; cmd.f[0] = rf; cmd.f[1] = gf; cmd.f[2] = bf; cmd.f[3] = 1.0f;
    if (record(c, cmd)) return;
    c->current_color = rgba(rf, gf, bf, 1.0f);
}

void glColor4ub(GLubyte r, GLubyte g, GLubyte b, GLubyte a)
{
    MTGL_CTX();
    float rf = r / 255.0f, gf = g / 255.0f, bf = b / 255.0f, af = a / 255.0f;
    ListCmd cmd; cmd.op = OP_COLOR; cmd.f[0] = rf; cmd.f[1] = gf; cmd.f[2] = bf; cmd.f[3] = af;
    if (record(c, cmd)) return;
    c->current_color = rgba(rf, gf, bf, af);
}

void glTexCoord2f(GLfloat s, GLfloat t)
{
    MTGL_CTX();
    ListCmd cmd; cmd.op = OP_TEXCOORD; cmd.f[0] = s; cmd.f[1] = t;
    if (record(c, cmd)) return;
    c->current_texcoord[0] = s; c->current_texcoord[1] = t;
}

void glNormal3f(GLfloat nx, GLfloat ny, GLfloat nz)
{
    MTGL_CTX();
    ListCmd cmd; cmd.op = OP_NORMAL; cmd.f[0] = nx; cmd.f[1] = ny; cmd.f[2] = nz;
    if (record(c, cmd)) return;
    c->current_normal[0] = nx; c->current_normal[1] = ny; c->current_normal[2] = nz;
}

/* ================================================================ vertex arrays */
static const uint8_t *array_base(GLState *c, const ArrayPointer &a) /* get_array_pointer, gl_api.c:1745-1755 */
{
    if (c->bound_array_buffer) {
        const uint8_t *base = buffer_host_data(c, c->bound_array_buffer);
        return base ? base + (size_t)a.pointer : nullptr;
    }
    return (const uint8_t *)a.pointer;
}

static GLsizei resolved_stride(const ArrayPointer &a)
{
    if (a.stride != 0) return a.stride;
    return a.size * (a.type == GL_FLOAT ? 4 : 1);
}

static void array_element(const ArrayPointer &a, const uint8_t *base, GLint index, float *out, int want)
{ /* get_array_element, gl_api.c:1758-1797 */
    GLsizei stride = resolved_stride(a);
    if (!base || index < 0 || stride <= 0) {
        for (int i = 0; i < want; i++) out[i] = (i < 3) ? 0.0f : 1.0f;
        return;
    }
    const uint8_t *p = base + (size_t)index * (size_t)stride;
    for (int i = 0; i < a.size && i < want; i++) {
        if (a.type == GL_FLOAT) { float f; std::memcpy(&f, p + 4 * i, 4); out[i] = f; }
        else if (a.type == GL_UNSIGNED_BYTE) out[i] = p[i] / 255.0f;
    }
    for (int i = a.size; i < want; i++) out[i] = (i == 3) ? 1.0f : 0.0f;
}

static GLuint fetch_index(GLenum type, const void *data, GLsizei i)
{
    if (type == GL_UNSIGNED_SHORT) return ((const GLushort *)data)[i];
    if (type == GL_UNSIGNED_INT) return ((const GLuint *)data)[i];
    return ((const GLubyte *)data)[i];
}

/* The reference's own loop: every element goes through the public immediate-mode calls. */
static void draw_expanded(GLState *c, GLenum mode, GLsizei count, GLint first, GLenum index_type, const void *indices)
{
    const uint8_t *vb = array_base(c, c->vertex_pointer);
    const uint8_t *cb = (c->client_state & 2u) ? array_base(c, c->color_pointer) : nullptr;
    const uint8_t *tb = (c->client_state & 4u) ? array_base(c, c->texcoord_pointer) : nullptr;
    const uint8_t *nb = (c->client_state & 8u) ? array_base(c, c->normal_pointer) : nullptr;
    glBegin(mode);
    for (GLsizei i = 0; i < count; i++) {
        GLint idx = index_type ? (GLint)fetch_index(index_type, indices, i) : first + i;
        float v[4], col[4], t[2], n[3];
        array_element(c->vertex_pointer, vb, idx, v, 4);
        if (cb) { array_element(c->color_pointer, cb, idx, col, 4); glColor4f(col[0], col[1], col[2], col[3]); }
        if (tb) { array_element(c->texcoord_pointer, tb, idx, t, 2); glTexCoord2f(t[0], t[1]); }
        if (nb) { array_element(c->normal_pointer, nb, idx, n, 3); glNormal3f(n[0], n[1], n[2]); }
        if (c->vertex_pointer.size == 2) glVertex2f(v[0], v[1]);
        else glVertex3f(v[0], v[1], v[2]);
    }
    glEnd();
}

static void describe_attrib(GLState *c, const ArrayPointer &a, bool enabled, mtgl_attrib *d)
{
    std::memset(d, 0, sizeof *d);
    if (!enabled) return;
    d->enabled = 1;
    d->buffer = c->bound_array_buffer;
    d->offset = (uint64_t)(size_t)a.pointer;
    d->stride = (uint32_t)resolved_stride(a);
    d->size = (uint16_t)a.size;
    d->type = (a.type == GL_FLOAT) ? MTGL_TYPE_F32 : MTGL_TYPE_U8;
}

/* host-side read of one element with a bounds check (the device does the same check) */
static bool host_element(GLState *c, GLuint buffer, const ArrayPointer &a, GLint idx, float *out, int want)
{
    GLsizei stride = resolved_stride(a);
    if (idx < 0 || stride <= 0) {           /* get_array_element (gl_api.c:1761-1782): the defaults, which then become current */
        for (int i = 0; i < want; i++) out[i] = (i < 3) ? 0.0f : 1.0f;
        return true;
    }
    size_t comp = (a.type == GL_FLOAT) ? 4 : 1;
    uint64_t off = (uint64_t)(size_t)a.pointer + (uint64_t)idx * (uint64_t)stride;
    uint8_t raw[16];
    if (!buffer_read(c, buffer, off, comp * (size_t)a.size, raw)) return false;
    ArrayPointer one = a;
    array_element(one, raw, 0, out, want);
    return true;
}

/* Device-side attribute fetch: the draw is queued as one record and the vertex stage reads the
 * buffer-object mirror in HBM.  Falls back to the reference's element-by-element loop whenever
 * the call has observable host-side structure (list compilation, nesting, client memory). */
static void draw_arrays_common(GLState *c, GLenum mode, GLsizei count, GLint first, GLenum index_type,
                               const void *indices, bool indices_in_buffer, uint64_t index_offset)
{
    Buffer *vbuf = c->bound_array_buffer ? get_buffer(c, c->bound_array_buffer) : nullptr;
    bool fast = vbuf && vbuf->has_data && !compiling(c) && !c->inside_begin_end && mode <= GL_POLYGON &&
                c->staged.size() == c->prim_first && resolved_stride(c->vertex_pointer) > 0;
    if (!fast) {
        if (indices_in_buffer) {
            const uint8_t *ib = buffer_host_data(c, c->bound_element_buffer);
            indices = ib ? ib + index_offset : nullptr;
        }
        if (index_type && !indices) return;
        draw_expanded(c, mode, count, first, index_type, indices);
        return;
    }
    const GLuint vbo = c->bound_array_buffer;
    c->primitive_mode = mode;
    if (count == 0) return;

    mtgl_draw d;
    std::memset(&d, 0, sizeof d);
    d.mode = mode;
    d.count = (uint32_t)count;
    d.source = MTGL_SRC_ARRAYS;
    d.first = first;
    d.index_type = index_type;
    if (index_type) {
        if (indices_in_buffer) {
            d.index_buffer = c->bound_element_buffer;
            d.index_offset = index_offset;
        } else {
            size_t isz = (index_type == GL_UNSIGNED_INT) ? 4 : (index_type == GL_UNSIGNED_SHORT ? 2 : 1);
            size_t at = (c->blob.size() + 3) & ~(size_t)3;
            c->blob.resize(at + isz * (size_t)count);
            std::memcpy(c->blob.data() + at, indices, isz * (size_t)count);
            d.index_offset = at;
        }
    }
    describe_attrib(c, c->vertex_pointer, true, &d.position);
    describe_attrib(c, c->color_pointer, (c->client_state & 2u) != 0, &d.color);
    describe_attrib(c, c->texcoord_pointer, (c->client_state & 4u) != 0, &d.texcoord);
    describe_attrib(c, c->normal_pointer, (c->client_state & 8u) != 0, &d.normal);
    d.cur_color[0] = c->current_color.r; d.cur_color[1] = c->current_color.g;
    d.cur_color[2] = c->current_color.b; d.cur_color[3] = c->current_color.a;
    d.cur_texcoord[0] = c->current_texcoord[0]; d.cur_texcoord[1] = c->current_texcoord[1];
    d.cur_normal[0] = c->current_normal[0]; d.cur_normal[1] = c->current_normal[1]; d.cur_normal[2] = c->current_normal[2];
    d.vertex_state = current_state_block(c);

    /* host-visible side effects of the per-element glColor4f/glTexCoord2f/glNormal3f calls:
     * the "current" attributes end up holding the last element's values (gl_api.c:1826-1842) */
    GLint last = first + (count - 1);
    if (index_type) {
        size_t isz = (index_type == GL_UNSIGNED_INT) ? 4 : (index_type == GL_UNSIGNED_SHORT ? 2 : 1);
        uint32_t raw = 0;
        if (indices_in_buffer) buffer_read(c, c->bound_element_buffer, index_offset + isz * (size_t)(count - 1), isz, &raw);
        else std::memcpy(&raw, (const uint8_t *)indices + isz * (size_t)(count - 1), isz);
        last = (GLint)raw;
    }
    float tmp[4];
    if (d.color.enabled && host_element(c, vbo, c->color_pointer, last, tmp, 4)) {
        float r = tmp[0], g = tmp[1], b = tmp[2], a = tmp[3];
        if (!finite_f(r)) r = 0.0f;
        if (!finite_f(g)) g = 0.0f;
        if (!finite_f(b)) b = 0.0f;
        if (!finite_f(a)) a = 1.0f;
        c->current_color = rgba(sat(r), sat(g), sat(b), sat(a));
    }
    if (d.texcoord.enabled && host_element(c, vbo, c->texcoord_pointer, last, tmp, 2)) {
        c->current_texcoord[0] = tmp[0]; c->current_texcoord[1] = tmp[1];
    }
    if (d.normal.enabled && host_element(c, vbo, c->normal_pointer, last, tmp, 3)) {
        c->current_normal[0] = tmp[0]; c->current_normal[1] = tmp[1]; c->current_normal[2] = tmp[2];
    }
    apply_color_material(c, c->current_color);

    d.raster_state = raster_state_block(c);
    c->draws.push_back(d);
    if (c->draws.size() > kMaxBatchDraws) flush_batch(c);
}

} // extern "C"

namespace mtgl {

/* One compiled run of a display list as an array draw from the list's device buffer.  The
 * reference replays the list call by call (execute_list, gl_api.c:2331-2439): glBegin, then per vertex glColor4f /
 * glTexCoord2f / glNormal3f + emit_vertex, then glEnd -- exactly the sequence glDrawArrays runs per element
 * (1826-1849), so the array path produces the same vertices.  Attributes the run never sets come from the current
 * values at call time (disabled arrays); the ones it sets stay current afterwards. */
bool draw_list_run(GLState *c, GLuint id, const ListRun &r)
{
    if (compiling(c) || c->inside_begin_end || c->staged.size() != c->prim_first || r.count == 0) return false;
    /* the colour of every vertex also drives the material while GL_COLOR_MATERIAL is on (gl_api.c:285-312): that walk
     * is left to the per-call path */
    if (r.has_color && (c->caps & MTGL_CAP_LIGHTING) && (c->caps & MTGL_CAP_COLOR_MATERIAL)) return false;
    c->primitive_mode = r.mode;
    mtgl_draw d;
    std::memset(&d, 0, sizeof d);
    d.mode = r.mode;
    d.count = r.count;
    d.source = MTGL_SRC_ARRAYS;
    d.first = (int32_t)r.first;
    auto attrib = [&](mtgl_attrib &a, bool on, uint32_t offset, uint16_t size) {
        a.enabled = on ? 1u : 0u;
        if (!on) return;
        a.buffer = MTGL_LIST_BUFFER_BASE + id; a.offset = offset; a.stride = 48; a.size = size; a.type = MTGL_TYPE_F32;
    };
    attrib(d.position, true, 0, 3);
    attrib(d.color, r.has_color, 12, 4);
    attrib(d.texcoord, r.has_texcoord, 28, 2);
    attrib(d.normal, r.has_normal, 36, 3);
    d.cur_color[0] = c->current_color.r; d.cur_color[1] = c->current_color.g;
    d.cur_color[2] = c->current_color.b; d.cur_color[3] = c->current_color.a;
    d.cur_texcoord[0] = c->current_texcoord[0]; d.cur_texcoord[1] = c->current_texcoord[1];
    d.cur_normal[0] = c->current_normal[0]; d.cur_normal[1] = c->current_normal[1]; d.cur_normal[2] = c->current_normal[2];
    d.vertex_state = current_state_block(c);
    if (r.has_color) c->current_color = rgba(r.last_color[0], r.last_color[1], r.last_color[2], r.last_color[3]);
    if (r.has_texcoord) { c->current_texcoord[0] = r.last_texcoord[0]; c->current_texcoord[1] = r.last_texcoord[1]; }
    if (r.has_normal) { c->current_normal[0] = r.last_normal[0]; c->current_normal[1] = r.last_normal[1]; c->current_normal[2] = r.last_normal[2]; }
    apply_color_material(c, c->current_color);      /* as after an array draw (draw_arrays_common) */
    d.raster_state = raster_state_block(c);
    c->draws.push_back(d);
    c->list_runs_drawn++;
    if (c->draws.size() > kMaxBatchDraws) flush_batch(c);
    return true;
}

} // namespace mtgl

extern "C" {

void glDrawArrays(GLenum mode, GLint first, GLsizei count) /* gl_api.c:1799-1852 */
{
    MTGL_CTX();
    if (count < 0) { set_error(c, GL_INVALID_VALUE); return; }
    if (!(c->client_state & 1u)) return;
    if (c->bound_array_buffer) {          /* get_array_pointer: a bound buffer without storage draws nothing */
        Buffer *b = get_buffer(c, c->bound_array_buffer);
        if (!b || !b->has_data) return;
    } else if (!c->vertex_pointer.pointer) return;
    draw_arrays_common(c, mode, count, first, 0, nullptr, false, 0);
}

void glDrawElements(GLenum mode, GLsizei count, GLenum type, const GLvoid *indices) /* gl_api.c:1854-1941 */
{
    MTGL_CTX();
    if (count < 0) { set_error(c, GL_INVALID_VALUE); return; }
    if (type != GL_UNSIGNED_BYTE && type != GL_UNSIGNED_SHORT && type != GL_UNSIGNED_INT) {
        set_error(c, GL_INVALID_ENUM);
        return;
    }
    bool have_vertices;
    if (c->bound_array_buffer) {
        Buffer *b = get_buffer(c, c->bound_array_buffer);
        have_vertices = b && b->has_data;
    } else have_vertices = c->vertex_pointer.pointer != nullptr;
    const void *index_data = indices;
    bool in_buffer = false;
    uint64_t offset = 0;
    if (c->bound_element_buffer) {
        Buffer *b = get_buffer(c, c->bound_element_buffer);
        if (!b || !b->has_data) return;
        offset = (uint64_t)(uintptr_t)indices;
        if (offset >= b->size) { set_error(c, GL_INVALID_VALUE); return; }
        index_data = nullptr;               /* resolved lazily: the indices live in the element buffer */
        in_buffer = true;
    }
    if (!(c->client_state & 1u) || !have_vertices || (!in_buffer && !index_data)) return;
    draw_arrays_common(c, mode, count, 0, type, index_data, in_buffer, offset);
}

/* ================================================================ pixel rectangles */
void glReadPixels(GLint x, GLint y, GLsizei width, GLsizei height, GLenum format, GLenum type, GLvoid *pixels)
{ /* gl_api.c:1180-1230.  The rectangle is flipped, cropped and repacked on the device (k_pixels.cu); only its bytes
   * cross PCIe.  Necessarily a synchronisation point. */
    MTGL_CTX();
    if (type != GL_UNSIGNED_BYTE || !pixels) return;
    flush_batch(c);
    if (mtgl_dev_read_pixels(c->dev, x, y, width, height, format, pixels) != MTGL_OK) set_error(c, GL_INVALID_OPERATION);
}

/* glDrawPixels (gl_api.c:1286-1373) on the device: the queued draws go first (glFlush, no wait), then the rectangle
 * with the per-fragment state sampled now; the call returns as soon as the pixels have been handed over. */
void glDrawPixels(GLsizei width, GLsizei height, GLenum format, GLenum type, const GLvoid *pixels)
{
    MTGL_CTX();
    if (type != GL_UNSIGNED_BYTE || !pixels || !c->raster_pos_valid) return;
    if (width <= 0 || height <= 0) return;
    flush_batch(c);
    mtgl_pixel_rect r;
    std::memset(&r, 0, sizeof r);
    r.x = c->raster_pos_x; r.y = c->raster_pos_y; r.width = width; r.height = height;
    r.format = format;
    r.caps = c->caps & (MTGL_CAP_ALPHA_TEST | MTGL_CAP_DEPTH_TEST | MTGL_CAP_BLEND);
    r.alpha_func = c->alpha_func; r.alpha_ref = c->alpha_ref;
    r.depth_func = c->depth_func; r.depth_mask = c->depth_mask ? 1u : 0u;
    r.blend_src = c->blend_src; r.blend_dst = c->blend_dst;
    if (mtgl_dev_draw_pixels(c->dev, &r, pixels) != MTGL_OK) set_error(c, GL_INVALID_OPERATION);
}

static void raster_pos(GLState *c, float x, float y, float z) /* gl_api.c:1375-1424 */
{
    float v[4] = { x, y, z, 1.0f }, eye[4], clip[4];
    mat_vec(c->modelview[c->modelview_depth], v, eye);
    mat_vec(c->projection[c->projection_depth], eye, clip);
    if (clip[3] <= 0.0f) { c->raster_pos_valid = GL_FALSE; return; }
    float nx = clip[0] / clip[3], ny = clip[1] / clip[3];
    int32_t sx = (int32_t)((nx + 1.0f) * 0.5f * c->viewport_w + c->viewport_x);
    int32_t sy = (int32_t)((1.0f - ny) * 0.5f * c->viewport_h + c->viewport_y);
    c->raster_pos_x = sx;
    c->raster_pos_y = c->fb_height - 1 - sy;
    c->raster_pos_valid = GL_TRUE;
}

void glRasterPos2i(GLint x, GLint y) { MTGL_CTX(); raster_pos(c, (float)x, (float)y, 0.0f); }
void glRasterPos3f(GLfloat x, GLfloat y, GLfloat z) { MTGL_CTX(); raster_pos(c, x, y, z); }
void glRasterPos2f(GLfloat x, GLfloat y) { glRasterPos3f(x, y, 0.0f); }

} // extern "C"
