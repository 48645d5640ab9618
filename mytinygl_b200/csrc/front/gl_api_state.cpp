/*
 * gl_api_state.cpp -- the GL state machine: capabilities, matrix stacks, lighting, fog, texture
 * and buffer objects, display lists and glGet*.  Pure host bookkeeping, kept identical in
 * behaviour (accepted enums, error codes, defaults, clamping) to the reference's src/gl_api.c
 * so that the library stays a drop-in; every setter that can influence rendering marks the
 * state snapshot dirty so that the next vertex / glEnd captures a fresh mtgl_state block.
 */
#include "front_internal.h"

#include <cmath>
#include <cstdlib>
#include <cstring>

using namespace mtgl;

namespace mtgl {

bool record(GLState *c, const ListCmd &cmd) /* record_cmd, lists.c:185-197 */
{
    if (c->list_index == 0) return false;
    DisplayList *l = get_list(c, c->list_index);
    if (l) l->cmds.push_back(cmd);
    return c->list_mode == GL_COMPILE;
}

static ListCmd make_cmd(ListOp op, GLenum e0 = 0, GLenum e1 = 0)
{
    ListCmd cmd;
    std::memset(&cmd, 0, sizeof cmd);
    cmd.op = op; cmd.e[0] = e0; cmd.e[1] = e1;
    return cmd;
}

static uint32_t cap_bit(GLenum cap) /* cap_to_flag, gl_api.c:351-367 */
{
    switch (cap) {
    case GL_DEPTH_TEST: return MTGL_CAP_DEPTH_TEST;
    case GL_CULL_FACE: return MTGL_CAP_CULL_FACE;
    case GL_BLEND: return MTGL_CAP_BLEND;
    case GL_TEXTURE_2D: return MTGL_CAP_TEXTURE_2D;
    case GL_LIGHTING: return MTGL_CAP_LIGHTING;
    case GL_FOG: return MTGL_CAP_FOG;
    case GL_NORMALIZE: return MTGL_CAP_NORMALIZE;
    case GL_COLOR_MATERIAL: return MTGL_CAP_COLOR_MATERIAL;
    case GL_ALPHA_TEST: return MTGL_CAP_ALPHA_TEST;
    case GL_SCISSOR_TEST: return MTGL_CAP_SCISSOR_TEST;
    case GL_STENCIL_TEST: return MTGL_CAP_STENCIL_TEST;
    default: return 0;
    }
}

static void post_multiply(GLState *c, const float *m)
{
    float *cur = current_matrix(c);
    mat_mul(cur, m, cur);
    mark_matrix_dirty(c);
}

static bool is_compare_func(GLenum f) { return f >= GL_NEVER && f <= GL_ALWAYS; }

} // namespace mtgl

extern "C" {

/* ================================================================ capabilities */
void glEnable(GLenum cap) /* gl_api.c:371-388 */
{
    MTGL_CTX();
    if (record(c, make_cmd(OP_ENABLE, cap))) return;
    uint32_t bit = cap_bit(cap);
    if (bit) { c->caps |= bit; mark_state_dirty(c); return; }
    if (cap >= GL_LIGHT0 && cap <= GL_LIGHT7) { c->lights[cap - GL_LIGHT0].enabled = GL_TRUE; mark_state_dirty(c); return; }
    set_error(c, GL_INVALID_ENUM);
}

void glDisable(GLenum cap) /* gl_api.c:390-407 */
{
    MTGL_CTX();
    if (record(c, make_cmd(OP_DISABLE, cap))) return;
    uint32_t bit = cap_bit(cap);
    if (bit) { c->caps &= ~bit; mark_state_dirty(c); return; }
    if (cap >= GL_LIGHT0 && cap <= GL_LIGHT7) { c->lights[cap - GL_LIGHT0].enabled = GL_FALSE; mark_state_dirty(c); return; }
    set_error(c, GL_INVALID_ENUM);
}

GLboolean glIsEnabled(GLenum cap) /* gl_api.c:3151-3174 */
{
    MTGL_CTX_RET(GL_FALSE);
    uint32_t bit = cap_bit(cap);
    if (bit) return (c->caps & bit) ? GL_TRUE : GL_FALSE;
    if (cap >= GL_LIGHT0 && cap <= GL_LIGHT7) return c->lights[cap - GL_LIGHT0].enabled;
    set_error(c, GL_INVALID_ENUM);
    return GL_FALSE;
}

void glClearColor(GLclampf r, GLclampf g, GLclampf b, GLclampf a) { MTGL_CTX(); c->clear_color = rgba(r, g, b, a); }
void glClearDepth(GLclampd depth) { MTGL_CTX(); c->clear_depth = depth; }
void glClearStencil(GLint s) { MTGL_CTX(); c->stencil_clear = s; }

void glViewport(GLint x, GLint y, GLsizei w, GLsizei h)
{
    MTGL_CTX();
    c->viewport_x = x; c->viewport_y = y; c->viewport_w = w; c->viewport_h = h;
    mark_state_dirty(c);
}

void glScissor(GLint x, GLint y, GLsizei w, GLsizei h) /* gl_api.c:1044-1055 */
{
    MTGL_CTX();
    if (w < 0 || h < 0) { set_error(c, GL_INVALID_VALUE); return; }
    c->scissor_x = x; c->scissor_y = y; c->scissor_w = w; c->scissor_h = h;
    mark_state_dirty(c);
}

void glDepthRange(GLclampd n, GLclampd f) /* gl_api.c:1057-1067 */
{
    MTGL_CTX();
    if (n < 0.0) n = 0.0;
    if (n > 1.0) n = 1.0;
    if (f < 0.0) f = 0.0;
    if (f > 1.0) f = 1.0;
    c->depth_near = n; c->depth_far = f;
    mark_state_dirty(c);
}

/* ================================================================ matrix stacks (gl_api.c:483-610) */
void glMatrixMode(GLenum mode)
{
    MTGL_CTX();
    if (record(c, make_cmd(OP_MATRIX_MODE, mode))) return;
    c->matrix_mode = mode;
}

void glLoadIdentity(void)
{
    MTGL_CTX();
    if (record(c, make_cmd(OP_LOAD_IDENTITY))) return;
    mat_identity(current_matrix(c));
    mark_matrix_dirty(c);
}

void glPushMatrix(void)
{
    MTGL_CTX();
    if (record(c, make_cmd(OP_PUSH_MATRIX))) return;
    GLint *depth = current_depth(c);
    if (*depth >= kMatrixStackDepth - 1) { set_error(c, GL_STACK_OVERFLOW); return; }
    float *src = current_matrix(c);
    (*depth)++;
    std::memcpy(current_matrix(c), src, 64);
}

void glPopMatrix(void)
{
    MTGL_CTX();
    if (record(c, make_cmd(OP_POP_MATRIX))) return;
    GLint *depth = current_depth(c);
    if (*depth <= 0) { set_error(c, GL_STACK_UNDERFLOW); return; }
    (*depth)--;
    mark_matrix_dirty(c);
}

void glLoadMatrixf(const GLfloat *m)
{
    MTGL_CTX();
    ListCmd cmd = make_cmd(OP_LOAD_MATRIX);
    std::memcpy(cmd.f, m, 64);
    if (record(c, cmd)) return;
    std::memcpy(current_matrix(c), m, 64);
    mark_matrix_dirty(c);
}

void glMultMatrixf(const GLfloat *m)
{
    MTGL_CTX();
    ListCmd cmd = make_cmd(OP_MULT_MATRIX);
    std::memcpy(cmd.f, m, 64);
    if (record(c, cmd)) return;
    post_multiply(c, m);
}

void glTranslatef(GLfloat x, GLfloat y, GLfloat z)
{
    MTGL_CTX();
    ListCmd cmd = make_cmd(OP_TRANSLATE);
    cmd.f[0] = x; cmd.f[1] = y; cmd.f[2] = z;
    if (record(c, cmd)) return;
    float t[16];
    mat_identity(t);
    t[12] = x; t[13] = y; t[14] = z;
    post_multiply(c, t);
}

void glScalef(GLfloat x, GLfloat y, GLfloat z)
{
    MTGL_CTX();
    ListCmd cmd = make_cmd(OP_SCALE);
    cmd.f[0] = x; cmd.f[1] = y; cmd.f[2] = z;
    if (record(c, cmd)) return;
    float s[16];
    mat_identity(s);
    s[0] = x; s[5] = y; s[10] = z;
    post_multiply(c, s);
}

void glRotatef(GLfloat angle, GLfloat x, GLfloat y, GLfloat z) /* mat4_rotate, graphics.h:172-186 */
{
    MTGL_CTX();
    ListCmd cmd = make_cmd(OP_ROTATE);
    cmd.f[0] = angle; cmd.f[1] = x; cmd.f[2] = y; cmd.f[3] = z;
    if (record(c, cmd)) return;
    float rad = angle * 3.14159265358979f / 180.0f;
    float cs = cosf(rad), sn = sinf(rad);
    float len = sqrtf(x * x + y * y + z * z);
    float r[16];
    mat_identity(r);
    if (len != 0) {
        x /= len; y /= len; z /= len;
        float k = 1 - cs;
        r[0] = x * x * k + cs;      r[1] = y * x * k + z * sn;  r[2] = x * z * k - y * sn;
        r[4] = x * y * k - z * sn;  r[5] = y * y * k + cs;      r[6] = y * z * k + x * sn;
        r[8] = x * z * k + y * sn;  r[9] = y * z * k - x * sn;  r[10] = z * z * k + cs;
    }
    post_multiply(c, r);
}

void glOrtho(GLdouble l, GLdouble r, GLdouble b, GLdouble t, GLdouble n, GLdouble f) /* graphics.h:188-205 */
{
    MTGL_CTX();
    ListCmd cmd = make_cmd(OP_ORTHO);
    cmd.d[0] = l; cmd.d[1] = r; cmd.d[2] = b; cmd.d[3] = t; cmd.d[4] = n; cmd.d[5] = f;
    if (record(c, cmd)) return;
    float left = (float)l, right = (float)r, bottom = (float)b, top = (float)t, zn = (float)n, zf = (float)f;
    float rl = right - left, tb = top - bottom, fn = zf - zn;
    float m[16];
    mat_identity(m);
    if (!(rl == 0.0f || tb == 0.0f || fn == 0.0f)) {
        m[0] = 2.0f / rl; m[5] = 2.0f / tb; m[10] = -2.0f / fn;
        m[12] = -(right + left) / rl; m[13] = -(top + bottom) / tb; m[14] = -(zf + zn) / fn;
    }
    post_multiply(c, m);
}

void glFrustum(GLdouble l, GLdouble r, GLdouble b, GLdouble t, GLdouble n, GLdouble f) /* graphics.h:207-224 */
{
    MTGL_CTX();
    ListCmd cmd = make_cmd(OP_FRUSTUM);
    cmd.d[0] = l; cmd.d[1] = r; cmd.d[2] = b; cmd.d[3] = t; cmd.d[4] = n; cmd.d[5] = f;
    if (record(c, cmd)) return;
    float left = (float)l, right = (float)r, bottom = (float)b, top = (float)t, zn = (float)n, zf = (float)f;
    float rl = right - left, tb = top - bottom, fn = zf - zn;
    float m[16];
    mat_identity(m);
    if (!(rl == 0.0f || tb == 0.0f || fn == 0.0f)) {
        std::memset(m, 0, sizeof m);
        m[0] = 2.0f * zn / rl; m[5] = 2.0f * zn / tb;
        m[8] = (right + left) / rl; m[9] = (top + bottom) / tb; m[10] = -(zf + zn) / fn; m[11] = -1;
        m[14] = -2.0f * zf * zn / fn;
    }
    post_multiply(c, m);
}

/* ================================================================ textures (gl_api.c:756-885, textures.c:58-269) */
void glGenTextures(GLsizei n, GLuint *ids)
{
    MTGL_CTX();
    if (n < 0) { set_error(c, GL_INVALID_VALUE); return; }
    for (GLsizei i = 0; i < n; i++) {
        GLuint id = 0;
        for (size_t k = 0; k < c->textures.size() && !id; k++)
            if (!c->textures[k].allocated) id = (GLuint)k + 1;          /* slot reuse, textures.c:61-67 */
        if (!id && c->textures.size() < (size_t)kMaxTextures) {
            c->textures.emplace_back();
            id = (GLuint)c->textures.size();
        }
        if (id) {
            c->textures[id - 1] = Texture();
            c->textures[id - 1].allocated = true;
        }
        ids[i] = id;
    }
}

void glDeleteTextures(GLsizei n, const GLuint *ids)
{
    MTGL_CTX();
    if (n < 0) { set_error(c, GL_INVALID_VALUE); return; }
    for (GLsizei i = 0; i < n; i++) {
        if (ids[i] == c->bound_texture_2d) { c->bound_texture_2d = 0; mark_state_dirty(c); }
        Texture *t = get_texture(c, ids[i]);
        if (!t) continue;
        flush_batch(c);                       /* queued draws may still sample it */
        mtgl_dev_texture_delete(c->dev, ids[i]);
        *t = Texture();
    }
}

void glBindTexture(GLenum target, GLuint id)
{
    MTGL_CTX();
    if (record(c, make_cmd(OP_BIND_TEXTURE, target, id))) return;
    if (target != GL_TEXTURE_2D) { set_error(c, GL_INVALID_ENUM); return; }
    c->bound_texture_2d = id;
    mark_state_dirty(c);
}

void glTexImage2D(GLenum target, GLint level, GLint internalformat, GLsizei w, GLsizei h, GLint border,
                  GLenum format, GLenum type, const GLvoid *pixels) /* gl_api.c:794-834 */
{
    (void)internalformat; (void)border;
    MTGL_CTX();
    if (target != GL_TEXTURE_2D) { set_error(c, GL_INVALID_ENUM); return; }
    if (level != 0 || type != GL_UNSIGNED_BYTE) { set_error(c, GL_INVALID_VALUE); return; }
    if (w < 0 || h < 0) { set_error(c, GL_INVALID_VALUE); return; }
    Texture *t = get_texture(c, c->bound_texture_2d);
    if (!t) return;
    int comps = (format == GL_RGBA) ? 4 : (format == GL_RGB) ? 3 : (format == GL_LUMINANCE) ? 1 :
                (format == GL_LUMINANCE_ALPHA) ? 2 : 0;
    if (!comps) { set_error(c, GL_INVALID_ENUM); return; }
    if (w <= 0 || h <= 0 || w > kMaxTextureSize || h > kMaxTextureSize) return;   /* textures.c:143-144: upload refused */
    const uint8_t *src = (const uint8_t *)pixels;
    size_t n = (size_t)w * h;
    std::vector<uint32_t> px(n);
    for (size_t i = 0; i < n; i++) {          /* textures.c:164-171, 198-204, 231-234, 261-266 */
        uint32_t r, g, b, a = 0xFF;
        switch (comps) {
        case 4: r = src[i * 4]; g = src[i * 4 + 1]; b = src[i * 4 + 2]; a = src[i * 4 + 3]; break;
        case 3: r = src[i * 3]; g = src[i * 3 + 1]; b = src[i * 3 + 2]; break;
        case 2: r = g = b = src[i * 2]; a = src[i * 2 + 1]; break;
        default: r = g = b = src[i]; break;
        }
        px[i] = (a << 24) | (b << 16) | (g << 8) | r;
    }
    flush_batch(c);                           /* queued draws sample the previous image */
    int rc = mtgl_dev_texture_image(c->dev, c->bound_texture_2d, w, h, px.data());
    if (rc != MTGL_OK) { set_error(c, GL_OUT_OF_MEMORY); return; }
    t->pixels.swap(px);
    t->width = w; t->height = h;
    mark_state_dirty(c);
}

void glTexParameteri(GLenum target, GLenum pname, GLint param) /* gl_api.c:836-885 */
{
    MTGL_CTX();
    if (target != GL_TEXTURE_2D) { set_error(c, GL_INVALID_ENUM); return; }
    Texture *t = get_texture(c, c->bound_texture_2d);
    if (!t) return;
    bool plain = (param == GL_NEAREST || param == GL_LINEAR);
    bool mip = (param == GL_NEAREST_MIPMAP_NEAREST || param == GL_LINEAR_MIPMAP_NEAREST ||
                param == GL_NEAREST_MIPMAP_LINEAR || param == GL_LINEAR_MIPMAP_LINEAR);
    bool wrap = (param == GL_REPEAT || param == GL_CLAMP || param == GL_CLAMP_TO_EDGE);
    switch (pname) {
    case GL_TEXTURE_MIN_FILTER: if (!plain && !mip) { set_error(c, GL_INVALID_ENUM); return; } t->min_filter = param; break;
    case GL_TEXTURE_MAG_FILTER: if (!plain) { set_error(c, GL_INVALID_ENUM); return; } t->mag_filter = param; break;
    case GL_TEXTURE_WRAP_S: if (!wrap) { set_error(c, GL_INVALID_ENUM); return; } t->wrap_s = param; break;
    case GL_TEXTURE_WRAP_T: if (!wrap) { set_error(c, GL_INVALID_ENUM); return; } t->wrap_t = param; break;
    default: set_error(c, GL_INVALID_ENUM); return;
    }
    mark_state_dirty(c);
}

void glTexEnvi(GLenum target, GLenum pname, GLint param) /* gl_api.c:3193-3219 */
{
    MTGL_CTX();
    if (target != GL_TEXTURE_ENV || pname != GL_TEXTURE_ENV_MODE) { set_error(c, GL_INVALID_ENUM); return; }
    switch (param) {
    case GL_MODULATE: case GL_DECAL: case GL_REPLACE: case GL_BLEND: case GL_ADD:
        c->tex_env_mode = (GLenum)param;
        mark_state_dirty(c);
        break;
    default: set_error(c, GL_INVALID_ENUM); break;
    }
}

void glTexEnvf(GLenum target, GLenum pname, GLfloat param) { glTexEnvi(target, pname, (GLint)param); }

void glTexEnvfv(GLenum target, GLenum pname, const GLfloat *params)
{
    MTGL_CTX();
    if (!params) return;
    if (target != GL_TEXTURE_ENV) { set_error(c, GL_INVALID_ENUM); return; }
    if (pname == GL_TEXTURE_ENV_MODE) glTexEnvi(target, pname, (GLint)params[0]);
    else if (pname == GL_TEXTURE_ENV_COLOR) { c->tex_env_color = rgba(params[0], params[1], params[2], params[3]); mark_state_dirty(c); }
    else set_error(c, GL_INVALID_ENUM);
}

GLboolean glIsTexture(GLuint id)
{
    MTGL_CTX_RET(GL_FALSE);
    return (id != 0 && get_texture(c, id)) ? GL_TRUE : GL_FALSE;
}

/* ================================================================ buffer objects (gl_api.c:1516-1616, vbo.c) */
void glGenBuffers(GLsizei n, GLuint *ids)
{
    MTGL_CTX();
    if (n < 0) { set_error(c, GL_INVALID_VALUE); return; }
    for (GLsizei i = 0; i < n; i++) {
        GLuint id = 0;
        for (size_t k = 0; k < c->buffers.size() && !id; k++)
            if (!c->buffers[k].allocated) id = (GLuint)k + 1;
        if (!id && c->buffers.size() < (size_t)kMaxBuffers) {
            c->buffers.emplace_back();
            id = (GLuint)c->buffers.size();
        }
        if (id) {
            c->buffers[id - 1] = Buffer();
            c->buffers[id - 1].allocated = true;
        }
        ids[i] = id;
    }
}

void glDeleteBuffers(GLsizei n, const GLuint *ids)
{
    MTGL_CTX();
    if (n < 0) { set_error(c, GL_INVALID_VALUE); return; }
    for (GLsizei i = 0; i < n; i++) {
        if (ids[i] == c->bound_array_buffer) c->bound_array_buffer = 0;
        if (ids[i] == c->bound_element_buffer) c->bound_element_buffer = 0;
        Buffer *b = get_buffer(c, ids[i]);
        if (!b) continue;
        flush_batch(c);
        mtgl_dev_buffer_delete(c->dev, ids[i]);
        *b = Buffer();
    }
}

void glBindBuffer(GLenum target, GLuint id)
{
    MTGL_CTX();
    if (target == GL_ARRAY_BUFFER) c->bound_array_buffer = id;
    else if (target == GL_ELEMENT_ARRAY_BUFFER) c->bound_element_buffer = id;
    else set_error(c, GL_INVALID_ENUM);
}

static GLuint buffer_for_target(GLState *c, GLenum target, bool *ok)
{
    *ok = true;
    if (target == GL_ARRAY_BUFFER) return c->bound_array_buffer;
    if (target == GL_ELEMENT_ARRAY_BUFFER) return c->bound_element_buffer;
    *ok = false;
    return 0;
}

static void buffer_data_common(GLenum target, GLsizeiptr size, const GLvoid *data, GLenum usage, bool pinned_async);

void glBufferData(GLenum target, GLsizeiptr size, const GLvoid *data, GLenum usage)
{
    buffer_data_common(target, size, data, usage, false);
}

/* include/mtgl_context.h: the source is page-locked and stays unchanged until glFinish() -- the copy is only queued */
void mtglBufferDataPinned(unsigned target, long size, const void *pinned_data, unsigned usage)
{
    buffer_data_common((GLenum)target, (GLsizeiptr)size, pinned_data, (GLenum)usage, pinned_data != nullptr);
}

static void buffer_data_common(GLenum target, GLsizeiptr size, const GLvoid *data, GLenum usage, bool pinned_async)
{
    MTGL_CTX();
    if (size < 0) { set_error(c, GL_INVALID_VALUE); return; }
    bool ok;
    GLuint id = buffer_for_target(c, target, &ok);
    if (!ok) { set_error(c, GL_INVALID_ENUM); return; }
    Buffer *b = get_buffer(c, id);
    if (!b) return;
    flush_batch(c);                           /* queued draws read the previous contents */
    b->usage = usage;
    b->peeks_frozen = false;                  /* the contents are known again */
    if (size <= 0) {                          /* vbo.c:126-134 */
        b->data.clear(); b->data.shrink_to_fit();
        b->has_data = false; b->host_valid = false; b->size = 0;
        b->peeks.clear();
        mtgl_dev_buffer_data(c->dev, id, 0, nullptr);
        return;
    }
    if (!data) b->peeks.clear();
    else {                                    /* the bytes the host has looked at before, from the new contents */
        size_t keep = 0;
        for (Buffer::Peek &p : b->peeks)
            if (p.off + p.n <= (uint64_t)size) {
                std::memcpy(p.raw, (const uint8_t *)data + p.off, p.n);
                b->peeks[keep++] = p;
            }
        b->peeks.resize(keep);
    }
    /* fresh storage; contents undefined when data == NULL.  The HBM mirror is filled straight from the
     * caller's memory (a pinned pointer is DMA'd without a bounce); only small buffers also keep a host copy. */
    b->has_data = true;
    b->size = (uint64_t)size;
    if ((uint64_t)size <= kHostMirrorLimit) {
        b->data.resize((size_t)size);
        if (data) std::memcpy(b->data.data(), data, (size_t)size);
        b->host_valid = true;
    } else {
        b->data.clear(); b->data.shrink_to_fit();
        b->host_valid = false;
    }
    const int rc = pinned_async ? mtgl_dev_buffer_data_pinned(c->dev, id, (uint64_t)size, data) : mtgl_dev_buffer_data(c->dev, id, (uint64_t)size, data);
    if (rc != MTGL_OK) {
        set_error(c, GL_OUT_OF_MEMORY);
        b->has_data = false; b->host_valid = false; b->size = 0;
    }
}

void glBufferSubData(GLenum target, GLintptr offset, GLsizeiptr size, const GLvoid *data)
{
    MTGL_CTX();
    if (offset < 0 || size < 0) { set_error(c, GL_INVALID_VALUE); return; }
    bool ok;
    GLuint id = buffer_for_target(c, target, &ok);
    if (!ok) { set_error(c, GL_INVALID_ENUM); return; }
    Buffer *b = get_buffer(c, id);
    if (!b) { set_error(c, GL_INVALID_OPERATION); return; }
    if (!b->has_data || !data || (uint64_t)offset + (uint64_t)size > b->size) { set_error(c, GL_INVALID_VALUE); return; }
    flush_batch(c);
    if (b->host_valid) std::memcpy(b->data.data() + offset, data, (size_t)size);
    for (Buffer::Peek &p : b->peeks) {        /* patch the overlapping bytes */
        const uint64_t lo = std::max<uint64_t>(p.off, (uint64_t)offset), hi = std::min<uint64_t>(p.off + p.n, (uint64_t)offset + (uint64_t)size);
        if (lo < hi) std::memcpy(p.raw + (lo - p.off), (const uint8_t *)data + (lo - (uint64_t)offset), (size_t)(hi - lo));
    }
    if (mtgl_dev_buffer_sub_data(c->dev, id, (uint64_t)offset, (uint64_t)size, data) != MTGL_OK)
        set_error(c, GL_INVALID_VALUE);
}

GLboolean glIsBuffer(GLuint id)
{
    MTGL_CTX_RET(GL_FALSE);
    return (id != 0 && get_buffer(c, id)) ? GL_TRUE : GL_FALSE;
}

/* ================================================================ client arrays (gl_api.c:1620-1742) */
static uint32_t client_bit(GLenum array)
{
    switch (array) {
    case GL_VERTEX_ARRAY: return 1u;
    case GL_COLOR_ARRAY: return 2u;
    case GL_TEXTURE_COORD_ARRAY: return 4u;
    case GL_NORMAL_ARRAY: return 8u;
    default: return 0;
    }
}

void glEnableClientState(GLenum array)
{
    MTGL_CTX();
    uint32_t b = client_bit(array);
    if (b) c->client_state |= b; else set_error(c, GL_INVALID_ENUM);
}

void glDisableClientState(GLenum array)
{
    MTGL_CTX();
    uint32_t b = client_bit(array);
    if (b) c->client_state &= ~b; else set_error(c, GL_INVALID_ENUM);
}

static void set_pointer(GLState *c, ArrayPointer *a, GLint size, GLint lo, GLint hi, GLenum type, GLsizei stride, const void *p)
{
    if (size < lo || size > hi) { set_error(c, GL_INVALID_VALUE); return; }
    if (type != GL_FLOAT && type != GL_UNSIGNED_BYTE) { set_error(c, GL_INVALID_ENUM); return; }
    if (stride < 0) { set_error(c, GL_INVALID_VALUE); return; }
    a->size = size; a->type = type; a->stride = stride; a->pointer = p;
}

void glVertexPointer(GLint size, GLenum type, GLsizei stride, const GLvoid *p) { MTGL_CTX(); set_pointer(c, &c->vertex_pointer, size, 2, 4, type, stride, p); }
void glColorPointer(GLint size, GLenum type, GLsizei stride, const GLvoid *p) { MTGL_CTX(); set_pointer(c, &c->color_pointer, size, 3, 4, type, stride, p); }
void glTexCoordPointer(GLint size, GLenum type, GLsizei stride, const GLvoid *p) { MTGL_CTX(); set_pointer(c, &c->texcoord_pointer, size, 1, 4, type, stride, p); }
void glNormalPointer(GLenum type, GLsizei stride, const GLvoid *p) { MTGL_CTX(); set_pointer(c, &c->normal_pointer, 3, 3, 3, type, stride, p); }

/* ================================================================ lighting (gl_api.c:1945-2262) */
/* floats a light / material parameter is made of: only those are read from the caller's array (a 3-float spot
 * direction or a 1-float shininess must not be read as four) */
static int light_param_count(GLenum pname)
{
    switch (pname) {
    case GL_AMBIENT: case GL_DIFFUSE: case GL_SPECULAR: case GL_POSITION: return 4;
    case GL_SPOT_DIRECTION: return 3;
    default: return 1;
    }
}

static int material_param_count(GLenum pname)
{
    switch (pname) {
    case GL_AMBIENT: case GL_DIFFUSE: case GL_SPECULAR: case GL_EMISSION: case GL_AMBIENT_AND_DIFFUSE: return 4;
    default: return 1;
    }
}

void glLightfv(GLenum light, GLenum pname, const GLfloat *params)
{
    MTGL_CTX();
    if (!params) return;
    GLfloat p[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
    std::memcpy(p, params, sizeof(GLfloat) * (size_t)light_param_count(pname));
    ListCmd cmd = make_cmd(OP_LIGHTFV, light, pname);
    std::memcpy(cmd.f, p, 16);
    if (record(c, cmd)) return;
    if (light < GL_LIGHT0 || light > GL_LIGHT7) { set_error(c, GL_INVALID_ENUM); return; }
    Light &l = c->lights[light - GL_LIGHT0];
    switch (pname) {
    case GL_AMBIENT: l.ambient = rgba(p[0], p[1], p[2], p[3]); break;
    case GL_DIFFUSE: l.diffuse = rgba(p[0], p[1], p[2], p[3]); break;
    case GL_SPECULAR: l.specular = rgba(p[0], p[1], p[2], p[3]); break;
    case GL_POSITION: mat_vec(c->modelview[c->modelview_depth], p, l.position); break;   /* eye space at call time */
    case GL_SPOT_DIRECTION: {
        float in[4] = { p[0], p[1], p[2], 0.0f }, out[4];
        mat_vec(c->modelview[c->modelview_depth], in, out);
        l.spot_direction[0] = out[0]; l.spot_direction[1] = out[1]; l.spot_direction[2] = out[2];
        break;
    }
    case GL_SPOT_EXPONENT: l.spot_exponent = p[0]; break;
    case GL_SPOT_CUTOFF: l.spot_cutoff = p[0]; break;
    case GL_CONSTANT_ATTENUATION: l.att_constant = p[0]; break;
    case GL_LINEAR_ATTENUATION: l.att_linear = p[0]; break;
    case GL_QUADRATIC_ATTENUATION: l.att_quadratic = p[0]; break;
    default: break;
    }
    mark_state_dirty(c);
}

void glLightf(GLenum light, GLenum pname, GLfloat param)
{
    MTGL_CTX();
    ListCmd cmd = make_cmd(OP_LIGHTF, light, pname);
    cmd.f[0] = param;
    if (record(c, cmd)) return;
    GLfloat p[4] = { param, param, param, param };
    glLightfv(light, pname, p);
}

void glLighti(GLenum light, GLenum pname, GLint param) { glLightf(light, pname, (GLfloat)param); }

void glLightiv(GLenum light, GLenum pname, const GLint *p)
{
    GLfloat f[4] = { (GLfloat)p[0], (GLfloat)p[1], (GLfloat)p[2], (GLfloat)p[3] };
    glLightfv(light, pname, f);
}

void glMaterialfv(GLenum face, GLenum pname, const GLfloat *params)
{
    MTGL_CTX();
    if (!params) return;
    GLfloat p[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
    std::memcpy(p, params, sizeof(GLfloat) * (size_t)material_param_count(pname));
    ListCmd cmd = make_cmd(OP_MATERIALFV, face, pname);
    std::memcpy(cmd.f, p, 16);
    if (record(c, cmd)) return;
    if (face != GL_FRONT && face != GL_BACK && face != GL_FRONT_AND_BACK) { set_error(c, GL_INVALID_ENUM); return; }
    Material *m[2] = { (face != GL_BACK) ? &c->material_front : nullptr, (face != GL_FRONT) ? &c->material_back : nullptr };
    Rgba v = rgba(p[0], p[1], p[2], p[3]);
    bool known = true;
    for (Material *mat : m) {
        if (!mat) continue;
        switch (pname) {
        case GL_AMBIENT: mat->ambient = v; break;
        case GL_DIFFUSE: mat->diffuse = v; break;
        case GL_SPECULAR: mat->specular = v; break;
        case GL_EMISSION: mat->emission = v; break;
        case GL_SHININESS: mat->shininess = p[0]; break;
        case GL_AMBIENT_AND_DIFFUSE: mat->ambient = v; mat->diffuse = v; break;
        default: known = false; break;
        }
    }
    if (!known) set_error(c, GL_INVALID_ENUM);
    mark_state_dirty(c);
}

void glMaterialf(GLenum face, GLenum pname, GLfloat param)
{
    MTGL_CTX();
    ListCmd cmd = make_cmd(OP_MATERIALF, face, pname);
    cmd.f[0] = param;
    if (record(c, cmd)) return;
    GLfloat p[4] = { param, param, param, param };
    glMaterialfv(face, pname, p);
}

void glMateriali(GLenum face, GLenum pname, GLint param) { glMaterialf(face, pname, (GLfloat)param); }

void glMaterialiv(GLenum face, GLenum pname, const GLint *p)
{
    GLfloat f[4] = { (GLfloat)p[0], (GLfloat)p[1], (GLfloat)p[2], (GLfloat)p[3] };
    glMaterialfv(face, pname, f);
}

void glLightModelfv(GLenum pname, const GLfloat *p)
{
    MTGL_CTX();
    switch (pname) {
    case GL_LIGHT_MODEL_AMBIENT: c->light_model_ambient = rgba(p[0], p[1], p[2], p[3]); break;
    case GL_LIGHT_MODEL_LOCAL_VIEWER: c->light_model_local_viewer = (GLboolean)(p[0] != 0.0f); break;
    case GL_LIGHT_MODEL_TWO_SIDE: c->light_model_two_side = (GLboolean)(p[0] != 0.0f); break;
    default: break;
    }
    mark_state_dirty(c);
}

void glLightModelf(GLenum pname, GLfloat param)
{
    GLfloat p[4] = { param, param, param, param };
    glLightModelfv(pname, p);
}

void glLightModeli(GLenum pname, GLint param) { glLightModelf(pname, (GLfloat)param); }

void glLightModeliv(GLenum pname, const GLint *p)
{
    GLfloat f[4] = { (GLfloat)p[0], (GLfloat)p[1], (GLfloat)p[2], (GLfloat)p[3] };
    glLightModelfv(pname, f);
}

void glColorMaterial(GLenum face, GLenum mode)
{
    MTGL_CTX();
    if (face != GL_FRONT && face != GL_BACK && face != GL_FRONT_AND_BACK) { set_error(c, GL_INVALID_ENUM); return; }
    if (mode != GL_EMISSION && mode != GL_AMBIENT && mode != GL_DIFFUSE && mode != GL_SPECULAR &&
        mode != GL_AMBIENT_AND_DIFFUSE) { set_error(c, GL_INVALID_ENUM); return; }
    c->color_material_face = face;
    c->color_material_mode = mode;
    mark_state_dirty(c);
}

void glShadeModel(GLenum mode)
{
    MTGL_CTX();
    if (record(c, make_cmd(OP_SHADE_MODEL, mode))) return;
    if (mode != GL_FLAT && mode != GL_SMOOTH && mode != GL_PHONG) { set_error(c, GL_INVALID_ENUM); return; }
    c->shade_model = mode;
    mark_state_dirty(c);
}

static void put4(GLfloat *o, const Rgba &v) { o[0] = v.r; o[1] = v.g; o[2] = v.b; o[3] = v.a; }

void glGetLightfv(GLenum light, GLenum pname, GLfloat *o)
{
    MTGL_CTX();
    if (!o) return;
    if (light < GL_LIGHT0 || light > GL_LIGHT7) { set_error(c, GL_INVALID_ENUM); return; }
    const Light &l = c->lights[light - GL_LIGHT0];
    switch (pname) {
    case GL_AMBIENT: put4(o, l.ambient); break;
    case GL_DIFFUSE: put4(o, l.diffuse); break;
    case GL_SPECULAR: put4(o, l.specular); break;
    case GL_POSITION: std::memcpy(o, l.position, 16); break;
    case GL_SPOT_DIRECTION: std::memcpy(o, l.spot_direction, 12); break;
    case GL_SPOT_EXPONENT: o[0] = l.spot_exponent; break;
    case GL_SPOT_CUTOFF: o[0] = l.spot_cutoff; break;
    case GL_CONSTANT_ATTENUATION: o[0] = l.att_constant; break;
    case GL_LINEAR_ATTENUATION: o[0] = l.att_linear; break;
    case GL_QUADRATIC_ATTENUATION: o[0] = l.att_quadratic; break;
    default: set_error(c, GL_INVALID_ENUM); break;
    }
}

void glGetMaterialfv(GLenum face, GLenum pname, GLfloat *o)
{
    MTGL_CTX();
    if (!o) return;
    const Material *m;
    if (face == GL_FRONT) m = &c->material_front;
    else if (face == GL_BACK) m = &c->material_back;
    else { set_error(c, GL_INVALID_ENUM); return; }
    switch (pname) {
    case GL_AMBIENT: put4(o, m->ambient); break;
    case GL_DIFFUSE: put4(o, m->diffuse); break;
    case GL_SPECULAR: put4(o, m->specular); break;
    case GL_EMISSION: put4(o, m->emission); break;
    case GL_SHININESS: o[0] = m->shininess; break;
    default: set_error(c, GL_INVALID_ENUM); break;
    }
}

/* ================================================================ fog (gl_api.c:1434-1512) */
static bool fog_mode_ok(GLenum m) { return m == GL_LINEAR || m == GL_EXP || m == GL_EXP2; }

void glFogi(GLenum pname, GLint param)
{
    MTGL_CTX();
    if (pname != GL_FOG_MODE || !fog_mode_ok((GLenum)param)) { set_error(c, GL_INVALID_ENUM); return; }
    c->fog_mode = (GLenum)param;
    mark_state_dirty(c);
}

void glFogfv(GLenum pname, const GLfloat *p)
{
    MTGL_CTX();
    switch (pname) {
    case GL_FOG_MODE:
        if (!fog_mode_ok((GLenum)(int)p[0])) { set_error(c, GL_INVALID_ENUM); return; }
        c->fog_mode = (GLenum)(int)p[0];
        break;
    case GL_FOG_DENSITY:
        if (p[0] < 0.0f) { set_error(c, GL_INVALID_VALUE); return; }
        c->fog_density = p[0];
        break;
    case GL_FOG_START: c->fog_start = p[0]; break;
    case GL_FOG_END: c->fog_end = p[0]; break;
    case GL_FOG_COLOR: c->fog_color = rgba(p[0], p[1], p[2], p[3]); break;
    default: set_error(c, GL_INVALID_ENUM); return;
    }
    mark_state_dirty(c);
}

void glFogf(GLenum pname, GLfloat param)
{
    MTGL_CTX();
    if (pname == GL_FOG_COLOR) { set_error(c, GL_INVALID_ENUM); return; }   /* scalar entry point has no colour case */
    glFogfv(pname, &param);
}

/* ================================================================ raster / fragment state */
void glCullFace(GLenum mode)
{
    MTGL_CTX();
    if (record(c, make_cmd(OP_CULL_FACE, mode))) return;
    if (mode != GL_FRONT && mode != GL_BACK && mode != GL_FRONT_AND_BACK) { set_error(c, GL_INVALID_ENUM); return; }
    c->cull_face_mode = mode;
    mark_state_dirty(c);
}

void glFrontFace(GLenum mode)
{
    MTGL_CTX();
    if (record(c, make_cmd(OP_FRONT_FACE, mode))) return;
    if (mode != GL_CW && mode != GL_CCW) { set_error(c, GL_INVALID_ENUM); return; }
    c->front_face = mode;
    mark_state_dirty(c);
}

void glPolygonMode(GLenum face, GLenum mode)
{
    MTGL_CTX();
    if (face != GL_FRONT && face != GL_BACK && face != GL_FRONT_AND_BACK) { set_error(c, GL_INVALID_ENUM); return; }
    if (mode != GL_POINT && mode != GL_LINE && mode != GL_FILL) { set_error(c, GL_INVALID_ENUM); return; }
    if (face != GL_BACK) c->polygon_mode_front = mode;
    if (face != GL_FRONT) c->polygon_mode_back = mode;
    mark_state_dirty(c);
}

void glLineWidth(GLfloat w)
{
    MTGL_CTX();
    if (w <= 0.0f || std::isnan(w) || std::isinf(w)) { set_error(c, GL_INVALID_VALUE); return; }
    c->line_width = w;
    mark_state_dirty(c);
}

void glPointSize(GLfloat s)
{
    MTGL_CTX();
    if (s <= 0.0f || std::isnan(s) || std::isinf(s)) { set_error(c, GL_INVALID_VALUE); return; }
    c->point_size = s;
    mark_state_dirty(c);
}

void glHint(GLenum target, GLenum mode) /* gl_api.c:1156-1176 */
{
    MTGL_CTX();
    if (mode != GL_DONT_CARE && mode != GL_FASTEST && mode != GL_NICEST) { set_error(c, GL_INVALID_ENUM); return; }
    switch (target) {
    case GL_PERSPECTIVE_CORRECTION_HINT: c->perspective_hint = mode; mark_state_dirty(c); break;
    case GL_POINT_SMOOTH_HINT: case GL_LINE_SMOOTH_HINT: case GL_FOG_HINT: break;
    default: set_error(c, GL_INVALID_ENUM); break;
    }
}

void glAlphaFunc(GLenum func, GLclampf ref)
{
    MTGL_CTX();
    if (!is_compare_func(func)) { set_error(c, GL_INVALID_ENUM); return; }
    c->alpha_func = func;
    if (ref < 0.0f) ref = 0.0f;
    if (ref > 1.0f) ref = 1.0f;
    c->alpha_ref = ref;
    mark_state_dirty(c);
}

void glStencilFunc(GLenum func, GLint ref, GLuint mask)
{
    MTGL_CTX();
    if (!is_compare_func(func)) { set_error(c, GL_INVALID_ENUM); return; }
    c->stencil_func = func; c->stencil_ref = ref; c->stencil_mask = mask;
    mark_state_dirty(c);
}

static bool stencil_op_ok(GLenum op)
{
    switch (op) {
    case GL_KEEP: case GL_ZERO: case GL_REPLACE: case GL_INCR: case GL_INCR_WRAP: case GL_DECR: case GL_DECR_WRAP: case GL_INVERT:
        return true;
    default: return false;
    }
}

void glStencilOp(GLenum sfail, GLenum dpfail, GLenum dppass)
{
    MTGL_CTX();
    if (!stencil_op_ok(sfail) || !stencil_op_ok(dpfail) || !stencil_op_ok(dppass)) { set_error(c, GL_INVALID_ENUM); return; }
    c->stencil_fail = sfail; c->stencil_zfail = dpfail; c->stencil_zpass = dppass;
    mark_state_dirty(c);
}

void glStencilMask(GLuint mask) { MTGL_CTX(); c->stencil_writemask = mask; mark_state_dirty(c); }

void glDepthFunc(GLenum func)
{
    MTGL_CTX();
    if (record(c, make_cmd(OP_DEPTH_FUNC, func))) return;
    if (!is_compare_func(func)) { set_error(c, GL_INVALID_ENUM); return; }
    c->depth_func = func;
    mark_state_dirty(c);
}

void glDepthMask(GLboolean flag)
{
    MTGL_CTX();
    if (record(c, make_cmd(OP_DEPTH_MASK, flag))) return;
    c->depth_mask = flag;
    mark_state_dirty(c);
}

static bool blend_factor_ok(GLenum f, bool src) /* gl_api.c:910-934 */
{
    switch (f) {
    case GL_ZERO: case GL_ONE: case GL_SRC_COLOR: case GL_ONE_MINUS_SRC_COLOR: case GL_DST_COLOR:
    case GL_ONE_MINUS_DST_COLOR: case GL_SRC_ALPHA: case GL_ONE_MINUS_SRC_ALPHA: case GL_DST_ALPHA:
    case GL_ONE_MINUS_DST_ALPHA: case GL_CONSTANT_COLOR: case GL_ONE_MINUS_CONSTANT_COLOR:
    case GL_CONSTANT_ALPHA: case GL_ONE_MINUS_CONSTANT_ALPHA:
        return true;
    case GL_SRC_ALPHA_SATURATE: return src;
    default: return false;
    }
}

void glBlendFunc(GLenum sfactor, GLenum dfactor)
{
    MTGL_CTX();
    if (record(c, make_cmd(OP_BLEND_FUNC, sfactor, dfactor))) return;
    if (!blend_factor_ok(sfactor, true) || !blend_factor_ok(dfactor, false)) { set_error(c, GL_INVALID_ENUM); return; }
    c->blend_src = sfactor; c->blend_dst = dfactor;
    mark_state_dirty(c);
}

void glColorMask(GLboolean r, GLboolean g, GLboolean b, GLboolean a)
{
    MTGL_CTX();
    c->color_mask[0] = r; c->color_mask[1] = g; c->color_mask[2] = b; c->color_mask[3] = a;
    mark_state_dirty(c);
}

void glPixelStorei(GLenum pname, GLint param) /* gl_api.c:1141-1154 */
{
    MTGL_CTX();
    if (pname == GL_PACK_ALIGNMENT || pname == GL_UNPACK_ALIGNMENT) {
        if (param != 1 && param != 2 && param != 4 && param != 8) set_error(c, GL_INVALID_VALUE);
    } else set_error(c, GL_INVALID_ENUM);
}

GLenum glGetError(void)
{
    MTGL_CTX_RET(GL_NO_ERROR);
    GLenum e = c->error;
    c->error = GL_NO_ERROR;
    return e;
}

/* ================================================================ display lists (gl_api.c:2268-2499, lists.c) */
GLuint glGenLists(GLsizei range)
{
    MTGL_CTX_RET(0);
    if (range < 0) { set_error(c, GL_INVALID_VALUE); return 0; }
    if (range == 0) return 0;
    /* first contiguous run of free slots (find_free_range, lists.c:39-56) */
    int run = 0, start = -1;
    for (size_t i = 0; i < c->lists.size(); i++) {
        if (!c->lists[i].allocated) {
            if (run == 0) start = (int)i;
            if (++run >= range) break;
        } else run = 0;
    }
    if (run < range) {
        if (c->lists.size() + (size_t)range > (size_t)kMaxLists) return 0;
        start = (int)c->lists.size();
        c->lists.resize(c->lists.size() + (size_t)range);
    }
    for (GLsizei i = 0; i < range; i++) {
        c->lists[start + i] = DisplayList();
        c->lists[start + i].allocated = true;
    }
    return (GLuint)start + 1;
}

/* drop a list's compiled geometry (queued draws may still read the buffer: they go first) */
static void release_list_buffer(GLState *c, GLuint id, DisplayList *l)
{
    l->runs.clear();
    if (!l->has_buffer) return;
    flush_batch(c);
    mtgl_dev_buffer_data(c->dev, MTGL_LIST_BUFFER_BASE + id, 0, nullptr);
    l->has_buffer = false;
}

void glDeleteLists(GLuint list, GLsizei range)
{
    MTGL_CTX();
    if (range < 0) { set_error(c, GL_INVALID_VALUE); return; }
    for (GLsizei i = 0; i < range; i++) {
        DisplayList *l = get_list(c, list + i);
        if (l) { release_list_buffer(c, list + i, l); *l = DisplayList(); }
    }
}

void glNewList(GLuint list, GLenum mode)
{
    MTGL_CTX();
    if (c->list_index != 0) { set_error(c, GL_INVALID_OPERATION); return; }
    if (mode != GL_COMPILE && mode != GL_COMPILE_AND_EXECUTE) { set_error(c, GL_INVALID_ENUM); return; }
    DisplayList *l = get_list(c, list);
    if (!l) { set_error(c, GL_INVALID_VALUE); return; }
    release_list_buffer(c, list, l);
    l->cmds.clear();
    l->valid = false;
    c->list_index = list;
    c->list_mode = mode;
}

/* Find the glBegin ... glEnd stretches made of nothing but vertices and attribute calls and move their vertices into
 * a device buffer.  A stretch qualifies when every attribute it sets is set before its first
 * vertex -- then each vertex's colour / texture coordinate / normal is known at compile time; an attribute it never
 * sets is the current value at call time.  Everything else keeps being replayed call by call. */
static void compile_list_runs(GLState *c, GLuint id, DisplayList *l)
{
    static const bool disabled = std::getenv("MTGL_NO_LIST_RUNS") != nullptr;      /* A/B switch for measurements */
    if (disabled) return;
    std::vector<float> verts;                   /* 12 floats per vertex: position 3, colour 4, texture coordinate 2, normal 3 */
    const std::vector<ListCmd> &cm = l->cmds;
    for (size_t i = 0; i < cm.size(); i++) {
        if (cm[i].op != OP_BEGIN || cm[i].e[0] > GL_POLYGON) continue;
        ListRun r;
        r.begin_cmd = i; r.mode = cm[i].e[0]; r.first = (uint32_t)(verts.size() / 12);
        float col[4] = { 0, 0, 0, 1 }, tex[2] = { 0, 0 }, nrm[3] = { 0, 0, 1 };
        bool ok = true, seen_vertex = false, closed = false;
        size_t j = i + 1;
        const size_t rollback = verts.size();
        for (; j < cm.size(); j++) {
            const ListCmd &k = cm[j];
            if (k.op == OP_END) { closed = true; break; }
            if (k.op == OP_VERTEX) {
                seen_vertex = true;
                const float v[12] = { k.f[0], k.f[1], k.f[2], col[0], col[1], col[2], col[3], tex[0], tex[1], nrm[0], nrm[1], nrm[2] };
                verts.insert(verts.end(), v, v + 12);
            } else if (k.op == OP_COLOR) {
                if (seen_vertex && !r.has_color) ok = false;    /* the first vertices would need the caller's colour */
                r.has_color = true; std::memcpy(col, k.f, sizeof col);
            } else if (k.op == OP_TEXCOORD) {
                if (seen_vertex && !r.has_texcoord) ok = false;
                r.has_texcoord = true; std::memcpy(tex, k.f, sizeof tex);
            } else if (k.op == OP_NORMAL) {
                if (seen_vertex && !r.has_normal) ok = false;
                r.has_normal = true; std::memcpy(nrm, k.f, sizeof nrm);
            } else { ok = false; break; }                       /* a state change, a nested list ...: not pure geometry */
        }
        r.count = (uint32_t)((verts.size() - rollback) / 12);
        if (!ok || !closed || r.count < kListRunMinVertices) {
            verts.resize(rollback);
            if (closed) i = j;                                  /* resume behind the glEnd */
            continue;
        }
        r.end_cmd = j;
        std::memcpy(r.last_color, col, sizeof col); std::memcpy(r.last_texcoord, tex, sizeof tex); std::memcpy(r.last_normal, nrm, sizeof nrm);
        l->runs.push_back(r);
        i = j;
    }
    if (l->runs.empty()) return;
    if (mtgl_dev_buffer_data(c->dev, MTGL_LIST_BUFFER_BASE + id, (uint64_t)verts.size() * 4, verts.data()) == MTGL_OK) l->has_buffer = true;
    else l->runs.clear();                                       /* no room: replay as before */
}

void glEndList(void)
{
    MTGL_CTX();
    if (c->list_index == 0) { set_error(c, GL_INVALID_OPERATION); return; }
    const GLuint id = c->list_index;
    DisplayList *l = get_list(c, id);
    c->list_index = 0;
    c->list_mode = 0;
    if (l) {
        l->valid = true;
        if (id <= 1024) compile_list_runs(c, id, l);
    }
}

static void replay(GLState *c, GLuint id) /* execute_list, gl_api.c:2331-2439: replays through the public API */
{
    DisplayList *l = get_list(c, id);
    if (!l || !l->valid) return;
    size_t next_run = 0;
    for (size_t i = 0; i < l->cmds.size(); i++) {
        /* compiled geometry: one array draw instead of the calls between glBegin and glEnd */
        l = get_list(c, id);
        while (next_run < l->runs.size() && l->runs[next_run].begin_cmd < i) next_run++;
        if (next_run < l->runs.size() && l->runs[next_run].begin_cmd == i && l->has_buffer) {
            const ListRun r = l->runs[next_run++];
            if (draw_list_run(c, id, r)) { i = r.end_cmd; continue; }
        }
        /* copy: a replayed call may append to this very list (COMPILE_AND_EXECUTE recursion) */
        ListCmd cmd = get_list(c, id)->cmds[i];
        switch (cmd.op) {
        case OP_BEGIN: glBegin(cmd.e[0]); break;
        case OP_END: glEnd(); break;
        case OP_VERTEX: glVertex3f(cmd.f[0], cmd.f[1], cmd.f[2]); break;
        case OP_COLOR: glColor4f(cmd.f[0], cmd.f[1], cmd.f[2], cmd.f[3]); break;
        case OP_TEXCOORD: glTexCoord2f(cmd.f[0], cmd.f[1]); break;
        case OP_NORMAL: glNormal3f(cmd.f[0], cmd.f[1], cmd.f[2]); break;
        case OP_TRANSLATE: glTranslatef(cmd.f[0], cmd.f[1], cmd.f[2]); break;
        case OP_ROTATE: glRotatef(cmd.f[0], cmd.f[1], cmd.f[2], cmd.f[3]); break;
        case OP_SCALE: glScalef(cmd.f[0], cmd.f[1], cmd.f[2]); break;
        case OP_PUSH_MATRIX: glPushMatrix(); break;
        case OP_POP_MATRIX: glPopMatrix(); break;
        case OP_LOAD_IDENTITY: glLoadIdentity(); break;
        case OP_MULT_MATRIX: glMultMatrixf(cmd.f); break;
        case OP_LOAD_MATRIX: glLoadMatrixf(cmd.f); break;
        case OP_MATRIX_MODE: glMatrixMode(cmd.e[0]); break;
        case OP_ORTHO: glOrtho(cmd.d[0], cmd.d[1], cmd.d[2], cmd.d[3], cmd.d[4], cmd.d[5]); break;
        case OP_FRUSTUM: glFrustum(cmd.d[0], cmd.d[1], cmd.d[2], cmd.d[3], cmd.d[4], cmd.d[5]); break;
        case OP_ENABLE: glEnable(cmd.e[0]); break;
        case OP_DISABLE: glDisable(cmd.e[0]); break;
        case OP_BIND_TEXTURE: glBindTexture(cmd.e[0], cmd.e[1]); break;
        case OP_BLEND_FUNC: glBlendFunc(cmd.e[0], cmd.e[1]); break;
        case OP_DEPTH_FUNC: glDepthFunc(cmd.e[0]); break;
        case OP_DEPTH_MASK: glDepthMask((GLboolean)cmd.e[0]); break;
        case OP_CULL_FACE: glCullFace(cmd.e[0]); break;
        case OP_FRONT_FACE: glFrontFace(cmd.e[0]); break;
        case OP_SHADE_MODEL: glShadeModel(cmd.e[0]); break;
        case OP_LIGHTF: glLightf(cmd.e[0], cmd.e[1], cmd.f[0]); break;
        case OP_LIGHTFV: glLightfv(cmd.e[0], cmd.e[1], cmd.f); break;
        case OP_MATERIALF: glMaterialf(cmd.e[0], cmd.e[1], cmd.f[0]); break;
        case OP_MATERIALFV: glMaterialfv(cmd.e[0], cmd.e[1], cmd.f); break;
        case OP_CALL_LIST: glCallList(cmd.e[0]); break;
        }
    }
}

void glCallList(GLuint list)
{
    MTGL_CTX();
    if (record(c, make_cmd(OP_CALL_LIST, list))) return;
    if (c->list_call_depth >= (GLuint)kMaxListDepth) { set_error(c, GL_STACK_OVERFLOW); return; }
    c->list_call_depth++;
    replay(c, list);
    c->list_call_depth--;
}

void glCallLists(GLsizei n, GLenum type, const GLvoid *lists)
{
    MTGL_CTX();
    if (n < 0) { set_error(c, GL_INVALID_VALUE); return; }
    for (GLsizei i = 0; i < n; i++) {
        GLuint off;
        switch (type) {
        case GL_UNSIGNED_BYTE: off = ((const GLubyte *)lists)[i]; break;
        case GL_UNSIGNED_SHORT: off = ((const GLushort *)lists)[i]; break;
        case GL_UNSIGNED_INT: off = ((const GLuint *)lists)[i]; break;
        default: set_error(c, GL_INVALID_ENUM); return;
        }
        glCallList(c->list_base + off);
    }
}

void glListBase(GLuint base) { MTGL_CTX(); c->list_base = base; }

GLboolean glIsList(GLuint list)
{
    MTGL_CTX_RET(GL_FALSE);
    DisplayList *l = get_list(c, list);
    return (l && l->valid) ? GL_TRUE : GL_FALSE;
}

/* ================================================================ queries (gl_api.c:2505-3189) */
static bool get_integer(GLState *c, GLenum pname, GLint *o)
{
    switch (pname) {
    case GL_VIEWPORT: o[0] = c->viewport_x; o[1] = c->viewport_y; o[2] = c->viewport_w; o[3] = c->viewport_h; break;
    case GL_MATRIX_MODE: o[0] = c->matrix_mode; break;
    case GL_MODELVIEW_STACK_DEPTH: o[0] = c->modelview_depth + 1; break;
    case GL_PROJECTION_STACK_DEPTH: o[0] = c->projection_depth + 1; break;
    case GL_TEXTURE_STACK_DEPTH: o[0] = c->texture_depth + 1; break;
    case GL_SHADE_MODEL: o[0] = c->shade_model; break;
    case GL_COLOR_MATERIAL_FACE: o[0] = c->color_material_face; break;
    case GL_COLOR_MATERIAL_PARAMETER: o[0] = c->color_material_mode; break;
    case GL_FOG_MODE: o[0] = c->fog_mode; break;
    case GL_LIGHT_MODEL_LOCAL_VIEWER: o[0] = c->light_model_local_viewer; break;
    case GL_LIGHT_MODEL_TWO_SIDE: o[0] = c->light_model_two_side; break;
    case GL_CULL_FACE_MODE: o[0] = c->cull_face_mode; break;
    case GL_FRONT_FACE: o[0] = c->front_face; break;
    case GL_POLYGON_MODE: o[0] = c->polygon_mode_front; o[1] = c->polygon_mode_back; break;
    case GL_TEXTURE_BINDING_2D: o[0] = c->bound_texture_2d; break;
    case GL_TEXTURE_ENV_MODE: o[0] = c->tex_env_mode; break;
    case GL_SCISSOR_BOX: o[0] = c->scissor_x; o[1] = c->scissor_y; o[2] = c->scissor_w; o[3] = c->scissor_h; break;
    case GL_ALPHA_TEST_FUNC: o[0] = c->alpha_func; break;
    case GL_STENCIL_FUNC: o[0] = c->stencil_func; break;
    case GL_STENCIL_VALUE_MASK: o[0] = c->stencil_mask; break;
    case GL_STENCIL_REF: o[0] = c->stencil_ref; break;
    case GL_STENCIL_FAIL: o[0] = c->stencil_fail; break;
    case GL_STENCIL_PASS_DEPTH_FAIL: o[0] = c->stencil_zfail; break;
    case GL_STENCIL_PASS_DEPTH_PASS: o[0] = c->stencil_zpass; break;
    case GL_STENCIL_WRITEMASK: o[0] = c->stencil_writemask; break;
    case GL_STENCIL_CLEAR_VALUE: o[0] = c->stencil_clear; break;
    case GL_DEPTH_FUNC: o[0] = c->depth_func; break;
    case GL_DEPTH_WRITEMASK: o[0] = c->depth_mask; break;
    case GL_BLEND_SRC: o[0] = c->blend_src; break;
    case GL_BLEND_DST: o[0] = c->blend_dst; break;
    case GL_COLOR_WRITEMASK: for (int i = 0; i < 4; i++) o[i] = c->color_mask[i]; break;
    case GL_UNPACK_ALIGNMENT: case GL_PACK_ALIGNMENT: o[0] = 4; break;
    case GL_PERSPECTIVE_CORRECTION_HINT: o[0] = c->perspective_hint; break;
    case GL_FOG_HINT: case GL_LINE_SMOOTH_HINT: case GL_POINT_SMOOTH_HINT: case GL_POLYGON_SMOOTH_HINT: o[0] = GL_DONT_CARE; break;
    case GL_ARRAY_BUFFER_BINDING: o[0] = c->bound_array_buffer; break;
    case GL_ELEMENT_ARRAY_BUFFER_BINDING: o[0] = c->bound_element_buffer; break;
    case GL_VERTEX_ARRAY_SIZE: o[0] = c->vertex_pointer.size; break;
    case GL_VERTEX_ARRAY_TYPE: o[0] = c->vertex_pointer.type; break;
    case GL_VERTEX_ARRAY_STRIDE: o[0] = c->vertex_pointer.stride; break;
    case GL_COLOR_ARRAY_SIZE: o[0] = c->color_pointer.size; break;
    case GL_COLOR_ARRAY_TYPE: o[0] = c->color_pointer.type; break;
    case GL_COLOR_ARRAY_STRIDE: o[0] = c->color_pointer.stride; break;
    case GL_NORMAL_ARRAY_TYPE: o[0] = c->normal_pointer.type; break;
    case GL_NORMAL_ARRAY_STRIDE: o[0] = c->normal_pointer.stride; break;
    case GL_TEXTURE_COORD_ARRAY_SIZE: o[0] = c->texcoord_pointer.size; break;
    case GL_TEXTURE_COORD_ARRAY_TYPE: o[0] = c->texcoord_pointer.type; break;
    case GL_TEXTURE_COORD_ARRAY_STRIDE: o[0] = c->texcoord_pointer.stride; break;
    case GL_LIST_BASE: o[0] = c->list_base; break;
    case GL_LIST_INDEX: o[0] = c->list_index; break;
    case GL_LIST_MODE: o[0] = c->list_index ? c->list_mode : 0; break;
    case GL_CURRENT_RASTER_POSITION: o[0] = c->raster_pos_x; o[1] = c->raster_pos_y; o[2] = 0; o[3] = 1; break;
    case GL_CURRENT_RASTER_POSITION_VALID: o[0] = c->raster_pos_valid; break;
    case GL_RENDER_MODE: o[0] = 0x1C00; break;
    case GL_MAX_LIGHTS: o[0] = kMaxLights; break;
    case GL_MAX_CLIP_PLANES: o[0] = 6; break;
    case GL_MAX_TEXTURE_SIZE: o[0] = kMaxTextureSize; break;
    case GL_MAX_3D_TEXTURE_SIZE: case GL_MAX_CUBE_MAP_TEXTURE_SIZE: o[0] = 0; break;
    case GL_MAX_PIXEL_MAP_TABLE: o[0] = 256; break;
    case GL_MAX_ATTRIB_STACK_DEPTH: case GL_MAX_CLIENT_ATTRIB_STACK_DEPTH: o[0] = 16; break;
    case GL_MAX_MODELVIEW_STACK_DEPTH: case GL_MAX_PROJECTION_STACK_DEPTH: case GL_MAX_TEXTURE_STACK_DEPTH:
        o[0] = kMatrixStackDepth; break;
    case GL_MAX_NAME_STACK_DEPTH: o[0] = 64; break;
    case GL_MAX_VIEWPORT_DIMS: o[0] = 16384; o[1] = 16384; break;
    case GL_MAX_TEXTURE_UNITS: o[0] = 1; break;
    case GL_MAX_ELEMENTS_VERTICES: case GL_MAX_ELEMENTS_INDICES: o[0] = 65536; break;
    case GL_SUBPIXEL_BITS: o[0] = 4; break;
    case GL_INDEX_BITS: o[0] = 0; break;
    case GL_RED_BITS: case GL_GREEN_BITS: case GL_BLUE_BITS: case GL_ALPHA_BITS: o[0] = 8; break;
    case GL_DEPTH_BITS: o[0] = 32; break;
    case GL_STENCIL_BITS: o[0] = 8; break;
    case GL_ACCUM_RED_BITS: case GL_ACCUM_GREEN_BITS: case GL_ACCUM_BLUE_BITS: case GL_ACCUM_ALPHA_BITS: o[0] = 0; break;
    case GL_AUX_BUFFERS: o[0] = 0; break;
    case GL_DOUBLEBUFFER: o[0] = GL_TRUE; break;
    case GL_STEREO: o[0] = GL_FALSE; break;
    case GL_RGBA_MODE: o[0] = GL_TRUE; break;
    case GL_INDEX_MODE: o[0] = GL_FALSE; break;
    case GL_SAMPLE_BUFFERS: case GL_SAMPLES: o[0] = 0; break;
    default: return false;
    }
    return true;
}

void glGetIntegerv(GLenum pname, GLint *params)
{
    MTGL_CTX();
    if (!params) return;
    if (!get_integer(c, pname, params)) set_error(c, GL_INVALID_ENUM);
}

void glGetFloatv(GLenum pname, GLfloat *o)
{
    MTGL_CTX();
    if (!o) return;
    switch (pname) {
    case GL_MODELVIEW_MATRIX: std::memcpy(o, c->modelview[c->modelview_depth], 64); break;
    case GL_PROJECTION_MATRIX: std::memcpy(o, c->projection[c->projection_depth], 64); break;
    case GL_TEXTURE_MATRIX: std::memcpy(o, c->texture[c->texture_depth], 64); break;
    case GL_CURRENT_COLOR: case GL_CURRENT_RASTER_COLOR: put4(o, c->current_color); break;
    case GL_CURRENT_NORMAL: std::memcpy(o, c->current_normal, 12); break;
    case GL_CURRENT_TEXTURE_COORDS: o[0] = c->current_texcoord[0]; o[1] = c->current_texcoord[1]; break;
    case GL_CURRENT_RASTER_POSITION: o[0] = (GLfloat)c->raster_pos_x; o[1] = (GLfloat)c->raster_pos_y; o[2] = 0.0f; o[3] = 1.0f; break;
    case GL_DEPTH_RANGE: o[0] = (GLfloat)c->depth_near; o[1] = (GLfloat)c->depth_far; break;
    case GL_VIEWPORT: o[0] = (GLfloat)c->viewport_x; o[1] = (GLfloat)c->viewport_y; o[2] = (GLfloat)c->viewport_w; o[3] = (GLfloat)c->viewport_h; break;
    case GL_DEPTH_CLEAR_VALUE: o[0] = (GLfloat)c->clear_depth; break;
    case GL_COLOR_CLEAR_VALUE: put4(o, c->clear_color); break;
    case GL_FOG_COLOR: put4(o, c->fog_color); break;
    case GL_FOG_DENSITY: o[0] = c->fog_density; break;
    case GL_FOG_START: o[0] = c->fog_start; break;
    case GL_FOG_END: o[0] = c->fog_end; break;
    case GL_LIGHT_MODEL_AMBIENT: put4(o, c->light_model_ambient); break;
    case GL_ALPHA_TEST_REF: o[0] = c->alpha_ref; break;
    case GL_BLEND_COLOR: o[0] = o[1] = o[2] = o[3] = 0.0f; break;
    case GL_POINT_SIZE: o[0] = c->point_size; break;
    case GL_POINT_SIZE_RANGE: o[0] = 1.0f; o[1] = 64.0f; break;
    case GL_POINT_SIZE_GRANULARITY: o[0] = 1.0f; break;
    case GL_LINE_WIDTH: o[0] = c->line_width; break;
    case GL_LINE_WIDTH_RANGE: o[0] = 1.0f; o[1] = 16.0f; break;
    case GL_LINE_WIDTH_GRANULARITY: o[0] = 1.0f; break;
    case GL_POLYGON_OFFSET_FACTOR: case GL_POLYGON_OFFSET_UNITS: o[0] = 0.0f; break;
    case GL_TEXTURE_ENV_COLOR: put4(o, c->tex_env_color); break;
    case GL_SCISSOR_BOX: o[0] = (GLfloat)c->scissor_x; o[1] = (GLfloat)c->scissor_y; o[2] = (GLfloat)c->scissor_w; o[3] = (GLfloat)c->scissor_h; break;
    case GL_MAX_TEXTURE_LOD_BIAS: o[0] = 2.0f; break;
    default: {
        GLint iv[4] = { 0, 0, 0, 0 };                    /* integer query, first component only (gl_api.c:2989-3003) */
        if (get_integer(c, pname, iv)) o[0] = (GLfloat)iv[0];
        else set_error(c, GL_INVALID_ENUM);
        break;
    }
    }
}

void glGetDoublev(GLenum pname, GLdouble *o) /* gl_api.c:3007-3053 */
{
    MTGL_CTX();
    if (!o) return;
    GLfloat f[16];
    glGetFloatv(pname, f);
    int n = 1;
    switch (pname) {
    case GL_MODELVIEW_MATRIX: case GL_PROJECTION_MATRIX: case GL_TEXTURE_MATRIX: n = 16; break;
    case GL_CURRENT_COLOR: case GL_CURRENT_RASTER_COLOR: case GL_CURRENT_RASTER_POSITION: case GL_VIEWPORT:
    case GL_SCISSOR_BOX: case GL_COLOR_CLEAR_VALUE: case GL_FOG_COLOR: case GL_LIGHT_MODEL_AMBIENT:
    case GL_BLEND_COLOR: case GL_TEXTURE_ENV_COLOR: case GL_COLOR_WRITEMASK: n = 4; break;
    case GL_CURRENT_NORMAL: n = 3; break;
    case GL_CURRENT_TEXTURE_COORDS: case GL_DEPTH_RANGE: case GL_POINT_SIZE_RANGE: case GL_LINE_WIDTH_RANGE:
    case GL_MAX_VIEWPORT_DIMS: case GL_POLYGON_MODE: n = 2; break;
    default: break;
    }
    for (int i = 0; i < n; i++) o[i] = (GLdouble)f[i];
}

void glGetBooleanv(GLenum pname, GLboolean *o) /* gl_api.c:3055-3149 */
{
    MTGL_CTX();
    if (!o) return;
    uint32_t bit = cap_bit(pname);
    if (bit) { o[0] = (c->caps & bit) ? GL_TRUE : GL_FALSE; return; }
    switch (pname) {
    case GL_DEPTH_WRITEMASK: o[0] = c->depth_mask; return;
    case GL_COLOR_WRITEMASK: for (int i = 0; i < 4; i++) o[i] = c->color_mask[i]; return;
    case GL_DOUBLEBUFFER: case GL_RGBA_MODE: o[0] = GL_TRUE; return;
    case GL_STEREO: case GL_INDEX_MODE: o[0] = GL_FALSE; return;
    case GL_CURRENT_RASTER_POSITION_VALID: o[0] = c->raster_pos_valid; return;
    case GL_LIGHT_MODEL_LOCAL_VIEWER: o[0] = c->light_model_local_viewer; return;
    case GL_LIGHT_MODEL_TWO_SIDE: o[0] = c->light_model_two_side; return;
    case GL_VERTEX_ARRAY: case GL_COLOR_ARRAY: case GL_NORMAL_ARRAY: case GL_TEXTURE_COORD_ARRAY:
        o[0] = (c->client_state & client_bit(pname)) ? GL_TRUE : GL_FALSE; return;
    default: break;
    }
    if (pname >= GL_LIGHT0 && pname <= GL_LIGHT7) { o[0] = c->lights[pname - GL_LIGHT0].enabled; return; }
    GLint iv[16] = { 0 };
    glGetIntegerv(pname, iv);
    o[0] = (iv[0] != 0) ? GL_TRUE : GL_FALSE;
}

const GLubyte *glGetString(GLenum name) /* gl_api.c:3176-3189 */
{
    MTGL_CTX_RET(nullptr);
    switch (name) {
    case GL_VENDOR: return (const GLubyte *)"zbufferoverflow";
    case GL_RENDERER: return (const GLubyte *)"MyTinyGL B200 (sm_100a) Renderer";
    case GL_VERSION: return (const GLubyte *)"1.5 MyTinyGL-B200";
    case GL_EXTENSIONS: return (const GLubyte *)"";
    default: set_error(c, GL_INVALID_ENUM); return nullptr;
    }
}

} // extern "C"
