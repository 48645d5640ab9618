/* front_internal.h -- declarations shared by the front-end translation units. */
#ifndef MTGL_FRONT_INTERNAL_H
#define MTGL_FRONT_INTERNAL_H

#include "gl_state.h"

namespace mtgl {

extern thread_local GLState *g_ctx;     /* current context (gl_api.c:14-23 keeps it thread local too) */

/* sticky first error (gl_api.c:30-35) */
void set_error(GLState *c, GLenum e);

/* batching */
void mark_state_dirty(GLState *c);      /* any state the device reads has changed */
void mark_matrix_dirty(GLState *c);     /* only one of the three current matrices has changed */
void flush_batch(GLState *c);                 /* submit queued clear + draws to the device */
void sync_device(GLState *c);                 /* flush + wait */
void emit_vertex(GLState *c, float x, float y, float z);
void end_primitive(GLState *c);

/* display-list recording: returns true when the call must not execute (GL_COMPILE) */
bool record(GLState *c, const ListCmd &cmd);
inline bool compiling(const GLState *c) { return c->list_index != 0; }

/* IEEE single-precision helpers restating src/graphics.h (column-major 4x4) */
void mat_identity(float *m);
void mat_mul(const float *a, const float *b, float *out);      /* graphics.h:140-152 */
void mat_vec(const float *m, const float *v, float *out);      /* graphics.h:131-138 */
float *current_matrix(GLState *c);
GLint *current_depth(GLState *c);

Texture *get_texture(GLState *c, GLuint id);
Buffer *get_buffer(GLState *c, GLuint id);
DisplayList *get_list(GLState *c, GLuint id);
/* compiled display-list geometry (gl_front.cpp): queue run r of list 'id' as one array draw; false = replay it instead */
bool draw_list_run(GLState *c, GLuint id, const ListRun &r);
/* host view of a buffer object: materialises the mirror of an HBM-only buffer on demand */
const uint8_t *buffer_host_data(GLState *c, GLuint id);
/* copy n bytes at 'offset' out of a buffer object (small device read-back when there is no host mirror) */
bool buffer_read(GLState *c, GLuint id, uint64_t offset, uint64_t n, void *out);

uint32_t pack_rgba(Rgba c);                                   /* graphics.h:337-348 */
inline Rgba rgba(float r, float g, float b, float a) { Rgba c = { r, g, b, a }; return c; }

} // namespace mtgl

#define MTGL_CTX()            GLState *c = mtgl::g_ctx; if (!c) return
#define MTGL_CTX_RET(v)       GLState *c = mtgl::g_ctx; if (!c) return (v)

#endif
