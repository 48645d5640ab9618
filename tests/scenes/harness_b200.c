/*
 * harness_b200.c -- headless harness around the B200-native library (and around the
 * front end + CPU oracle test build).  Same four entry points as oracle/ref_shim.c so that the
 * Python side drives both libraries identically.
 */
#include <GL/gl.h>
#include <mtgl_context.h>

#include <string.h>

void *mtgl_harness_create(int w, int h)
{
    GLState *c = gl_create_context(w, h);
    if (c) gl_make_current(c);
    return c;
}

void mtgl_harness_destroy(void *h)
{
    gl_make_current(NULL);
    gl_destroy_context((GLState *)h);
}

void mtgl_harness_make_current(void *h) { gl_make_current((GLState *)h); }

int mtgl_harness_read(void *h, uint32_t *color, float *depth, uint8_t *stencil)
{
    GLState *c = (GLState *)h;
    if (!c) return -1;
    unsigned planes = (color ? MTGL_PLANE_COLOR : 0) | (depth ? MTGL_PLANE_DEPTH : 0) | (stencil ? MTGL_PLANE_STENCIL : 0);
    const mtgl_framebuffer *fb = mtgl_map_framebuffer(c, planes);
    if (!fb) return -1;
    size_t n = (size_t)fb->width * (size_t)fb->height;
    if (color) memcpy(color, fb->color, n * sizeof(uint32_t));
    if (depth) memcpy(depth, fb->depth, n * sizeof(float));
    if (stencil) memcpy(stencil, fb->stencil, n);
    return 0;
}

void *mtgl_harness_device(void *h) { return mtgl_context_device((GLState *)h); }

#ifdef MTGL_HARNESS_ORACLE
const char *mtgl_harness_kind(void) { return "front+oracle"; }
#else
const char *mtgl_harness_kind(void) { return "b200"; }
#endif
