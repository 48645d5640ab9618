/*
 * ref_shim.c -- TEST INFRASTRUCTURE ONLY.  Headless harness around the UNMODIFIED reference
 * (compiled from /root/reference/src/*.c by the top-level Makefile into oracle/_ref/).  It is the
 * only file that looks inside the reference's GLState: everything else drives the reference
 * through its public gl* API.  Recipe: SURVEY.md Appendix D.
 */
#include "src/mytinygl.h"

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

/* planes laid out as in framebuffer.h:19-25: row 0 = top, pitch = width */
int mtgl_harness_read(void *h, uint32_t *color, float *depth, uint8_t *stencil)
{
    GLState *c = (GLState *)h;
    if (!c) return -1;
    size_t n = (size_t)c->framebuffer.width * (size_t)c->framebuffer.height;
    if (color) memcpy(color, c->framebuffer.color, n * sizeof(uint32_t));
    if (depth) memcpy(depth, c->framebuffer.depth, n * sizeof(float));
    if (stencil) memcpy(stencil, c->framebuffer.stencil, n);
    return 0;
}

const char *mtgl_harness_kind(void) { return "reference"; }
