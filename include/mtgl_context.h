/*
 * mtgl_context.h -- context management of the B200-native MyTinyGL.
 *
 * Mirrors the context API of the reference (src/mytinygl.h:228-233: gl_create_context,
 * gl_destroy_context, gl_make_current, gl_get_current_context).  GLState is opaque here; the one
 * thing applications reach into it for in the reference -- ctx->framebuffer.color, read by
 * mtgl_swap (include/mytinygl/sdl.h:76-81) -- is exposed through mtgl_map_framebuffer(), which
 * first makes the host mirror of the device planes current.
 */
#ifndef MTGL_CONTEXT_H
#define MTGL_CONTEXT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct GLState GLState;

/* NULL on failure (bad size, no CUDA device, out of memory) exactly like gl_api.c:81-90.
 * There is no CPU fallback: without a usable sm_100a device the call fails. */
GLState *gl_create_context(int32_t width, int32_t height);
void gl_destroy_context(GLState *ctx);
void gl_make_current(GLState *ctx);
GLState *gl_get_current_context(void);

/* framebuffer_t (src/framebuffer.h:19-25): row 0 = top, pitch = width */
typedef struct mtgl_framebuffer {
    int32_t width;
    int32_t height;
    uint32_t *color;    /* a<<24 | b<<16 | g<<8 | r */
    float *depth;
    uint8_t *stencil;
} mtgl_framebuffer;

#define MTGL_PLANE_COLOR   1u
#define MTGL_PLANE_DEPTH   2u
#define MTGL_PLANE_STENCIL 4u

/* Flush queued work, wait for the GPU and copy the selected planes into the context's host
 * mirror; returns the mirror (valid until the next gl* call on the context) or NULL on error. */
const mtgl_framebuffer *mtgl_map_framebuffer(GLState *ctx, unsigned planes);

/* Pipelined transfers for render loops that stream geometry in and frames out (extensions; include/mtgl_dev.h explains
 * the mechanics).  mtglBufferDataPinned is glBufferData (src/gl_api.c -> src/vbo.c:120-145) for a source in page-locked
 * host memory that stays unchanged until glFinish(): the call returns once the copy is queued.  mtglReadColorAsync hands
 * the queued work to the device and queues a copy of colour rows [y0, y1) (row 0 = top, as ctx->framebuffer.color) into
 * page-locked host memory; glFinish() waits for it. */
void mtglBufferDataPinned(unsigned target, long size, const void *pinned_data, unsigned usage);
void mtglReadColorAsync(int32_t y0, int32_t y1, uint32_t *pinned_color);

/* Back-end handle of a context (struct mtgl_dev*, include/mtgl_dev.h) for tools that need device
 * plane pointers, band ownership or timing counters. */
struct mtgl_dev *mtgl_context_device(GLState *ctx);

/* Device address and size of a buffer object's storage, for applications that fill it behind the API (each rank of a
 * multi-GPU application uploads a slice, an NCCL all-gather over NVLink completes the buffer).  The front end forgets what
 * it knew about the contents; glBufferData makes them known again.  Returns 0 on success. */
int mtgl_context_buffer_pointer(GLState *ctx, unsigned id, void **ptr, uint64_t *size);
/* The same for a buffer whose previous contents queued frames may still be reading: the name gets FRESH storage of its
 * current size (contents undefined; mtgl_dev_buffer_orphan) and *ptr its address.  Fill it on your own stream, then make
 * the context's stream (mtgl_dev_stream(mtgl_context_device(ctx))) wait for that work before the draws that read it.
 * `contents`: NULL, or host memory holding the bytes the buffer WILL contain once the fill has landed -- the front end
 * then takes the few bytes it needs for host-visible side effects (the last element of an array draw becomes the current
 * normal / colour / texture coordinate, gl_api.c:1826-1842) from there, as glBufferData does, instead of reading them
 * back from the device at every draw (which waits for everything queued). */
int mtgl_context_buffer_orphan(GLState *ctx, unsigned id, const void *contents, void **ptr, uint64_t *size);

/* Display-list geometry drawn as compiled array draws so far (glBegin ... glEnd stretches of a list that were queued as
 * one draw record instead of being replayed call by call) -- for tools and tests. */
uint64_t mtgl_context_list_runs_drawn(GLState *ctx);

/* Select the CUDA ordinal used by contexts created afterwards on this thread (-1 = current). */
void mtgl_set_device(int ordinal);

#ifdef __cplusplus
}
#endif

#endif /* MTGL_CONTEXT_H */
