/*
 * mytinygl/sdl.h -- SDL2 glue for the B200-native MyTinyGL: the drop-in for the reference's header of the same name
 * (include/mytinygl/sdl.h: mtgl_init, mtgl_swap, mtgl_destroy and the mtgl_window / mtgl_renderer / mtgl_texture /
 * mtgl_ctx globals the testbed programs use).
 *
 * One thing differs underneath.  The reference's mtgl_swap (include/mytinygl/sdl.h:74-84) hands SDL a pointer into
 * ctx->framebuffer.color, which is host memory the rasteriser has just written -- no GL call is involved.  Here the
 * colour plane lives in HBM and GLState is opaque, so mtgl_swap first asks for the plane with mtgl_map_framebuffer()
 * (include/mtgl_context.h): that call hands the queued frame to the GPU, waits for it and copies the plane into the
 * context's host mirror -- same layout (ABGR8888 words, row 0 = top, pitch = 4 * width), so the SDL texture format and
 * everything after it stay as they were.
 */
#ifndef MYTINYGL_SDL_H
#define MYTINYGL_SDL_H

#include <SDL2/SDL.h>
#include <stdint.h>

#include "../GL/gl.h"
#include "../mtgl_context.h"

static SDL_Window *mtgl_window = NULL;
static SDL_Renderer *mtgl_renderer = NULL;
static SDL_Texture *mtgl_texture = NULL;
static GLState *mtgl_ctx = NULL;

/* release whatever mtgl_init has created so far, newest first */
static inline void mtgl_sdl_teardown_(void)
{
    if (mtgl_ctx) { gl_destroy_context(mtgl_ctx); mtgl_ctx = NULL; }
    if (mtgl_texture) { SDL_DestroyTexture(mtgl_texture); mtgl_texture = NULL; }
    if (mtgl_renderer) { SDL_DestroyRenderer(mtgl_renderer); mtgl_renderer = NULL; }
    if (mtgl_window) { SDL_DestroyWindow(mtgl_window); mtgl_window = NULL; }
    SDL_Quit();
}

/* window + accelerated vsync'ed renderer + streaming ABGR8888 texture + a current GL context of the same size;
 * 0 on success, -1 on any failure (nothing is left allocated) -- as the reference's mtgl_init (sdl.h:21-72).
 * gl_create_context fails, and with it mtgl_init, when there is no sm_100-class GPU: there is no CPU fallback. */
static inline int mtgl_init(const char *title, int32_t width, int32_t height)
{
    if (SDL_Init(SDL_INIT_VIDEO) < 0) return -1;
    mtgl_window = SDL_CreateWindow(title, SDL_WINDOWPOS_CENTERED, SDL_WINDOWPOS_CENTERED, width, height, SDL_WINDOW_SHOWN);
    if (mtgl_window) mtgl_renderer = SDL_CreateRenderer(mtgl_window, -1, SDL_RENDERER_ACCELERATED | SDL_RENDERER_PRESENTVSYNC);
    if (mtgl_renderer) mtgl_texture = SDL_CreateTexture(mtgl_renderer, SDL_PIXELFORMAT_ABGR8888, SDL_TEXTUREACCESS_STREAMING, width, height);
    if (mtgl_texture) mtgl_ctx = gl_create_context(width, height);
    if (!mtgl_ctx) {
        mtgl_sdl_teardown_();
        return -1;
    }
    gl_make_current(mtgl_ctx);
    return 0;
}

/* present the frame: the observable point of a render loop, so this is where the back end synchronises */
static inline void mtgl_swap(void)
{
    const mtgl_framebuffer *fb = mtgl_map_framebuffer(mtgl_ctx, MTGL_PLANE_COLOR);
    if (!fb) return;                /* a device error has been recorded as the sticky GL error */
    SDL_UpdateTexture(mtgl_texture, NULL, fb->color, fb->width * (int)sizeof(uint32_t));
    SDL_RenderCopy(mtgl_renderer, mtgl_texture, NULL, NULL);
    SDL_RenderPresent(mtgl_renderer);
}

static inline void mtgl_destroy(void)
{
    mtgl_sdl_teardown_();
}

#endif /* MYTINYGL_SDL_H */
