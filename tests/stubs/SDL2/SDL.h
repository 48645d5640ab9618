/* Minimal stand-in for <SDL2/SDL.h> (SDL2 is not installed in the build image): just the declarations
 * include/mytinygl/sdl.h uses, so that the header can be compiled and its calls type-checked by tests/test_sdl_header.py. */
#ifndef STUB_SDL_H
#define STUB_SDL_H
#include <stdint.h>
typedef struct SDL_Window SDL_Window;
typedef struct SDL_Renderer SDL_Renderer;
typedef struct SDL_Texture SDL_Texture;
typedef struct SDL_Rect SDL_Rect;
#define SDL_INIT_VIDEO 0x20u
#define SDL_WINDOWPOS_CENTERED 0x2FFF0000
#define SDL_WINDOW_SHOWN 0x4u
#define SDL_RENDERER_ACCELERATED 0x2u
#define SDL_RENDERER_PRESENTVSYNC 0x4u
#define SDL_PIXELFORMAT_ABGR8888 0x16762004u
#define SDL_TEXTUREACCESS_STREAMING 1
int SDL_Init(uint32_t flags);
void SDL_Quit(void);
SDL_Window *SDL_CreateWindow(const char *title, int x, int y, int w, int h, uint32_t flags);
void SDL_DestroyWindow(SDL_Window *w);
SDL_Renderer *SDL_CreateRenderer(SDL_Window *w, int index, uint32_t flags);
void SDL_DestroyRenderer(SDL_Renderer *r);
SDL_Texture *SDL_CreateTexture(SDL_Renderer *r, uint32_t format, int access, int w, int h);
void SDL_DestroyTexture(SDL_Texture *t);
int SDL_UpdateTexture(SDL_Texture *t, const SDL_Rect *rect, const void *pixels, int pitch);
int SDL_RenderCopy(SDL_Renderer *r, SDL_Texture *t, const SDL_Rect *src, const SDL_Rect *dst);
void SDL_RenderPresent(SDL_Renderer *r);
#endif
