/* Minimal stand-in for SDL2's <SDL.h>, used ONLY to compile the unmodified
 * CobbleTrace reference sources headless (oracle/_ref, test infrastructure).
 * SDL2 is not installed in this image.  The reference uses SDL for threads,
 * logging, memcpy, delay and (in draw2d.h Blit, never called headless) for
 * presenting the framebuffer.  Threads map to pthreads; display calls are
 * declared and defined as no-ops.
 */
#ifndef CT_SDL_STUB_H
#define CT_SDL_STUB_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>
#include <unistd.h>
#include <time.h>
#include <pthread.h>

typedef uint32_t Uint32;
typedef uint8_t Uint8;

struct SDL_Renderer;
struct SDL_Texture;
struct SDL_Window;
typedef struct SDL_Renderer SDL_Renderer;
typedef struct SDL_Texture SDL_Texture;
typedef struct SDL_Window SDL_Window;

typedef struct SDL_Surface { int w, h; void *pixels; } SDL_Surface;
typedef struct SDL_Rect { int x, y, w, h; } SDL_Rect;

#define SDL_LIL_ENDIAN 1234
#define SDL_BIG_ENDIAN 4321
#define SDL_BYTEORDER SDL_LIL_ENDIAN

#define SDL_memcpy memcpy

extern int ct_sdl_stub_quiet; /* set to 1 to silence SDL_Log */

static inline void SDL_Log(const char *fmt, ...) {
    if (ct_sdl_stub_quiet) return;
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
}

static inline void SDL_Delay(Uint32 ms) { usleep((useconds_t)ms * 1000u); }

static inline Uint32 SDL_GetTicks(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (Uint32)(ts.tv_sec * 1000u + ts.tv_nsec / 1000000u);
}

typedef struct SDL_Thread { pthread_t handle; int (*fn)(void *); void *data; } SDL_Thread;
typedef int (*SDL_ThreadFunction)(void *);

static inline void *ct_sdl_thread_tramp(void *p) {
    SDL_Thread *t = (SDL_Thread *)p;
    t->fn(t->data);
    return NULL;
}

static inline SDL_Thread *SDL_CreateThread(SDL_ThreadFunction fn, const char *name, void *data) {
    (void)name;
    SDL_Thread *t = (SDL_Thread *)calloc(1, sizeof(SDL_Thread));
    if (!t) return NULL;
    t->fn = fn;
    t->data = data;
    if (pthread_create(&t->handle, NULL, ct_sdl_thread_tramp, t) != 0) { free(t); return NULL; }
    pthread_detach(t->handle);
    return t;
}

/* Display path (draw2d.h Blit): declared so the inline function compiles; never called. */
static inline SDL_Surface *SDL_CreateRGBSurfaceFrom(void *, int, int, int, int, Uint32, Uint32, Uint32, Uint32) { return NULL; }
static inline const char *SDL_GetError(void) { return "SDL stub"; }
static inline SDL_Texture *SDL_CreateTextureFromSurface(SDL_Renderer *, SDL_Surface *) { return NULL; }
static inline int SDL_RenderCopy(SDL_Renderer *, SDL_Texture *, const SDL_Rect *, const SDL_Rect *) { return 0; }
static inline void SDL_RenderPresent(SDL_Renderer *) {}
static inline void SDL_FreeSurface(SDL_Surface *) {}
static inline void SDL_DestroyTexture(SDL_Texture *) {}

#endif
