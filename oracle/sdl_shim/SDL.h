/*
 * SDL.h -- TEST INFRASTRUCTURE ONLY: a minimal stand-in for the few SDL2 names that the reference's
 * src/examples/Texture.h touches, so that the UNMODIFIED Texture.h compiles in place into
 * oracle/_ref/libswr_ref.so (SDL2 is not installed here).  Surfaces are plain 32-bit 0x00RRGGBB
 * arrays, exactly the layout Texture.h assumes (Texture.h:183-189, 228-232); the blit of an
 * identically formatted surface is a copy.
 */
#ifndef ORACLE_SDL_SHIM_H
#define ORACLE_SDL_SHIM_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>

typedef uint8_t Uint8;
typedef uint32_t Uint32;

struct SDL_Rect { int x, y, w, h; };

struct SDL_Surface {
    int w, h;
    int pitch;
    void *pixels;
};

inline SDL_Surface *SDL_CreateRGBSurface(Uint32, int width, int height, int depth, Uint32, Uint32, Uint32, Uint32)
{
    if (depth != 32 || width <= 0 || height <= 0) return NULL;
    SDL_Surface *s = (SDL_Surface *)malloc(sizeof(SDL_Surface));
    s->w = width;
    s->h = height;
    s->pitch = width * 4;
    s->pixels = calloc((size_t)width * height, 4);
    return s;
}

inline void SDL_FreeSurface(SDL_Surface *s)
{
    if (!s) return;
    free(s->pixels);
    free(s);
}

inline int SDL_BlitSurface(SDL_Surface *src, const SDL_Rect *, SDL_Surface *dst, SDL_Rect *)
{
    const int w = std::min(src->w, dst->w), h = std::min(src->h, dst->h);
    for (int y = 0; y < h; ++y)
        memcpy((char *)dst->pixels + (size_t)y * dst->pitch, (const char *)src->pixels + (size_t)y * src->pitch, (size_t)w * 4);
    return 0;
}

inline int SDL_LockSurface(SDL_Surface *) { return 0; }
inline void SDL_UnlockSurface(SDL_Surface *) {}
inline const char *SDL_GetError() { return "sdl shim"; }

#endif
