/**
   @file surface.h
   @brief The slice of SDL2's surface ABI the renderer touches.

   The reference renders into an `SDL_Surface` (renderer.h:8,44-50) and maps
   colours with `SDL_MapRGB` (renderer.c:39).  SDL2 is not part of this build
   environment, so unless VOX_HAVE_SDL2 is defined this header declares
   layout-compatible stand-ins for the leading members of SDL 2.0's public
   `SDL_Surface` / `SDL_PixelFormat` (the members the renderer reads: format,
   w, h, pitch, pixels; masks, losses, shifts).  With VOX_HAVE_SDL2 the real
   SDL headers are used and a surface obtained from SDL can be handed to
   vox_make_context_from_surface() unchanged.
**/
#ifndef VOXRND_SURFACE_H
#define VOXRND_SURFACE_H

#include "../voxvision.h"

#ifdef VOX_HAVE_SDL2
#include <SDL2/SDL.h>
#else
typedef uint8_t  Uint8;
typedef uint32_t Uint32;

typedef struct SDL_Window SDL_Window;   /* never instantiated without SDL */

typedef struct SDL_PixelFormat
{
    Uint32 format;
    void  *palette;
    Uint8  BitsPerPixel;
    Uint8  BytesPerPixel;
    Uint8  padding[2];
    Uint32 Rmask, Gmask, Bmask, Amask;
    Uint8  Rloss, Gloss, Bloss, Aloss;
    Uint8  Rshift, Gshift, Bshift, Ashift;
    int    refcount;
    struct SDL_PixelFormat *next;
} SDL_PixelFormat;

typedef struct SDL_Surface
{
    Uint32 flags;
    SDL_PixelFormat *format;
    int w, h;
    int pitch;
    void *pixels;
    void *userdata;
    int locked;
    void *list_blitmap;
    struct { int x, y, w, h; } clip_rect;
    void *map;
    int refcount;
} SDL_Surface;
#endif

/**
   \brief Create a headless 32-bpp surface (XRGB8888: R<<16 | G<<8 | B, the usual
   window-surface format) whose pixels live in page-locked host memory mapped
   into the GPU's address space, so the render kernel can store the frame
   straight into it.  Falls back to ordinary aligned memory when no CUDA device
   is present (rendering then fails, but contexts can still be built).
**/
VOX_EXPORT SDL_Surface* vox_make_surface (unsigned int width, unsigned int height);

/** Free a surface made by vox_make_surface(). */
VOX_EXPORT void vox_free_surface (SDL_Surface *surface);

/** SDL_MapRGB for a 32-bpp format: (r>>Rloss)<<Rshift | ... | Amask. */
VOX_EXPORT Uint32 vox_map_rgb (const SDL_PixelFormat *format, Uint8 r, Uint8 g, Uint8 b);

#endif
