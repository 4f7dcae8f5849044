/*
 * surface.c -- headless stand-in for the SDL window surface the reference
 * draws to (src/voxrnd/renderer.c:112-149), with its pixels in page-locked,
 * GPU-mapped host memory.
 */
#include <stdlib.h>
#include <string.h>
#include "private.h"

#define VOX_SURFACE_PINNED 0x56580001u   /* flags: pixels from voxgpu_host_alloc */
#define VOX_SURFACE_PLAIN  0x56580002u   /* flags: pixels from aligned_alloc */

SDL_Surface* vox_make_surface (unsigned int width, unsigned int height)
{
    SDL_Surface *s = calloc (1, sizeof *s);
    SDL_PixelFormat *f = calloc (1, sizeof *f);
    if (!s || !f) { free (s); free (f); return NULL; }
    /* SDL_PIXELFORMAT_RGB888: 32 bpp, X8 R8 G8 B8 */
    f->BitsPerPixel = 32; f->BytesPerPixel = 4;
    f->Rmask = 0x00ff0000u; f->Gmask = 0x0000ff00u; f->Bmask = 0x000000ffu; f->Amask = 0;
    f->Rshift = 16; f->Gshift = 8; f->Bshift = 0; f->Ashift = 0;
    f->Rloss = f->Gloss = f->Bloss = 0; f->Aloss = 8;
    f->refcount = 1;
    s->format = f;
    s->w = (int) width; s->h = (int) height;
    s->pitch = (int) width * 4;
    s->refcount = 1;
    size_t bytes = (size_t) width * height * 4;
    if (voxgpu_device_count () > 0 && (s->pixels = voxgpu_host_alloc (bytes)) != NULL)
        s->flags = VOX_SURFACE_PINNED;
    else
    {
        s->pixels = aligned_alloc (4096, (bytes + 4095) & ~(size_t) 4095);
        s->flags = VOX_SURFACE_PLAIN;
    }
    if (!s->pixels) { free (f); free (s); return NULL; }
    memset (s->pixels, 0, bytes);
    return s;
}

void vox_free_surface (SDL_Surface *s)
{
    if (!s) return;
    if (s->flags == VOX_SURFACE_PINNED) voxgpu_host_free (s->pixels);
    else if (s->flags == VOX_SURFACE_PLAIN) free (s->pixels);
    free (s->format);
    free (s);
}

Uint32 vox_map_rgb (const SDL_PixelFormat *f, Uint8 r, Uint8 g, Uint8 b)
{
    return ((Uint32) (r >> f->Rloss) << f->Rshift) |
           ((Uint32) (g >> f->Gloss) << f->Gshift) |
           ((Uint32) (b >> f->Bloss) << f->Bshift) | f->Amask;
}

int vox_surface_is_pinned (const SDL_Surface *s) { return s->flags == VOX_SURFACE_PINNED; }
