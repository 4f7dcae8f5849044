/*
 * screenshot.c -- vox_screenshot (reference: src/voxrnd/screenshot.c:7-36), for looking at what the
 * GPU drew.  The reference hands the surface to SDL_SaveBMP; here the 24-bit BMP is written
 * directly (bottom-up rows of B, G, R, padded to 4 bytes), channels taken apart with the
 * surface's own masks and shifts.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <sys/stat.h>
#include <sys/param.h>
#include "private.h"

static void put32 (unsigned char *p, unsigned v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; }

int vox_screenshot (const struct vox_rnd_ctx *ctx, const char *dirname)
{
    static int next = 0;                               /* screenshot.c:9: numbering continues per process */
    const char *data = getenv ("VOXVISION_DATA");
    const char *dir = dirname ? dirname : (data ? data : ".");
    char name[MAXPATHLEN];
    struct stat sb;
    for (;;)
    {
        if (snprintf (name, sizeof name, "%s/screenshot%i.bmp", dir, next++) >= (int) sizeof name)
        { vox_set_error ("vox_screenshot: path too long"); return 0; }
        if (stat (name, &sb) != 0) break;
    }
    if (errno != ENOENT) { vox_set_error ("vox_screenshot: %s: %s", name, strerror (errno)); return 0; }

    const SDL_Surface *s = ctx->surface;
    const SDL_PixelFormat *f = s->format;
    const int w = s->w, h = s->h, row = (3*w + 3) & ~3;
    unsigned char header[54] = { 'B', 'M' };
    put32 (header + 2, 54u + (unsigned) row * h);
    put32 (header + 10, 54);
    put32 (header + 14, 40);
    put32 (header + 18, (unsigned) w);
    put32 (header + 22, (unsigned) h);
    header[26] = 1; header[28] = 24;
    put32 (header + 34, (unsigned) row * h);
    put32 (header + 38, 2835); put32 (header + 42, 2835);

    FILE *out = fopen (name, "wb");
    unsigned char *line = calloc (1, row);
    if (!out || !line) { vox_set_error ("vox_screenshot: %s: %s", name, strerror (errno)); free (line); if (out) fclose (out); return 0; }
    int ok = fwrite (header, 1, sizeof header, out) == sizeof header;
    for (int y=h-1; y>=0 && ok; y--)
    {
        const Uint32 *pix = (const Uint32*) ((const char*) s->pixels + (size_t) y * s->pitch);
        for (int x=0; x<w; x++)
        {
            line[3*x]   = (Uint8) (((pix[x] & f->Bmask) >> f->Bshift) << f->Bloss);
            line[3*x+1] = (Uint8) (((pix[x] & f->Gmask) >> f->Gshift) << f->Gloss);
            line[3*x+2] = (Uint8) (((pix[x] & f->Rmask) >> f->Rshift) << f->Rloss);
        }
        ok = fwrite (line, 1, row, out) == (size_t) row;
    }
    free (line);
    if (fclose (out) != 0) ok = 0;
    if (!ok) vox_set_error ("vox_screenshot: writing %s failed", name);
    return ok;
}
