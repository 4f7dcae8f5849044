/*
 * texture.c -- the 32x32x32 grey texture voxels are shaded with.
 *
 * The reference fills it with 4-octave value noise from the external vn3d
 * library (src/voxrnd/renderer.c:54-71: vn_value_generator(3, 4),
 * vn_noise_3d(...) >> 24).  vn3d is neither vendored nor pinned by the
 * reference, so its bytes cannot be reproduced here; this is an independent,
 * deterministic, integer-only 4-octave value noise with the same role
 * (tileable with period 32, one byte per cell).  vox_context_set_texture()
 * installs any other table.
 */
#include "private.h"

static uint32_t lattice (uint32_t x, uint32_t y, uint32_t z, uint32_t octave)
{
    /* integer hash of a lattice point -> 16-bit value */
    uint32_t h = x * 0x8da6b343u ^ y * 0xd8163841u ^ z * 0xcb1ab31fu ^ octave * 0x9e3779b9u;
    h ^= h >> 15; h *= 0x2c1b3c6du;
    h ^= h >> 12; h *= 0x297a2d39u;
    h ^= h >> 15;
    return h & 0xffffu;
}

/* smoothstep on a fraction f/period, 16.16 fixed point */
static uint32_t fade (uint32_t f, uint32_t period)
{
    uint64_t t = ((uint64_t) f << 16) / period;           /* 0 .. 65535 */
    return (uint32_t) ((t * t * (3 * 65536ull - 2 * t)) >> 32);
}

static uint32_t lerp16 (uint32_t a, uint32_t b, uint32_t t)
{
    return (uint32_t) (((uint64_t) a * (65536u - t) + (uint64_t) b * t) >> 16);
}

static uint32_t octave_value (uint32_t x, uint32_t y, uint32_t z, uint32_t octave)
{
    uint32_t period = 16u >> octave;                      /* 16, 8, 4, 2 cells */
    uint32_t cells = 32u / period;
    uint32_t x0 = x / period, y0 = y / period, z0 = z / period;
    uint32_t x1 = (x0 + 1) % cells, y1 = (y0 + 1) % cells, z1 = (z0 + 1) % cells;
    uint32_t tx = fade (x % period, period), ty = fade (y % period, period), tz = fade (z % period, period);
    uint32_t c00 = lerp16 (lattice (x0,y0,z0,octave), lattice (x1,y0,z0,octave), tx);
    uint32_t c10 = lerp16 (lattice (x0,y1,z0,octave), lattice (x1,y1,z0,octave), tx);
    uint32_t c01 = lerp16 (lattice (x0,y0,z1,octave), lattice (x1,y0,z1,octave), tx);
    uint32_t c11 = lerp16 (lattice (x0,y1,z1,octave), lattice (x1,y1,z1,octave), tx);
    return lerp16 (lerp16 (c00, c10, ty), lerp16 (c01, c11, ty), tz);
}

void vox_default_texture (Uint8 *texture)
{
    int p = 0;
    for (uint32_t i=0; i<32; i++)
        for (uint32_t j=0; j<32; j++)
            for (uint32_t k=0; k<32; k++)
            {
                /* octave weights 8:4:2:1 */
                uint32_t v = 8*octave_value (i,j,k,0) + 4*octave_value (i,j,k,1) +
                             2*octave_value (i,j,k,2) + octave_value (i,j,k,3);
                v /= 15;                                  /* 0 .. 65535 */
                texture[p++] = (Uint8) (64 + (v * 191u >> 16));   /* keep away from black: 0 means "no hit" */
            }
}
