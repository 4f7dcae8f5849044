/*
 * render_loop.h -- TEST INFRASTRUCTURE ONLY (see oracle_api.h).
 *
 * CPU restatement of the reference's vox_render block loop
 * (src/voxrnd/renderer.c:200-319), get_color (renderer.c:27-52, no light
 * manager) and the net effect of vox_redraw/copy_squares
 * (renderer.c:321-327, copy-helper.c:5-40: squares de-tiled into a linear
 * surface, pixels that hit nothing read 0).  renderer.c itself cannot be
 * compiled in this image (clang Blocks, libdispatch, SDL2, vn3d are absent),
 * so both oracle libraries share this restatement and plug in their own ray
 * search and camera: the port its restated ones, the reference glue the
 * reference's compiled functions.  OpenMP schedule(dynamic) over 4x4 blocks
 * stands in for dispatch_apply (renderer.c:214).
 */
#ifndef VOX_ORACLE_RENDER_LOOP_H
#define VOX_ORACLE_RENDER_LOOP_H

#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "oracle_api.h"

typedef struct
{
    const void *scene;   /* root node (may be NULL) */
    void *camera;
    /* vox_ray_tree_intersection: returns leaf or NULL, writes res on hit */
    const void* (*ray) (const void *node, const float *origin, const float *dir, float *res);
    void (*screen2world) (void *camera, float *dir, int sx, int sy);
    void (*get_position) (void *camera, float *pos);
    float (*sqr_metric) (const float *a, const float *b);  /* vox_sqr_metric, geom.c:133-139 */
    float (*sqr_norm) (const float *a);                    /* vox_sqr_norm, geom.c:141-147 */
    int32_t (*node_id) (void *idctx, const void *node); /* for the optional leaf_id output */
    void *idctx;
    float voxel[3];
    /* vox_get_light (lights.c:104-116) when a light manager is attached, else NULL */
    void (*get_light) (void *lights, const float *inter, float *light);
    void *lights;
} orc_render_hooks;

/* renderer.c:23-25 */
static const int orc_rendering_order[16] = {
    0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15
};

/* SDL_MapRGB for a >8bpp format: SDL2 src/video/SDL_pixels.c
 * (r >> Rloss) << Rshift | ... | Amask.  SDL2 is not vendored by the reference
 * and no reference test covers it: colour parity is "unpinned" (SURVEY 8c). */
static inline uint32_t orc_map_rgb (const orc_pixfmt *f, uint8_t r, uint8_t g, uint8_t b)
{
    return ((uint32_t)(r >> f->rloss) << f->rshift) |
           ((uint32_t)(g >> f->gloss) << f->gshift) |
           ((uint32_t)(b >> f->bloss) << f->bshift) | f->amask;
}

/* renderer.c:27-52 */
static inline uint32_t orc_get_color_lit (const orc_render_hooks *hk, const uint8_t *texture, const orc_pixfmt *fmt,
                                          const float voxel[3], const float inter[3])
{
    int x = abs ((int)(inter[0] / voxel[0])) & 31;
    int y = abs ((int)(inter[1] / voxel[1])) & 31;
    int z = abs ((int)(inter[2] / voxel[2])) & 31;
    uint8_t c = texture[x*32*32 + y*32 + z];
    if (!hk->get_light) return orc_map_rgb (fmt, c, c, c);
    float light[4];
    uint8_t r, g, b;
    hk->get_light (hk->lights, inter, light);
    r = fminf (light[0], 1.0) * c;
    g = fminf (light[1], 1.0) * c;
    b = fminf (light[2], 1.0) * c;
    return orc_map_rgb (fmt, r, g, b);
}
#define orc_get_color(texture, fmt, voxel, inter) orc_get_color_lit (hk, texture, fmt, voxel, inter)

/* ---- lights over a FLAT M-tree (the layout the product uploads, voxgpu_lights_flat): the
 * traversal of mtree.c:463-493 and the sum of lights.c:91-116, restated.  `sqr` is the
 * implementation's vox_sqr_metric. */
static inline void orc_flat_light (const orc_lights_flat *f, float (*sqr) (const float*, const float*),
                                   const float *p, float *light)
{
    light[0] = f->ambient[0]; light[1] = f->ambient[1]; light[2] = f->ambient[2];
    uint32_t i = 0;
    while (i < f->n_nodes)
    {
        const float *b = f->node_bounds + 4*i;
        const uint32_t *l = f->node_link + 4*i;
        if (!(sqr (b, p) < b[3])) { i = l[0]; continue; }
        if (l[1])
            for (uint32_t k=l[2]; k<l[2]+l[3]; k++)
            {
                const float *g = f->sphere_geom + 4*k, *c = f->sphere_light + 4*k;
                if (l[3] != 1 && !(sqr (g, p) < g[3])) continue;       /* mtree.c:475-482 */
                float dist = fminf (c[3], sqrtf (sqr (g, p)));           /* lights.c:95-101 */
                float add = 1 - dist / c[3];
                float r = light[0] + add * c[0];
                float gg = light[1] + add * c[1];
                float bb = light[2] + add * c[2];
                light[0] = r; light[1] = gg; light[2] = bb;
            }
        i++;
    }
}

#define ORC_LENGTH_THRESHOLD 1.69   /* renderer.c:197 */
#define ORC_MAX_DIST 150            /* renderer.c:198 */

static int orc_render_loop (const orc_render_hooks *hk, int w, int h, unsigned quality,
                            const uint8_t *texture, const orc_pixfmt *fmt,
                            uint32_t *pixels, int32_t *leaf_id, float *hit, int threads)
{
    /* renderer.c:97-100 geometry contract, renderer.c:179-195 quality contract */
    if ((w & 0xf) || (h & 0x3)) return 0;
    if ((quality & 12u) && ((quality & 3u) != ORC_QUALITY_ADAPTIVE)) return 0;
    if ((quality & 3u) > 2u || (quality & 12u) > 8u) return 0;

    const int ws = w >> 2, hs = h >> 2;
    const long squares_num = (long)ws * hs;
    const int rnd_mode = quality & 3u;
    const int merge_mode = quality & 12u;

    memset (pixels, 0, sizeof (uint32_t) * (size_t)w * h);
    if (leaf_id) for (size_t i=0; i<(size_t)w*h; i++) leaf_id[i] = -1;

    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
    for (long cs=0; cs<squares_num; cs++)
    {
        const void *corner1, *corner2, *leaf = NULL;
        float dir1[3], dir2[3], inter1[3], inter2[3], origin[3];
        uint32_t square[16];
        int written[16];
        int sq_leaf[16];
        float sq_hit[16][3];
        int block_rnd_mode = rnd_mode;
        int i, xstart, ystart, istart = 0, iend = 16;
        uint32_t color;

        memset (square, 0, sizeof square);
        memset (written, 0, sizeof written);
        ystart = (int)(cs / ws) << 2;
        xstart = (int)(cs % ws) << 2;

        hk->get_position (hk->camera, origin);
        int block_merge_mode = 0;

        if (block_rnd_mode == ORC_QUALITY_ADAPTIVE)
        {
            /* renderer.c:240-281 */
            block_merge_mode = (merge_mode == ORC_QUALITY_RAY_MERGE);
            istart = 1;
            hk->screen2world (hk->camera, dir1, xstart, ystart);
            corner1 = hk->ray (hk->scene, origin, dir1, inter1);
            block_rnd_mode = ORC_QUALITY_FAST;
            leaf = corner1;
            if (corner1 != NULL)
            {
                iend = 15;
                color = orc_get_color (texture, fmt, hk->voxel, inter1);
                square[0] = color; written[0] = 1;
                sq_leaf[0] = hk->node_id ? hk->node_id (hk->idctx, corner1) : 0;
                memcpy (sq_hit[0], inter1, sizeof inter1);
                hk->screen2world (hk->camera, dir2, xstart + 3, ystart + 3);
                corner2 = hk->ray (hk->scene, origin, dir2, inter2);
                if (corner2 != NULL)
                {
                    color = orc_get_color (texture, fmt, hk->voxel, inter2);
                    square[15] = color; written[15] = 1;
                    sq_leaf[15] = hk->node_id ? hk->node_id (hk->idctx, corner2) : 0;
                    memcpy (sq_hit[15], inter2, sizeof inter2);
                    float d1 = hk->sqr_metric (inter1, origin);
                    float d2 = hk->sqr_norm (dir1);
                    float criteria = d1 / d2 * hk->sqr_metric (dir1, dir2);
                    float dist = hk->sqr_metric (inter1, inter2);
                    /* float/float promoted to double against the double literal */
                    if (dist/criteria > ORC_LENGTH_THRESHOLD)
                        block_rnd_mode = ORC_QUALITY_BEST;

                    if (d1 > ORC_MAX_DIST * ORC_MAX_DIST)
                    {
                        block_merge_mode = (block_rnd_mode == ORC_QUALITY_FAST &&
                                            merge_mode == ORC_QUALITY_RAY_MERGE_ACCURATE)? 1:
                            block_merge_mode;
                    }
                    else block_merge_mode = 0;
                }
            }
        }

        int prev_p = 0;
        /* renderer.c:284-316 */
        for (i=istart; i<iend; i++)
        {
            int p = orc_rendering_order[i];
            int y = p/4;
            int x = p%4;
            int merge = block_merge_mode && (i & 1);

            hk->screen2world (hk->camera, dir1, x+xstart, y+ystart);
            if (!merge)
            {
                if (block_rnd_mode == ORC_QUALITY_FAST)
                {
                    if (leaf != NULL)
                        leaf = hk->ray (leaf, origin, dir1, inter1);
                    if (leaf == NULL)
                        leaf = hk->ray (hk->scene, origin, dir1, inter1);
                }
                else leaf = hk->ray (hk->scene, origin, dir1, inter1);
            }

            if (leaf != NULL)
            {
                color = (merge)? square[prev_p]: orc_get_color (texture, fmt, hk->voxel, inter1);
                square[p] = color; written[p] = 1;
                sq_leaf[p] = hk->node_id ? hk->node_id (hk->idctx, leaf) : 0;
                memcpy (sq_hit[p], inter1, sizeof inter1);
            }
            prev_p = p;
        }

        /* vox_redraw: copy_squares de-tiles square p -> (xstart + p%4, ystart + p/4) */
        for (int p=0; p<16; p++)
        {
            size_t pos = (size_t)(ystart + p/4) * w + (xstart + p%4);
            pixels[pos] = square[p];
            if (written[p])
            {
                if (leaf_id) leaf_id[pos] = sq_leaf[p];
                if (hit) memcpy (hit + 3*pos, sq_hit[p], 3*sizeof(float));
            }
        }
    }
    return 1;
}

#endif
