/*
 * renderer.c -- renderer context and vox_render / vox_redraw (reference:
 * src/voxrnd/renderer.c).  The per-block CPU loop of renderer.c:214-318 is a
 * handful of kernel launches (one per band of the frame, voxgpu_present_render);
 * every band is DMA-copied into the surface's page-locked pixels in linear
 * layout while the next band renders, so copy_squares (copy-helper.c:5-40) has
 * no counterpart.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "private.h"

int vox_surface_is_pinned (const SDL_Surface *s);

static int bad_geometry (unsigned int width, unsigned int height)
{
    /* renderer.c:97-100 */
    return (width & 0xf) || (height & 0x3);
}

static struct vox_rnd_ctx_full* allocate_context (void)
{
    /* renderer.c:73-81: zeroed, texture made, ADAPTIVE by default */
    struct vox_rnd_ctx_full *ctx = calloc (1, sizeof *ctx);
    if (!ctx) return NULL;
    ctx->texture = malloc (VOX_TEXTURE_SIZE);
    vox_default_texture (ctx->texture);
    ctx->texture_dirty = 1;
    ctx->quality = VOX_QUALITY_ADAPTIVE;
    ctx->device = -1;
    return ctx;
}

/* Device-side resources; only possible (and only needed) with a GPU present. */
static int context_device_init (struct vox_rnd_ctx_full *ctx)
{
    if (ctx->device >= 0) return 1;
    if (voxgpu_device_count () <= 0)
    {
        vox_set_error ("no CUDA device available (%s); voxvision-b200 has no CPU renderer", voxgpu_last_error ());
        return 0;
    }
    SDL_Surface *s = ctx->pub.surface;
    size_t bytes = (size_t) s->pitch * s->h;
    if (!vox_surface_is_pinned (s))
    {
        /* a caller-provided surface: page-lock and map its pixels in place */
        if (voxgpu_host_register (s->pixels, bytes))
        { vox_set_error ("cannot map surface pixels: %s", voxgpu_last_error ()); return 0; }
        ctx->registered = 1;
    }
    ctx->d_texture = voxgpu_malloc (VOX_TEXTURE_SIZE);
    ctx->stream = voxgpu_stream_create ();
    ctx->present = voxgpu_present_create (s->w, s->h, 0);
    if (!ctx->d_texture || !ctx->stream || !ctx->present)
    { vox_set_error ("context device setup failed: %s", voxgpu_last_error ()); return 0; }
    ctx->device = voxgpu_get_device ();
    return 1;
}

struct vox_rnd_ctx* vox_make_context_from_surface (SDL_Surface *surface)
{
    /* renderer.c:102-110 */
    if (!surface || bad_geometry (surface->w, surface->h)) return NULL;
    if (!surface->format || surface->format->BitsPerPixel != 32) return NULL;   /* renderer.c:139 */
    struct vox_rnd_ctx_full *ctx = allocate_context ();
    if (!ctx) return NULL;
    ctx->pub.surface = surface;
    return &ctx->pub;
}

struct vox_rnd_ctx* vox_make_context_and_window (unsigned int width, unsigned int height)
{
    /* renderer.c:112-149, minus the SDL window */
    if (bad_geometry (width, height)) return NULL;
    struct vox_rnd_ctx_full *ctx = allocate_context ();
    if (!ctx) return NULL;
    ctx->pub.surface = vox_make_surface (width, height);
    if (!ctx->pub.surface) { vox_destroy_context (&ctx->pub); return NULL; }
    ctx->owns_surface = 1;
    return &ctx->pub;
}

void vox_destroy_context (struct vox_rnd_ctx *pub)
{
    /* renderer.c:151-161 */
    struct vox_rnd_ctx_full *ctx = (struct vox_rnd_ctx_full*) pub;
    if (!ctx) return;
    if (ctx->stream) voxgpu_synchronize (ctx->stream);
    voxgpu_present_destroy (ctx->present);
    if (ctx->registered) voxgpu_host_unregister (pub->surface->pixels);
    if (ctx->owns_surface) vox_free_surface (pub->surface);
    voxgpu_free (ctx->d_texture);
    voxgpu_free (ctx->d_dir_table);
    voxgpu_host_free (ctx->dir_table);
    voxgpu_stream_destroy (ctx->stream);
    free (ctx->texture);
    free (ctx);
}

void vox_context_set_camera (struct vox_rnd_ctx *ctx, struct vox_camera *camera)
{
    /* renderer.c:163-167 */
    ctx->camera = camera;
    camera->iface->set_window_size (camera, ctx->surface->w, ctx->surface->h);
}

void vox_context_set_scene (struct vox_rnd_ctx *ctx, struct vox_node *scene)
{
    ctx->scene = scene;
}

void vox_context_set_light_manager (struct vox_rnd_ctx *ctx, struct vox_light_manager *light_manager)
{
    /* renderer.c:173-177; NULL detaches */
    ctx->light_manager = light_manager;
}

int vox_context_set_quality (struct vox_rnd_ctx *pub, unsigned int quality)
{
    /* renderer.c:179-195 */
    struct vox_rnd_ctx_full *ctx = (struct vox_rnd_ctx_full*) pub;
    if ((quality & VOX_QUALITY_RM_MASK) &&
        ((quality & VOX_QUALITY_MODE_MASK) != VOX_QUALITY_ADAPTIVE)) return 0;
    if ((quality & VOX_QUALITY_MODE_MASK) > VOX_QUALITY_ADAPTIVE) return 0;
    if ((quality & VOX_QUALITY_RM_MASK) > VOX_QUALITY_RAY_MERGE_ACCURATE) return 0;
    ctx->quality = quality;
    return 1;
}

void vox_context_set_texture (struct vox_rnd_ctx *pub, const Uint8 *texture)
{
    struct vox_rnd_ctx_full *ctx = (struct vox_rnd_ctx_full*) pub;
    memcpy (ctx->texture, texture, VOX_TEXTURE_SIZE);
    ctx->texture_dirty = 1;
}

const Uint8* vox_context_get_texture (const struct vox_rnd_ctx *pub)
{
    return ((const struct vox_rnd_ctx_full*) pub)->texture;
}

float vox_context_last_kernel_ms (struct vox_rnd_ctx *pub)
{
    struct vox_rnd_ctx_full *ctx = (struct vox_rnd_ctx_full*) pub;
    if (!ctx->timed) return -1.0f;
    return voxgpu_present_last_ms (ctx->present);
}

/* A camera without `describe`: evaluate its host screen2world for every pixel. */
static int tabulate_camera (struct vox_rnd_ctx_full *ctx, voxgpu_camera *cam)
{
    struct vox_camera *camera = ctx->pub.camera;
    int w = ctx->pub.surface->w, h = ctx->pub.surface->h;
    size_t bytes = (size_t) w * h * 4 * sizeof (float);
    if (!ctx->dir_table)
    {
        ctx->dir_table = voxgpu_host_alloc (bytes);
        ctx->d_dir_table = voxgpu_malloc (bytes);
        if (!ctx->dir_table || !ctx->d_dir_table) return 0;
    }
    for (int y=0; y<h; y++)
        for (int x=0; x<w; x++)
        {
            vox_dot d;
            camera->iface->screen2world (camera, d, x, y);
            memcpy (ctx->dir_table + 4*((size_t) y*w + x), d, 3*sizeof (float));
        }
    if (voxgpu_memcpy_h2d (ctx->d_dir_table, ctx->dir_table, bytes, ctx->stream) ||
        voxgpu_synchronize (ctx->stream)) return 0;
    cam->kind = VOXGPU_CAMERA_TABLE;
    cam->dir_table = ctx->d_dir_table;
    return 1;
}

void vox_render (struct vox_rnd_ctx *pub)
{
    struct vox_rnd_ctx_full *ctx = (struct vox_rnd_ctx_full*) pub;
    SDL_Surface *s = pub->surface;
    ctx->timed = 0;
    vox_set_error ("");
    if (!pub->camera) { vox_set_error ("vox_render: no camera"); return; }
    if (!pub->scene)
    {
        /* every ray misses a NULL tree (search.c:22): the frame is all zeros */
        memset (s->pixels, 0, (size_t) s->pitch * s->h);
        return;
    }
    if (!context_device_init (ctx)) return;
    voxgpu_tree *tree = vox_tree_device_handle (pub->scene);
    if (!tree) return;

    if (ctx->texture_dirty)
    {
        if (voxgpu_memcpy_h2d (ctx->d_texture, ctx->texture, VOX_TEXTURE_SIZE, ctx->stream) ||
            voxgpu_synchronize (ctx->stream))
        { vox_set_error ("vox_render: %s", voxgpu_last_error ()); return; }
        ctx->texture_dirty = 0;
    }

    voxgpu_camera cam;
    memset (&cam, 0, sizeof cam);
    struct vox_camera_desc desc;
    if (pub->camera->iface->describe && pub->camera->iface->describe (pub->camera, &desc))
    {
        cam.kind = desc.kind == VOX_CAMERA_DESC_DOOM ? VOXGPU_CAMERA_DOOM : VOXGPU_CAMERA_SIMPLE;
        memcpy (cam.position, desc.position, sizeof cam.position);
        memcpy (cam.rotation, desc.rotation, sizeof cam.rotation);
        cam.mul = desc.mul; cam.xsub = desc.xsub; cam.ysub = desc.ysub;
        cam.sinphi = desc.sinphi; cam.cosphi = desc.cosphi; cam.k = desc.k;
    }
    else
    {
        vox_dot pos;
        pub->camera->iface->get_position (pub->camera, pos);
        memcpy (cam.position, pos, sizeof cam.position);
        if (!tabulate_camera (ctx, &cam)) { vox_set_error ("vox_render: %s", voxgpu_last_error ()); return; }
    }

    voxgpu_frame fr;
    memset (&fr, 0, sizeof fr);
    fr.w = s->w; fr.h = s->h; fr.y0 = 0; fr.y1 = s->h;
    fr.pitch_pixels = s->pitch / 4;
    fr.quality = ctx->quality;
    fr.fmt.amask = s->format->Amask;
    fr.fmt.rloss = s->format->Rloss; fr.fmt.gloss = s->format->Gloss; fr.fmt.bloss = s->format->Bloss;
    fr.fmt.rshift = s->format->Rshift; fr.fmt.gshift = s->format->Gshift; fr.fmt.bshift = s->format->Bshift;

    if (pub->light_manager)
    {
        /* renderer.c:41-49: the kernel shades through a flattened copy of the manager's M-tree */
        fr.lights = vox_light_manager_device (pub->light_manager);
        if (!fr.lights) return;
    }

    if (voxgpu_present_render (ctx->present, tree, &cam, &fr, ctx->d_texture, s->pixels, (size_t) s->pitch))
    { vox_set_error ("vox_render: %s", voxgpu_last_error ()); return; }
    ctx->timed = 1;
}

void vox_redraw (struct vox_rnd_ctx *pub)
{
    /* renderer.c:321-327: after this call surface->pixels holds the frame.
       There is no window to update. */
    struct vox_rnd_ctx_full *ctx = (struct vox_rnd_ctx_full*) pub;
    if (ctx->present && voxgpu_present_wait (ctx->present))
        vox_set_error ("vox_redraw: %s", voxgpu_last_error ());
}
