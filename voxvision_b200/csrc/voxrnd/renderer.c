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
    ctx->want_devices = 1;
    const char *env = getenv ("VOXVISION_DEVICES");
    if (env) ctx->want_devices = (!strcmp (env, "all") || atoi (env) <= 0) ? 0 : atoi (env);
    return ctx;
}

/* Drop everything that lives on a device (the context stays usable: the next vox_render sets it up again). */
static void context_device_release (struct vox_rnd_ctx_full *ctx)
{
    int cur = voxgpu_get_device ();
    if (ctx->stream) voxgpu_synchronize (ctx->stream);
    voxgpu_multi_destroy (ctx->multi);
    ctx->multi = NULL;
    if (ctx->device >= 0) voxgpu_set_device (ctx->device);
    voxgpu_present_destroy (ctx->present);
    ctx->present = NULL;
    voxgpu_free (ctx->d_dir_table);
    ctx->d_dir_table = NULL;
    voxgpu_stream_destroy (ctx->stream);
    ctx->stream = NULL;
    for (int i=0; i<VOX_CTX_MAX_DEVICES; i++)
    {
        if (!ctx->d_texture[i]) continue;
        voxgpu_set_device (ctx->devices[i]);
        voxgpu_free (ctx->d_texture[i]);
        ctx->d_texture[i] = NULL;
    }
    if (cur >= 0) voxgpu_set_device (cur);
    ctx->device = -1;
    ctx->ndev = 0;
    ctx->texture_dirty = 1;
}

/* Device-side resources; only possible (and only needed) with a GPU present.  Restartable: a step that
   succeeded is not repeated when a later one failed and the next vox_render tries again. */
static int context_device_init (struct vox_rnd_ctx_full *ctx)
{
    if (ctx->device >= 0) return 1;
    const int visible = voxgpu_device_count ();
    if (visible <= 0)
    {
        vox_set_error ("no CUDA device available (%s); voxvision-b200 has no CPU renderer", voxgpu_last_error ());
        return 0;
    }
    SDL_Surface *s = ctx->pub.surface;
    size_t bytes = (size_t) s->pitch * s->h;
    if (!ctx->registered && !vox_surface_is_pinned (s))
    {
        /* a caller-provided surface: page-lock and map its pixels in place (for every device) */
        if (voxgpu_host_register (s->pixels, bytes))
        { vox_set_error ("cannot map surface pixels: %s", voxgpu_last_error ()); return 0; }
        ctx->registered = 1;
    }
    const int primary = voxgpu_get_device ();
    if (primary < 0) { vox_set_error ("context device setup failed: %s", voxgpu_last_error ()); return 0; }
    /* devices[0] = the current device, then the others in ordinal order */
    int want = ctx->want_devices <= 0 ? visible : ctx->want_devices;
    if (want > visible) want = visible;
    if (want > VOX_CTX_MAX_DEVICES) want = VOX_CTX_MAX_DEVICES;
    if (s->pitch != s->w * 4) want = 1;           /* shares of a frame are copied as blocks of four whole rows */
    int n = 0;
    ctx->devices[n++] = primary;
    for (int d=0; d<visible && n<want; d++) if (d != primary) ctx->devices[n++] = d;

    int ok = 1;
    for (int i=0; i<n && ok; i++)
    {
        if (ctx->d_texture[i]) continue;
        ok = voxgpu_set_device (ctx->devices[i]) == 0 && (ctx->d_texture[i] = voxgpu_malloc (VOX_TEXTURE_SIZE)) != NULL;
    }
    voxgpu_set_device (primary);
    if (ok && !ctx->stream) ok = (ctx->stream = voxgpu_stream_create ()) != NULL;
    if (ok && !ctx->present) ok = (ctx->present = voxgpu_present_create (s->w, s->h, 0)) != NULL;
    if (ok && n > 1 && !ctx->multi) ok = (ctx->multi = voxgpu_multi_create (s->w, s->h, n, ctx->devices)) != NULL;
    if (!ok) { vox_set_error ("context device setup failed: %s", voxgpu_last_error ()); return 0; }
    ctx->ndev = n;
    ctx->device = primary;
    ctx->texture_dirty = 1;
    return 1;
}

int vox_context_set_devices (struct vox_rnd_ctx *pub, int n)
{
    struct vox_rnd_ctx_full *ctx = (struct vox_rnd_ctx_full*) pub;
    const int visible = voxgpu_device_count ();
    if (n < 0) n = 1;
    if (n != ctx->want_devices && ctx->device >= 0) context_device_release (ctx);
    ctx->want_devices = n;
    if (visible <= 0) return 0;
    int use = n == 0 ? visible : (n < visible ? n : visible);
    return use > VOX_CTX_MAX_DEVICES ? VOX_CTX_MAX_DEVICES : use;
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
    context_device_release (ctx);
    if (ctx->registered) voxgpu_host_unregister (pub->surface->pixels);
    if (ctx->owns_surface) vox_free_surface (pub->surface);
    voxgpu_host_free (ctx->dir_table);
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
    return ctx->timed == 2 ? voxgpu_multi_last_ms (ctx->multi) : voxgpu_present_last_ms (ctx->present);
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
        vox_redraw (pub);                               /* a frame still on its way would land on top of it */
        memset (s->pixels, 0, (size_t) s->pitch * s->h);
        return;
    }
    if (!context_device_init (ctx)) return;
    if (voxgpu_get_device () != ctx->device && voxgpu_set_device (ctx->device))
    { vox_set_error ("vox_render: %s", voxgpu_last_error ()); return; }

    /* any node is a scene (renderer.c:297 searches inside returned leaves): the searches start there */
    struct vox_tree_info info;
    vox_tree_get_info (pub->scene, &info);

    voxgpu_camera cam;
    memset (&cam, 0, sizeof cam);
    struct vox_camera_desc desc;
    int described = pub->camera->iface->describe && pub->camera->iface->describe (pub->camera, &desc);
    if (described)
    {
        cam.kind = desc.kind == VOX_CAMERA_DESC_DOOM ? VOXGPU_CAMERA_DOOM : VOXGPU_CAMERA_SIMPLE;
        memcpy (cam.position, desc.position, sizeof cam.position);
        memcpy (cam.rotation, desc.rotation, sizeof cam.rotation);
        cam.mul = desc.mul; cam.xsub = desc.xsub; cam.ysub = desc.ysub;
        cam.sinphi = desc.sinphi; cam.cosphi = desc.cosphi; cam.k = desc.k;
    }
    /* a frame is split over the devices when there are several, the mode is BEST and the camera is a POD */
    const int split = ctx->multi && ctx->ndev > 1 && (ctx->quality & VOX_QUALITY_MODE_MASK) == VOX_QUALITY_BEST && described;
    const int ndev = split ? ctx->ndev : 1;

    if (ctx->texture_dirty)
    {
        /* (every device gets it, whether or not this frame is split) */
        int bad = 0;
        for (int i=0; i<ctx->ndev && !bad; i++)
            bad = voxgpu_set_device (ctx->devices[i]) || voxgpu_memcpy_h2d (ctx->d_texture[i], ctx->texture, VOX_TEXTURE_SIZE, NULL) ||
                  voxgpu_synchronize (NULL);
        voxgpu_set_device (ctx->device);
        if (bad) { vox_set_error ("vox_render: %s", voxgpu_last_error ()); return; }
        ctx->texture_dirty = 0;
    }

    if (!described)
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
    fr.start_node = info.node_id;

    /* the replicas of the tree (uploaded on first use, again after an edit) and of the lights
       (renderer.c:41-49: the kernel shades through a flattened copy of the manager's M-tree) */
    const voxgpu_tree *trees[VOX_CTX_MAX_DEVICES];
    const voxgpu_lights *lights[VOX_CTX_MAX_DEVICES];
    int ok = 1;
    for (int i=0; i<ndev && ok; i++)
    {
        if (ndev > 1 && voxgpu_set_device (ctx->devices[i])) { vox_set_error ("vox_render: %s", voxgpu_last_error ()); ok = 0; break; }
        trees[i] = vox_tree_device_handle (pub->scene);
        lights[i] = pub->light_manager ? vox_light_manager_device (pub->light_manager) : NULL;
        ok = trees[i] && (!pub->light_manager || lights[i]);
    }
    if (ndev > 1) voxgpu_set_device (ctx->device);
    if (!ok) return;

    if (split)
    {
        if (ctx->present) voxgpu_present_wait (ctx->present);   /* a frame of another mode still on its way */
        if (voxgpu_multi_render (ctx->multi, trees, &cam, &fr, (const uint8_t *const *) ctx->d_texture,
                                 pub->light_manager ? lights : NULL, s->pixels, (size_t) s->pitch))
        { vox_set_error ("vox_render: %s", voxgpu_last_error ()); return; }
        ctx->timed = 2;
        return;
    }
    if (ctx->multi) voxgpu_multi_wait (ctx->multi);
    fr.lights = lights[0];
    if (voxgpu_present_render (ctx->present, trees[0], &cam, &fr, ctx->d_texture[0], s->pixels, (size_t) s->pitch))
    { vox_set_error ("vox_render: %s", voxgpu_last_error ()); return; }
    ctx->timed = 1;
}

void vox_redraw (struct vox_rnd_ctx *pub)
{
    /* renderer.c:321-327: after this call surface->pixels holds the frame.
       There is no window to update. */
    struct vox_rnd_ctx_full *ctx = (struct vox_rnd_ctx_full*) pub;
    if (ctx->multi && voxgpu_multi_wait (ctx->multi))
        vox_set_error ("vox_redraw: %s", voxgpu_last_error ());
    if (ctx->present && voxgpu_present_wait (ctx->present))
        vox_set_error ("vox_redraw: %s", voxgpu_last_error ());
}
