/**
   @file renderer.h
   @brief The renderer (reference: src/voxrnd/renderer.h).
**/
#ifndef VOXRND_RENDERER_H
#define VOXRND_RENDERER_H

#include "surface.h"
#include "../voxtrees/tree.h"
#include "camera.h"

struct vox_light_manager;   /* lights.h */

/**
   \brief A renderer context.  The five leading members are public, exactly as
   in the reference (renderer.h:44-72); everything after them is private.
**/
struct vox_rnd_ctx
{
    SDL_Window *window;       /**< always NULL here: there is no windowing */
    SDL_Surface *surface;     /**< surface the context draws to */
    struct vox_node *scene;   /**< set with vox_context_set_scene() */
    struct vox_camera *camera;/**< set with vox_context_set_camera() */
    struct vox_light_manager *light_manager; /**< set with vox_context_set_light_manager(), may be NULL */
};

/* Quality bits, renderer.h:81-127 */
#define VOX_QUALITY_BEST 0
#define VOX_QUALITY_FAST 1
#define VOX_QUALITY_ADAPTIVE 2
#define VOX_QUALITY_RAY_MERGE 4
#define VOX_QUALITY_RAY_MERGE_ACCURATE 8
#define VOX_QUALITY_RM_MASK 12
#define VOX_QUALITY_MODE_MASK 3

#define VOX_TEXTURE_SIDE 32
#define VOX_TEXTURE_SIZE (32*32*32)

/**
   \brief Make a renderer context from a surface (renderer.h:145, renderer.c:102-110).
   Width must be a multiple of 16 and height of 4, else NULL.  The surface is not
   owned by the context.  Its pixel memory is page-locked and mapped for the GPU
   for the lifetime of the context.
**/
VOX_EXPORT struct vox_rnd_ctx* vox_make_context_from_surface (SDL_Surface *surface);

/**
   \brief Create a context with its own (headless) surface (renderer.h:161).
   The reference opens an SDL window here (renderer.c:112-149); without SDL the
   context owns a vox_make_surface() surface and `window` stays NULL.
**/
VOX_EXPORT struct vox_rnd_ctx* vox_make_context_and_window (unsigned int width, unsigned int height);

/**
   \brief Finish the frame (renderer.h:169, renderer.c:321-327): wait for the
   kernel launched by vox_render(); afterwards `surface->pixels` holds the
   frame in linear layout with 0 in every pixel that hit nothing.
**/
VOX_EXPORT void vox_redraw (struct vox_rnd_ctx *ctx);

VOX_EXPORT void vox_context_set_scene (struct vox_rnd_ctx *ctx, struct vox_node *scene);
VOX_EXPORT void vox_context_set_camera (struct vox_rnd_ctx *ctx, struct vox_camera *camera);
VOX_EXPORT void vox_context_set_light_manager (struct vox_rnd_ctx *ctx,
                                               struct vox_light_manager *light_manager);

/** \brief Set rendering quality; 1 on success, 0 for an invalid value (renderer.c:179-195). **/
VOX_EXPORT int vox_context_set_quality (struct vox_rnd_ctx *ctx, unsigned int quality);

VOX_EXPORT void vox_destroy_context (struct vox_rnd_ctx *ctx);

/**
   \brief Render the scene (renderer.h:208, renderer.c:200-319): launches the
   primary-ray kernel on the context's stream and returns; callable from any
   thread.  vox_redraw() completes the frame.
**/
VOX_EXPORT void vox_render (struct vox_rnd_ctx *ctx);

/* ---- additions ---- */

/**
   \brief Replace the 32x32x32 grey texture (renderer.c:54-71).  The reference
   fills it from the un-vendored vn3d noise library; this build ships its own
   deterministic value-noise table, and callers wanting the reference's bytes
   can install them here.
**/
VOX_EXPORT void vox_context_set_texture (struct vox_rnd_ctx *ctx, const Uint8 *texture);
VOX_EXPORT const Uint8* vox_context_get_texture (const struct vox_rnd_ctx *ctx);

/**
   \brief Draw every frame of this context on up to `n` CUDA devices of the process (0 = all visible
   devices, 1 = only the device that is current at the first vox_render(): the default, or what
   $VOXVISION_DEVICES says).  This is the counterpart of the reference's fan-out over all cores inside
   vox_render (dispatch_apply over the blocks, renderer.c:214-226): device k draws the 4-row tile rows
   r with r % n == k and copies them into surface->pixels over its own PCIe link; the tree, the texture
   and the lights are replicated on first use.  Applies to quality BEST; the other quality modes, whose
   blocks depend on their own previous pixels, stay on the first device.  Returns the number of devices
   that will be used, 0 if there is no device.  vox_render / vox_redraw are called as before.
**/
VOX_EXPORT int vox_context_set_devices (struct vox_rnd_ctx *ctx, int n);

/** Time in milliseconds the GPU spent in the last completed vox_render() kernel. */
VOX_EXPORT float vox_context_last_kernel_ms (struct vox_rnd_ctx *ctx);

#endif
