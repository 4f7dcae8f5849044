/* private.h -- internals of libvoxrnd. */
#ifndef VOXRND_PRIVATE_H
#define VOXRND_PRIVATE_H

#include "voxrnd.h"
#include "voxtrees.h"
#include "voxgpu.h"

/* error channel lives in libvoxtrees */
VOX_EXPORT void vox_set_error (const char *fmt, ...);

/* dummy methods, reference: src/voxrnd/object.c:4-9 */
void vox_method_void_dummy (void *obj, ...);
void vox_method_dot_dummy (void *obj, vox_dot dot, ...);

/* quaternion helpers, reference: src/voxrnd/vect-ops.h:19-35 */
VOX_EXPORT void vox_rotate_vector (const vox_quat base, const vox_dot vector, vox_dot res);
VOX_EXPORT void vox_rotate_vector_q (const vox_quat base, const vox_quat vector, vox_quat res);
VOX_EXPORT void vox_quat_mul (const vox_quat q1, const vox_quat q2, vox_quat res);
VOX_EXPORT void vox_dot_normalize (vox_dot dot);
VOX_EXPORT void vox_quat_normalize (vox_quat quat);
VOX_EXPORT void vox_quat_set_identity (vox_quat quat);

#define VOX_CTX_MAX_DEVICES 16

/* the context behind the five public members (reference: renderer.h:25-38) */
struct vox_rnd_ctx_full
{
    struct vox_rnd_ctx pub;
    Uint8 *texture;               /* host copy, 32^3 */
    int texture_dirty;
    unsigned int quality;
    int owns_surface;
    int registered;               /* we page-locked a foreign surface */
    voxgpu_stream stream;         /* uploads (texture, tabulated directions) on the primary device */
    voxgpu_present *present;      /* primary device: framebuffer + band streams -> surface->pixels */
    int timed;                    /* 0: nothing to time, 1: `present` drew the last frame, 2: `multi` did */
    float *dir_table, *d_dir_table;   /* foreign cameras: tabulated screen2world */
    int device;                   /* primary device (the device current at the first vox_render), -1: not set up */
    /* a frame split over several devices of this process (vox_context_set_devices) */
    int want_devices;             /* 1: primary only (default), 0: all visible, n: up to n */
    int ndev;                     /* devices in use once set up; devices[0] is the primary */
    int devices[VOX_CTX_MAX_DEVICES];
    Uint8 *d_texture[VOX_CTX_MAX_DEVICES];   /* device copies, by position in devices[] */
    voxgpu_multi *multi;          /* ndev > 1: one worker + presenter per device */
};

void vox_default_texture (Uint8 *texture);

/* lights.c: the light set as the render kernel reads it, on the current device */
const voxgpu_lights* vox_light_manager_device (struct vox_light_manager *lm);

#endif
