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

/* the context behind the five public members (reference: renderer.h:25-38) */
struct vox_rnd_ctx_full
{
    struct vox_rnd_ctx pub;
    Uint8 *texture;               /* host copy, 32^3 */
    Uint8 *d_texture;             /* device copy */
    int texture_dirty;
    unsigned int quality;
    int owns_surface;
    int registered;               /* we page-locked a foreign surface */
    voxgpu_stream stream;         /* uploads (texture, tabulated directions) */
    voxgpu_present *present;      /* device framebuffer + band streams -> surface->pixels */
    int timed;
    float *dir_table, *d_dir_table;   /* foreign cameras: tabulated screen2world */
    int device;
};

void vox_default_texture (Uint8 *texture);

/* lights.c: the light set as the render kernel reads it, on the current device */
const voxgpu_lights* vox_light_manager_device (struct vox_light_manager *lm);

#endif
