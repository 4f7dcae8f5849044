/* cd.c -- collision detector (reference: src/voxrnd/cd.c). */
#include <stdlib.h>
#include <string.h>
#include "private.h"

struct vox_cd
{
    struct vox_camera *camera;
    struct vox_rnd_ctx *ctx;
    vox_dot last_position;
    float camera_radius;
};

struct vox_cd* vox_make_cd (void)
{
    struct vox_cd *cd = vox_alloc (sizeof *cd);
    memset (cd, 0, sizeof *cd);
    return cd;
}

void vox_destroy_cd (struct vox_cd *cd) { free (cd); }

void vox_cd_attach_camera (struct vox_cd *cd, struct vox_camera *camera, float radius)
{
    /* cd.c:20-25 */
    cd->camera = camera;
    cd->camera_radius = radius;
    camera->iface->get_position (camera, cd->last_position);
}

void vox_cd_attach_context (struct vox_cd *cd, struct vox_rnd_ctx *ctx)
{
    cd->ctx = ctx;
}

void vox_cd_gravity (struct vox_cd *cd, const vox_dot gravity)
{
    /* a stub in the reference too (cd.c:33-36) */
    (void) cd; (void) gravity;
}

void vox_cd_collide (struct vox_cd *cd)
{
    /* cd.c:38-55 */
    struct vox_camera *camera = cd->camera;
    struct vox_rnd_ctx *ctx = cd->ctx;
    if (!camera || !ctx) return;
    vox_dot position;
    camera->iface->get_position (camera, position);
    if (vox_dot_equalp (position, cd->last_position)) return;
    if (vox_tree_ball_collidep (ctx->scene, position, cd->camera_radius))
        camera->iface->set_property_dot (camera, "position", cd->last_position);
    else
        vox_dot_copy (cd->last_position, position);
}
