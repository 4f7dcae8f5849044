/*
 * cd.c -- collision detector (reference API: src/voxrnd/cd.h:25-60, behaviour: cd.c:20-55).
 *
 * The detector remembers the last camera position it accepted.  vox_cd_collide()
 * asks the GPU one ball query (the batched entry point, batch of one, through the
 * host-buffer pipeline: two small copies and a launch) whether a ball of the
 * camera's radius at the camera's new position touches the scene; if it does the
 * camera is put back where it was, otherwise the new position becomes the
 * accepted one.
 */
#include <stdlib.h>
#include <string.h>
#include "private.h"

struct vox_cd
{
    struct vox_rnd_ctx *ctx;        /* whose scene is probed */
    struct vox_camera *camera;      /* whose position is guarded */
    float radius;
    vox_dot accepted;               /* last position that did not collide */
};

struct vox_cd* vox_make_cd (void)
{
    return calloc (1, sizeof (struct vox_cd));
}

void vox_destroy_cd (struct vox_cd *cd)
{
    free (cd);
}

void vox_cd_attach_context (struct vox_cd *cd, struct vox_rnd_ctx *ctx)
{
    cd->ctx = ctx;
}

void vox_cd_attach_camera (struct vox_cd *cd, struct vox_camera *camera, float radius)
{
    cd->radius = radius;
    cd->camera = camera;
    camera->iface->get_position (camera, cd->accepted);     /* wherever it is now counts as free */
}

void vox_cd_gravity (struct vox_cd *cd, const vox_dot gravity)
{
    (void) cd; (void) gravity;      /* does nothing in the reference either (cd.c:33-36) */
}

void vox_cd_collide (struct vox_cd *cd)
{
    if (!cd->camera || !cd->ctx) return;
    vox_dot now;
    cd->camera->iface->get_position (cd->camera, now);
    if (vox_dot_equalp (now, cd->accepted)) return;          /* did not move: nothing to ask */

    uint8_t touched = 0;
    const float r = cd->radius;
    if (vox_tree_ball_collidep_batch (cd->ctx->scene, (const vox_dot*) &now, &r, 1, &touched) < 0) return;
    if (touched) cd->camera->iface->set_property_dot (cd->camera, "position", cd->accepted);
    else vox_dot_copy (cd->accepted, now);
}
