/**
   @file cd.h
   @brief Collision detection (reference: src/voxrnd/cd.h).
**/
#ifndef VOXRND_CD_H
#define VOXRND_CD_H

#include "renderer.h"

struct vox_cd;

VOX_EXPORT struct vox_cd* vox_make_cd (void);
VOX_EXPORT void vox_cd_attach_camera (struct vox_cd *cd, struct vox_camera *camera, float radius);
VOX_EXPORT void vox_cd_attach_context (struct vox_cd *cd, struct vox_rnd_ctx *ctx);
/** If the camera moved into a voxel, put it back (cd.c:38-55); one ball query on the GPU. */
VOX_EXPORT void vox_cd_collide (struct vox_cd *cd);
/** A stub in the reference as well (cd.c:33-36). */
VOX_EXPORT void vox_cd_gravity (struct vox_cd *cd, const vox_dot gravity);
VOX_EXPORT void vox_destroy_cd (struct vox_cd *cd);

#endif
