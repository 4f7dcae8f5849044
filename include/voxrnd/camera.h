/**
   @file camera.h
   @brief Camera interface (reference: src/voxrnd/camera.h).

   The vtable keeps the reference's ten slots in the reference's order
   (object.h:29-30 then camera.h:25-116) so a camera written against the
   reference header still works.  One optional slot is APPENDED: `describe`,
   through which a camera tells the renderer how its screen2world can be
   evaluated on the device.  A camera that leaves it NULL is still rendered:
   its host screen2world is tabulated into a direction buffer each frame.
**/
#ifndef VOXRND_CAMERA_H
#define VOXRND_CAMERA_H

#include "object.h"

struct vox_camera;

#define VOX_CAMERA_DESC_SIMPLE 0  /**< dir = q (x) (mul*sx - xsub, 1, ysub - mul*sy), simple-camera.c:21-31 */
#define VOX_CAMERA_DESC_DOOM   1  /**< yaw + z shear, doom-camera.c:24-57 */

/** POD description of a camera for device-side ray generation. */
struct vox_camera_desc
{
    int kind;            /**< VOX_CAMERA_DESC_* */
    float position[3];
    float mul, xsub, ysub;
    float rotation[4];   /**< SIMPLE: quaternion (w,i,j,k) */
    float sinphi, cosphi, k; /**< DOOM */
};

struct vox_camera_interface
{
    VOX_OBJECT (vox_camera)
    void (*screen2world) (const struct vox_camera *camera, vox_dot ray, int sx, int sy);
    void (*rotate_camera) (struct vox_camera *camera, const vox_dot delta);
    void (*move_camera) (struct vox_camera *camera, const vox_dot delta);
    void (*look_at) (struct vox_camera *camera, const vox_dot coord);
    void (*get_position) (const struct vox_camera *camera, vox_dot res);
    void (*set_window_size) (struct vox_camera *camera, int w, int h);
    struct vox_camera* (*construct_camera) (const struct vox_camera *camera);
    void (*destroy_camera) (struct vox_camera *camera);
    /* ---- appended, optional ---- */
    int (*describe) (const struct vox_camera *camera, struct vox_camera_desc *desc);
    /**< Fill `desc`, return 1; return 0 if the camera cannot be described. */
};

/** A generic camera; properties "fov", "position", "rotation" (camera.h:119-137). */
struct vox_camera
{
    struct vox_camera_interface *iface;
};

/** Overwrite the camera's methods with the non-NULL ones of `iface` (camera.h:146-147, camera.c:6-17). */
VOX_EXPORT void vox_use_camera_methods (struct vox_camera *camera,
                                        const struct vox_camera_interface *iface);

/** Allocate `camera->iface` and fill it with dummy methods (camera.h:160, camera.c:19-41). */
VOX_EXPORT void vox_init_camera (struct vox_camera *camera);

/** Methods of the camera module `name` ("simple-camera", "doom-camera") or NULL (camera.h:175). */
VOX_EXPORT struct vox_camera_interface* vox_camera_methods (const char *name);

#endif
