/*
 * doom-camera.c -- 5-DoF yaw + z-shear camera module (reference:
 * src/voxrnd/doom-camera.c).  Built as doom-camera.so.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "private.h"

struct doom_camera
{
    struct vox_camera_interface *iface;
    float xsub, ysub;
    vox_dot position;
    float phi, sinphi, cosphi, k;
    float mul, fov;
};

VOX_EXPORT struct vox_module* module_init (void);

static void yaw (const struct doom_camera *c, const float *v, float *res)
{
    /* doom-camera.c:24-34 transform_vector */
    float x = v[0]*c->cosphi - v[1]*c->sinphi;
    float y = v[1]*c->cosphi + v[0]*c->sinphi;
    float z = v[2];
    res[0] = x; res[1] = y; res[2] = z;
}

static void d_screen2world (const struct vox_camera *cam, vox_dot ray, int sx, int sy)
{
    /* doom-camera.c:36-57 */
    const struct doom_camera *c = (const void*) cam;
    vox_dot dir;
    dir[0] = c->mul*sx - c->xsub;
    dir[1] = 1;
    dir[2] = c->ysub - c->mul*sy;
    dir[2] += c->k;
    dir[3] = 0;
    yaw (c, dir, ray);
}

static void d_get_position (const struct vox_camera *cam, vox_dot res)
{
    const struct doom_camera *c = (const void*) cam;
    vox_dot_copy (res, c->position);
}

static void d_set_property_number (struct vox_camera *cam, const char *name, float value)
{
    struct doom_camera *c = (void*) cam;
    if (strcmp (name, "fov") == 0) c->fov = value;
}

static void d_set_property_dot (struct vox_camera *cam, const char *name, const vox_dot value)
{
    /* doom-camera.c:72-83 */
    struct doom_camera *c = (void*) cam;
    if (strcmp (name, "position") == 0) vox_dot_copy (c->position, value);
    else if (strcmp (name, "rotation") == 0)
    {
        c->k = value[0];
        c->phi = value[2];
        c->sinphi = sinf (c->phi);
        c->cosphi = cosf (c->phi);
    }
}

static void d_move_camera (struct vox_camera *cam, const vox_dot delta)
{
    struct doom_camera *c = (void*) cam;
    vox_dot tr;
    yaw (c, delta, tr);
    for (int i=0; i<3; i++) c->position[i] = c->position[i] + tr[i];
}

static void d_rotate_camera (struct vox_camera *cam, const vox_dot delta)
{
    /* doom-camera.c:94-102 */
    struct doom_camera *c = (void*) cam;
    c->k += delta[0];
    c->phi += delta[2];
    c->sinphi = sinf (c->phi);
    c->cosphi = cosf (c->phi);
}

static void d_look_at (struct vox_camera *cam, const vox_dot coord)
{
    /* doom-camera.c:104-115 */
    struct doom_camera *c = (void*) cam;
    float sub[3];
    for (int i=0; i<3; i++) sub[i] = coord[i] - c->position[i];
    c->k = sub[2] / hypotf (sub[0], sub[1]);
    c->phi = -atan2f (sub[0], sub[1]);
    c->cosphi = cosf (c->phi);
    c->sinphi = sinf (c->phi);
}

static void d_set_window_size (struct vox_camera *cam, int w, int h)
{
    /* doom-camera.c:117-136 */
    struct doom_camera *c = (void*) cam;
    if (w > h)
    {
        c->xsub = c->fov;
        c->ysub = c->fov*h/w;
        c->mul  = c->fov*2/w;
    }
    else
    {
        c->ysub = c->fov;
        c->xsub = c->fov*w/h;
        c->mul  = c->fov*2/h;
    }
}

static struct vox_camera* d_construct_camera (const struct vox_camera *cam)
{
    /* doom-camera.c:138-158 */
    const struct doom_camera *old = (const void*) cam;
    struct doom_camera *c = vox_alloc (sizeof *c);
    if (old) memcpy (c, old, sizeof *c);
    else
    {
        memset (c, 0, sizeof *c);
        c->fov = 1.0;
        c->cosphi = 1.0;
    }
    vox_init_camera ((struct vox_camera*) c);
    if (old) vox_use_camera_methods ((struct vox_camera*) c, old->iface);
    vox_use_camera_methods ((struct vox_camera*) c, (struct vox_camera_interface*) module_init ()->methods);
    return (struct vox_camera*) c;
}

static int d_describe (const struct vox_camera *cam, struct vox_camera_desc *desc)
{
    const struct doom_camera *c = (const void*) cam;
    memset (desc, 0, sizeof *desc);
    desc->kind = VOX_CAMERA_DESC_DOOM;
    memcpy (desc->position, c->position, 3*sizeof (float));
    desc->mul = c->mul; desc->xsub = c->xsub; desc->ysub = c->ysub;
    desc->sinphi = c->sinphi; desc->cosphi = c->cosphi; desc->k = c->k;
    return 1;
}

static struct vox_camera_interface doom_camera_interface =
{
    .set_property_number = d_set_property_number,
    .set_property_dot = d_set_property_dot,
    .screen2world = d_screen2world,
    .rotate_camera = d_rotate_camera,
    .move_camera = d_move_camera,
    .look_at = d_look_at,
    .get_position = d_get_position,
    .set_window_size = d_set_window_size,
    .construct_camera = d_construct_camera,
    .describe = d_describe,
};

static struct vox_module doom_camera_module =
{
    .type = CAMERA_MODULE,
    .methods = (struct vox_module_methods*) &doom_camera_interface
};

struct vox_module* module_init (void)
{
    return &doom_camera_module;
}
