/*
 * simple-camera.c -- 6-DoF quaternion camera module (reference:
 * src/voxrnd/simple-camera.c).  Built as simple-camera.so and loaded through
 * vox_camera_methods("simple-camera").  Adds the `describe` slot so that
 * screen2world runs inside the render kernel.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "private.h"

struct simple_camera
{
    struct vox_camera_interface *iface;
    float xsub, ysub;
    vox_dot position;
    vox_quat rotation;
    float mul, fov;
};

VOX_EXPORT struct vox_module* module_init (void);

static void s_screen2world (const struct vox_camera *cam, vox_dot ray, int sx, int sy)
{
    /* simple-camera.c:21-31 */
    const struct simple_camera *c = (const void*) cam;
    vox_dot v;
    vox_dot_set (v, c->mul*sx - c->xsub, 1, c->ysub - c->mul*sy);
    vox_rotate_vector (c->rotation, v, ray);
}

static void s_get_position (const struct vox_camera *cam, vox_dot res)
{
    const struct simple_camera *c = (const void*) cam;
    vox_dot_copy (res, c->position);
}

static void s_move_camera (struct vox_camera *cam, const vox_dot delta)
{
    /* simple-camera.c:39-46 */
    struct simple_camera *c = (void*) cam;
    vox_dot tr;
    vox_rotate_vector (c->rotation, delta, tr);
    for (int i=0; i<3; i++) c->position[i] = c->position[i] + tr[i];
}

static void s_rotate_camera (struct vox_camera *cam, const vox_dot delta)
{
    /* simple-camera.c:49-103: about camera Z, then Y, then X; then renormalise */
    struct simple_camera *c = (void*) cam;
    static const vox_quat orts[3] = { {0,1,0,0}, {0,0,1,0}, {0,0,0,1} };
    for (int i=2; i>=0; i--)
    {
        vox_quat ort;
        vox_rotate_vector_q (c->rotation, orts[i], ort);
        float sinang = sinf (delta[i]/2);
        float cosang = cosf (delta[i]/2);
        for (int j=0; j<3; j++) ort[j+1] = sinang*ort[j+1];
        ort[0] = cosang;
        vox_quat_mul (ort, c->rotation, c->rotation);
    }
    vox_quat_normalize (c->rotation);
}

static void s_look_at (struct vox_camera *cam, const vox_dot coord)
{
    /* simple-camera.c:105-118 */
    struct simple_camera *c = (void*) cam;
    vox_dot sub, rot;
    for (int i=0; i<3; i++) sub[i] = coord[i] - c->position[i];
    vox_quat_set_identity (c->rotation);
    vox_dot_set (rot, atan2f (sub[2], hypotf (sub[0], sub[1])), 0, -atan2f (sub[0], sub[1]));
    cam->iface->rotate_camera (cam, rot);
}

static void s_set_window_size (struct vox_camera *cam, int w, int h)
{
    /* simple-camera.c:120-136 */
    struct simple_camera *c = (void*) cam;
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

static struct vox_camera* s_construct_camera (const struct vox_camera *cam)
{
    /* simple-camera.c:138-158 */
    const struct simple_camera *old = (const void*) cam;
    struct simple_camera *c = vox_alloc (sizeof *c);
    if (old) memcpy (c, old, sizeof *c);
    else
    {
        memset (c, 0, sizeof *c);
        c->fov = 1.0;
        vox_quat_set_identity (c->rotation);
    }
    /* every camera owns its method table (the reference shares it between a
       camera and its copies and frees it twice on destruction) */
    vox_init_camera ((struct vox_camera*) c);
    if (old) vox_use_camera_methods ((struct vox_camera*) c, old->iface);
    vox_use_camera_methods ((struct vox_camera*) c, (struct vox_camera_interface*) module_init ()->methods);
    return (struct vox_camera*) c;
}

static void s_set_property_number (struct vox_camera *cam, const char *name, float value)
{
    struct simple_camera *c = (void*) cam;
    if (strcmp (name, "fov") == 0) c->fov = value;
}

static void s_set_property_dot (struct vox_camera *cam, const char *name, const vox_dot value)
{
    /* simple-camera.c:166-185 */
    struct simple_camera *c = (void*) cam;
    if (strcmp (name, "position") == 0) vox_dot_copy (c->position, value);
    else if (strcmp (name, "rotation") == 0)
    {
        vox_quat r0, r1, r2;
        vox_quat_set (r0, cosf (value[0]/2), sinf (value[0]/2), 0, 0);
        vox_quat_set (r1, cosf (value[1]/2), 0, sinf (value[1]/2), 0);
        vox_quat_set (r2, cosf (value[2]/2), 0, 0, sinf (value[2]/2));
        vox_quat_mul (r1, r0, c->rotation);
        vox_quat_mul (r2, c->rotation, c->rotation);
    }
}

static int s_describe (const struct vox_camera *cam, struct vox_camera_desc *desc)
{
    const struct simple_camera *c = (const void*) cam;
    memset (desc, 0, sizeof *desc);
    desc->kind = VOX_CAMERA_DESC_SIMPLE;
    memcpy (desc->position, c->position, 3*sizeof (float));
    memcpy (desc->rotation, c->rotation, 4*sizeof (float));
    desc->mul = c->mul; desc->xsub = c->xsub; desc->ysub = c->ysub;
    return 1;
}

static struct vox_camera_interface simple_camera_interface =
{
    .set_property_number = s_set_property_number,
    .set_property_dot = s_set_property_dot,
    .screen2world = s_screen2world,
    .rotate_camera = s_rotate_camera,
    .move_camera = s_move_camera,
    .look_at = s_look_at,
    .get_position = s_get_position,
    .set_window_size = s_set_window_size,
    .construct_camera = s_construct_camera,
    .describe = s_describe,
};

static struct vox_module simple_camera_module =
{
    .type = CAMERA_MODULE,
    .methods = (struct vox_module_methods*) &simple_camera_interface
};

struct vox_module* module_init (void)
{
    return &simple_camera_module;
}
