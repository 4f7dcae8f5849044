/* camera.c -- generic camera plumbing (reference: src/voxrnd/camera.c:6-48). */
#include <stdlib.h>
#include <string.h>
#include "private.h"

void vox_use_camera_methods (struct vox_camera *camera, const struct vox_camera_interface *iface)
{
    /* camera.c:6-17: slot-wise overlay of the non-NULL methods */
    size_t nfields = sizeof (struct vox_camera_interface) / sizeof (void*);
    void * const *src = (void * const *) iface;
    void **dst = (void**) camera->iface;
    for (size_t i=0; i<nfields; i++)
        if (src[i] != NULL) dst[i] = src[i];
}

static void generic_destroy_camera (struct vox_camera *camera)
{
    free (camera->iface);
    free (camera);
}

void vox_init_camera (struct vox_camera *camera)
{
    /* camera.c:25-41: every slot gets a harmless default; `describe` stays NULL
       so an undescribed camera is rendered through its host screen2world */
    struct vox_camera_interface *iface = calloc (1, sizeof *iface);
    iface->set_property_number = (void*) vox_method_void_dummy;
    iface->set_property_dot = (void*) vox_method_void_dummy;
    iface->screen2world = (void*) vox_method_dot_dummy;
    iface->rotate_camera = (void*) vox_method_void_dummy;
    iface->move_camera = (void*) vox_method_void_dummy;
    iface->look_at = (void*) vox_method_void_dummy;
    iface->get_position = (void*) vox_method_dot_dummy;
    iface->set_window_size = (void*) vox_method_void_dummy;
    iface->destroy_camera = generic_destroy_camera;
    camera->iface = iface;
}

struct vox_camera_interface* vox_camera_methods (const char *name)
{
    return (struct vox_camera_interface*) vox_load_module (name, CAMERA_MODULE);
}
