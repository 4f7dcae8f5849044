/**
   @file object.h
   @brief Common object definitions (reference: src/voxrnd/object.h).
**/
#ifndef VOXRND_OBJECT_H
#define VOXRND_OBJECT_H

#include "../voxvision.h"

/** The two generic property setters every voxrnd object starts with (object.h:29-30). */
#define VOX_OBJECT(type) void (*set_property_number) (struct type *obj, const char *name, float value); \
    void (*set_property_dot) (struct type *obj, const char *name, const vox_dot value);

struct vox_object
{
    struct vox_object_interface *iface;
};

struct vox_object_interface {
    VOX_OBJECT(vox_object)
};

#endif
