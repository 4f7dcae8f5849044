/**
   @file modules.h
   @brief Generic module loading (reference: src/voxrnd/modules.h).
**/
#ifndef VOXRND_MODULES_H
#define VOXRND_MODULES_H
#include "../voxvision.h"

#define CAMERA_MODULE 0

struct vox_module_methods;

/** What a module's `module_init()` returns (modules.h:26-36). */
struct vox_module {
    int type;
    struct vox_module_methods *methods;
};

/**
   \brief Load `<name>.so` and return its methods if its class is `type`
   (modules.h:54, modules.c:43-69).

   Search order: the directory named by the `VOXVISION_MODULES` environment
   variable, then the directory this library was loaded from (this replaces
   the reference's compile-time VOX_MODULE_PATH, voxvision.h.in:9).
**/
VOX_EXPORT struct vox_module_methods* vox_load_module (const char *name, int type);

#endif
