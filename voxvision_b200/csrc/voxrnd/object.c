/* object.c -- dummy object methods (reference: src/voxrnd/object.c:4-9). */
#include "private.h"

void vox_method_void_dummy (void *obj, ...) { (void) obj; }
void vox_method_dot_dummy (void *obj, vox_dot dot, ...)
{
    (void) obj;
    vox_dot_set (dot, 0, 0, 0);
}
