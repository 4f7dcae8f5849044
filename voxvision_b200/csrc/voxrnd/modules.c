/* modules.c -- dlopen-based module loader (reference: src/voxrnd/modules.c:43-69). */
#define _GNU_SOURCE
#include <sys/param.h>
#include <sys/stat.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "private.h"

static int is_regular_file (const char *path)
{
    struct stat sb;
    return !stat (path, &sb) && S_ISREG (sb.st_mode);
}

/* $VOXVISION_MODULES first, then the directory libvoxrnd.so itself sits in. */
static int find_module (const char *name, char *fullpath)
{
    const char *env = getenv ("VOXVISION_MODULES");
    if (env && snprintf (fullpath, MAXPATHLEN, "%s/%s.so", env, name) < MAXPATHLEN &&
        is_regular_file (fullpath)) return 1;

    Dl_info info;
    if (dladdr ((void*) find_module, &info) && info.dli_fname)
    {
        char dir[MAXPATHLEN];
        strncpy (dir, info.dli_fname, MAXPATHLEN - 1);
        dir[MAXPATHLEN - 1] = 0;
        char *slash = strrchr (dir, '/');
        if (slash)
        {
            *slash = 0;
            if (snprintf (fullpath, MAXPATHLEN, "%s/%s.so", dir, name) < MAXPATHLEN &&
                is_regular_file (fullpath)) return 1;
        }
    }
    return 0;
}

struct vox_module_methods* vox_load_module (const char *name, int type)
{
    char fullpath[MAXPATHLEN];
    if (!find_module (name, fullpath))
    {
        vox_set_error ("module %s.so not found (set VOXVISION_MODULES)", name);
        return NULL;
    }
    /* RTLD_NODELETE keeps the method table valid after dlclose (modules.c:50-55) */
    void *handle = dlopen (fullpath, RTLD_LAZY | RTLD_NODELETE);
    if (!handle) { vox_set_error ("dlopen %s: %s", fullpath, dlerror ()); return NULL; }
    struct vox_module* (*init) (void) = (struct vox_module* (*) (void)) dlsym (handle, "module_init");
    dlclose (handle);
    if (!init) { vox_set_error ("%s has no module_init()", fullpath); return NULL; }
    struct vox_module *module = init ();
    if (module->type != type) { vox_set_error ("%s is not a module of class %d", fullpath, type); return NULL; }
    return module->methods;
}
