/* error.c -- per-thread error text behind vox_last_error() (include/voxvision.h). */
#include <stdio.h>
#include <stdarg.h>
#include "store.h"

static __thread char vox_error_text[512];

const char* vox_last_error (void) { return vox_error_text; }

void vox_set_error (const char *fmt, ...)
{
    va_list ap;
    va_start (ap, fmt);
    vsnprintf (vox_error_text, sizeof vox_error_text, fmt, ap);
    va_end (ap);
}
