/*
 * geom.c -- the three exported metrics of the reference's geom.h:59-73
 * (scalar definitions, src/voxtrees/geom.c:125-147).  Host-side helpers for
 * user code; the device has its own copies next to the traversal.
 */
#include <math.h>
#include "store.h"

float vox_abs_metric (const vox_dot a, const vox_dot b)
{
    float r = 0;
    for (int i=0; i<3; i++) r += fabsf (a[i] - b[i]);
    return r;
}

float vox_sqr_metric (const vox_dot a, const vox_dot b)
{
    float r = 0;
    for (int i=0; i<3; i++) { float d = a[i] - b[i]; r += d*d; }
    return r;
}

float vox_sqr_norm (const vox_dot a)
{
    float r = 0;
    for (int i=0; i<3; i++) r += a[i]*a[i];
    return r;
}
