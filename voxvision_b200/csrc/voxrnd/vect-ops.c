/*
 * vect-ops.c -- quaternion helpers used by the cameras (reference, scalar
 * variant: src/voxrnd/vect-ops.c:8-75).  Operation order is the reference's,
 * so the camera state handed to the device matches the scalar oracle's.
 */
#include <math.h>
#include "private.h"

static void cross (const float *a, const float *b, float *r)
{
    /* vect-ops.c:8-16 */
    float ax = a[0], ay = a[1], az = a[2];
    float bx = b[0], by = b[1], bz = b[2];
    r[0] =  ay*bz - az*by;
    r[1] = -ax*bz + az*bx;
    r[2] =  ax*by - ay*bx;
}

void vox_quat_mul (const vox_quat q1, const vox_quat q2, vox_quat res)
{
    /* vect-ops.c:18-26 */
    float e1 = q1[0], i1 = q1[1], j1 = q1[2], k1 = q1[3];
    float e2 = q2[0], i2 = q2[1], j2 = q2[2], k2 = q2[3];
    float w = e1*e2 - i1*i2 - j1*j2 - k1*k2;
    float x = j1*k2 - k1*j2 + i1*e2 + e1*i2;
    float y = k1*i2 - i1*k2 + j1*e2 + e1*j2;
    float z = i1*j2 - j1*i2 + k1*e2 + e1*k2;
    vox_quat_set (res, w, x, y, z);
}

static void rotate3 (const float *q, const float *v, float *res)
{
    /* vect-ops.c:28-36: t = 2 (q.ijk x v); res = (v + q.w t) + (q.ijk x t) */
    float in[3] = { v[0], v[1], v[2] }, t[3], t2[3];
    cross (q+1, in, t);
    for (int i=0; i<3; i++) t[i] *= 2.0;
    cross (q+1, t, t2);
    for (int i=0; i<3; i++) res[i] = in[i] + q[0]*t[i] + t2[i];
}

void vox_rotate_vector (const vox_quat base, const vox_dot vector, vox_dot res)
{
    rotate3 (base, vector, res);
}

void vox_rotate_vector_q (const vox_quat base, const vox_quat vector, vox_quat res)
{
    /* vect-ops.c:38-46 */
    rotate3 (base, vector+1, res+1);
}

void vox_quat_normalize (vox_quat quat)
{
    /* vect-ops.c:48-57 */
    float sum = 0;
    for (int i=0; i<4; i++) sum += quat[i]*quat[i];
    float norm = sqrtf (sum);
    for (int i=0; i<4; i++) quat[i] /= norm;
}

void vox_dot_normalize (vox_dot dot)
{
    /* vect-ops.c:59-68 */
    float sum = 0;
    for (int i=0; i<3; i++) sum += dot[i]*dot[i];
    float norm = sqrtf (sum);
    for (int i=0; i<3; i++) dot[i] /= norm;
}

void vox_quat_set_identity (vox_quat quat)
{
    vox_quat_set (quat, 1, 0, 0, 0);
}
