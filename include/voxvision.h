/**
   @file voxvision.h
   @brief Types and declarations shared by the voxvision-b200 libraries.

   Replaces the header CMake generates from the reference's
   src/voxvision.h.in:1-87.  The reference forks its ABI at configure time:
   `vox_dot` is `float[4]`, 16-byte aligned, when SSE_INTRIN is on (the default
   build, CMakeLists.txt:19) and `float[3]` otherwise (voxvision.h.in:45-59).
   This implementation fixes the default (SSE) ABI: `float[4]`, aligned 16,
   lane 3 unused.  It also matches the 16-byte vector loads the device uses.
**/
#ifndef VOXVISION_H
#define VOXVISION_H

#include <stdlib.h>
#include <string.h>
#include <stddef.h>
#include <stdint.h>

#define VOX_VERSION_MAJOR 0
#define VOX_VERSION_MINOR 34
#define VOX_VERSION "0.34-b200"
#define VOX_EXPORT __attribute__((visibility ("default")))

/* Same data layout as voxvision.h.in:23-42: x, y, z, (unused) / w, i, j, k */
typedef float vox_dot[4] __attribute__ ((aligned (16)));
typedef float vox_quat[4] __attribute__ ((aligned (16)));

#define vox_alloc(size) aligned_alloc (16, ((size) + 0xf) & ~(size_t)0xf)
#define vox_dot_copy(d1, d2) memcpy ((d1), (d2), sizeof (vox_dot))
#define vox_dot_equalp(d1, d2) (memcmp ((d1), (d2), 3*sizeof (float)) == 0)
#define vox_dot_set(dot, x, y, z) do {          \
        (dot)[0] = (x);                         \
        (dot)[1] = (y);                         \
        (dot)[2] = (z);                         \
        (dot)[3] = 0;                           \
    } while (0)
#define vox_quat_copy(q1, q2) memcpy ((q1), (q2), sizeof (vox_quat))
#define vox_quat_set(q, w, i, j, k) do {  \
        (q)[0] = (w);                     \
        (q)[1] = (i);                     \
        (q)[2] = (j);                     \
        (q)[3] = (k);                     \
    } while (0)

/**
   \brief Structure representing a box (voxvision.h.in:79-83).
**/
struct vox_box
{
    vox_dot min; /**< \brief minimal coordinate */
    vox_dot max; /**< \brief maximal coordinate */
};
#define vox_box_copy(b1, b2) memcpy ((b1), (b2), sizeof (struct vox_box))

/**
   \brief Error text of the last query / render call made on the calling thread
   (CUDA failures included), or "" if that call succeeded.

   The reference has no error channel besides NULL / 0 returns
   (SURVEY.md 8b "Error conventions"); those conventions are kept and this
   accessor is added so that a device failure is never silent.
**/
VOX_EXPORT const char* vox_last_error (void);

#endif
