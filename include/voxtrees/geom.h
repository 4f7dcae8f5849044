/**
   @file geom.h
   @brief Exported geometry helpers (reference: src/voxtrees/geom.h:59-73).
**/
#ifndef VOXTREES_GEOM_H
#define VOXTREES_GEOM_H

#include "params.h"

#define vox_dot_add(a,b,res) do { \
    (res)[0] = (a)[0] + (b)[0]; \
    (res)[1] = (a)[1] + (b)[1]; \
    (res)[2] = (a)[2] + (b)[2]; \
    } while (0)
#define vox_dot_sub(a,b,res) do { \
    (res)[0] = (a)[0] - (b)[0]; \
    (res)[1] = (a)[1] - (b)[1]; \
    (res)[2] = (a)[2] - (b)[2]; \
    } while (0)
#define vox_dot_scmul(d,sc,res) do { \
    (res)[0] = (sc) * (d)[0]; \
    (res)[1] = (sc) * (d)[1]; \
    (res)[2] = (sc) * (d)[2]; \
    } while (0)

/** \brief L1 metric, (|dx|+|dy|)+|dz| (geom.c:125-131). **/
VOX_EXPORT float vox_abs_metric (const vox_dot dot1, const vox_dot dot2);
/** \brief Squared Euclidean metric (geom.c:133-139). **/
VOX_EXPORT float vox_sqr_metric (const vox_dot dot1, const vox_dot dot2);
/** \brief Squared Euclidean norm (geom.c:141-147). **/
VOX_EXPORT float vox_sqr_norm (const vox_dot dot);

#endif
