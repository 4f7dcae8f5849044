/**
   @file params.h
   @brief Global parameters used by the library (reference: src/voxtrees/params.h).
**/
#ifndef VOXTREES_PARAMS_H
#define VOXTREES_PARAMS_H

#include "../voxvision.h"

/**
   \brief Sides of voxel (params.h:38, defined {1,1,1} in tree.c:11).

   A mutable exported variable, read when a tree is built and captured into the
   device copy of that tree when it is uploaded.  As in the reference
   (tree.h:143-149) all trees must be rebuilt after changing it.
**/
VOX_EXPORT extern vox_dot vox_voxel;

#endif
