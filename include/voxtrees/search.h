/**
   @file search.h
   @brief Searching in a tree (reference: src/voxtrees/search.h).

   Both calls run on the GPU (a batch of one); there is no CPU path.  They
   return NULL / 0 and set vox_last_error() if no CUDA device is usable.
   For throughput use the batched forms in batch.h.
**/
#ifndef VOXTREES_SEARCH_H
#define VOXTREES_SEARCH_H

#include "tree.h"

/**
   \brief Find intersection of a tree and a ray (search.h:24-26, search.c:8-153).

   \return leaf where the intersection is found or NULL.  The returned node can
   itself be passed as `tree` (the renderer's FAST mode does, renderer.c:297).
   `res` is written only on a hit.
**/
VOX_EXPORT const struct vox_node*
vox_ray_tree_intersection (const struct vox_node* tree, const vox_dot origin,
                           const vox_dot dir, vox_dot res);

/**
   \brief Does a ball collide with voxels of a tree (search.h:36-37, search.c:155-183).
   \return 1 if a collision was found, 0 otherwise
**/
VOX_EXPORT int vox_tree_ball_collidep (const struct vox_node* tree, const vox_dot center,
                                       float radius);

#endif
