/**
   @file batch.h
   @brief Batched tree queries (new entry points; no reference counterpart).

   The reference only has one-ray / one-ball calls (search.h:24-37).  These run
   the same search for n independent queries in one CUDA launch.
**/
#ifndef VOXTREES_BATCH_H
#define VOXTREES_BATCH_H

#include "tree.h"

/**
   \brief vox_ray_tree_intersection() for n rays with HOST buffers.

   \param tree      root (or any node) to search from; NULL is an empty tree.
   \param leaf_ids  out, optional: breadth-first id of the hit leaf, -1 on a miss
                    (vox_tree_node() turns an id into a `struct vox_node*`).
   \param points    out, optional: hit points; untouched lanes on a miss.
   \param voxels    out, optional: index of the hit voxel among the voxels stored
                    in leaves (see vox_tree_arrays.dots), -1 for a dense leaf or miss.
   \return number of rays that hit, or -1 on error (see vox_last_error()).
**/
VOX_EXPORT long vox_ray_tree_intersection_batch (const struct vox_node *tree,
                                                 const vox_dot *origins, const vox_dot *dirs,
                                                 size_t n, int32_t *leaf_ids, vox_dot *points,
                                                 int32_t *voxels);

/**
   \brief vox_tree_ball_collidep() for n balls with HOST buffers.
   \param out  n bytes, 1 = collides.
   \return number of colliding balls, or -1 on error.
**/
VOX_EXPORT long vox_tree_ball_collidep_batch (const struct vox_node *tree,
                                              const vox_dot *centers, const float *radii,
                                              size_t n, uint8_t *out);

/**
   \brief Make sure the device copy of the tree exists on the current CUDA
   device (it is otherwise created lazily by the first query) and return the
   thin-layer handle (`voxgpu_tree*`, see voxgpu.h) for callers that drive the
   kernels with device buffers and their own streams.  NULL on error.
**/
VOX_EXPORT void* vox_tree_device_handle (const struct vox_node *tree);

#endif
