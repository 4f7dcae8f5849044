/**
   @file tree.h
   @brief Tree construction (reference: src/voxtrees/tree.h).

   `struct vox_node` is opaque, as it is for users of the reference
   (tree.h:39-48).  Behind it the tree is a flattened, breadth-first node
   array mirrored into GPU memory on first use.
**/
#ifndef VOXTREES_TREE_H
#define VOXTREES_TREE_H

#include "params.h"

struct vox_node;

/**
   \brief Turn a set of voxels into a tree (tree.h:60, tree.c:249-338).

   The set is destructively reordered exactly as the reference reorders it and
   can be freed afterwards.  Returns NULL for n == 0.  Runs on the host; the
   resulting node array is bit-identical to a breadth-first flattening of the
   reference's tree for the same input.
**/
VOX_EXPORT struct vox_node* vox_make_tree (vox_dot set[], size_t n);

/** \brief Free a tree and its device copies (tree.h:67). Pass the root. **/
VOX_EXPORT void vox_destroy_tree (struct vox_node *tree);

/** \brief Number of voxels in the (sub)tree (tree.h:72). **/
VOX_EXPORT size_t vox_voxels_in_tree (const struct vox_node *tree);

/** \brief Bounding box of the (sub)tree (tree.h:77). **/
VOX_EXPORT void vox_bounding_box (const struct vox_node *tree, struct vox_box *box);

/**
   \brief Rebuild a tree from its voxels (tree.h:78-86, tree.c:366-435).

   Returns a NEW, freshly balanced tree; the old one stays valid and must still
   be destroyed by the caller.
**/
VOX_EXPORT struct vox_node* vox_rebuild_tree (const struct vox_node *tree);

/**
   \brief Insert a voxel on the fly (tree.h:88-96, tree.c:437-621).

   `*tree_ptr` is the root (or NULL for an empty tree) and is updated.  `voxel`
   is aligned to the voxel grid in place.  The edited tree has the shape the
   reference's tree has after the same edit.  The GPU copy of the tree is
   refreshed lazily: the next query or vox_render() after a burst of edits
   re-flattens the tree and uploads it once.  As in the reference, edits must
   not overlap searches or renders of the same tree, and node pointers returned
   by earlier searches are dead after an edit.

   \return 1 on success, 0 if the voxel was already in the tree.
**/
VOX_EXPORT int vox_insert_voxel (struct vox_node **tree_ptr, vox_dot voxel);
/** \brief vox_insert_voxel() by coordinates (tree.h:98-104, tree.c:623-633). **/
VOX_EXPORT int vox_insert_voxel_coord (struct vox_node **tree_ptr, float x, float y, float z);

/**
   \brief Delete a voxel on the fly (tree.h:106-114, tree.c:635-855).

   When the last voxel goes, the tree is destroyed and `*tree_ptr` becomes NULL.
   \return 1 on success, 0 if there was no such voxel in the tree.
**/
VOX_EXPORT int vox_delete_voxel (struct vox_node **tree_ptr, vox_dot voxel);
/** \brief vox_delete_voxel() by coordinates (tree.h:116-122, tree.c:857-866). **/
VOX_EXPORT int vox_delete_voxel_coord (struct vox_node **tree_ptr, float x, float y, float z);

/** \brief Print a (sub)tree to stdout; node ids are breadth-first positions (tree.h:124-130, tree.c:868-918). **/
VOX_EXPORT void vox_dump_tree (const struct vox_node *tree);

/** \brief A tree that is one solid box (tree.h:135, tree.c:229-240). **/
VOX_EXPORT struct vox_node* vox_make_dense_leaf (const struct vox_box *box);

/** \brief Set the global voxel size (tree.h:143-149, tree.c:920-925). **/
VOX_EXPORT void vox_set_voxel_size (float x, float y, float z);

/* ------------------------------------------------------------------------
 * Additions (not in the reference): inspection / import of the flattened
 * layout.  They exist so that parity tests can compare this builder with
 * the reference's and run the device traversal on a tree the reference built.
 * ---------------------------------------------------------------------- */

/** Sizes of the flattened tree behind `tree` (whole tree, not the subtree). */
struct vox_tree_info
{
    uint64_t n_nodes, n_inner, n_leaf, n_dense, n_overflow, n_dots;
    uint32_t depth;     /**< root alone = 1 */
    uint32_t node_id;   /**< breadth-first id of the node `tree` points at */
};
VOX_EXPORT void vox_tree_get_info (const struct vox_node *tree, struct vox_tree_info *info);

/**
   Flattened arrays in the canonical packed layout (3 floats per point),
   breadth-first, non-NULL children of a node consecutive in subspace order.
   Used both for export (caller allocates from vox_tree_info) and import.
**/
struct vox_tree_arrays
{
    float    *bb_min;      /* 3*n_nodes */
    float    *bb_max;      /* 3*n_nodes */
    uint32_t *flags;       /* n_nodes: 0 inner, 1 leaf, 2 dense leaf, 5 overflowed leaf (tree.h:15-18) */
    uint32_t *dots_num;    /* n_nodes */
    float    *center;      /* 3*n_nodes */
    uint32_t *first_child; /* n_nodes */
    uint32_t *child_mask;  /* n_nodes */
    uint64_t *dots_off;    /* n_nodes */
    float    *dots;        /* 3*n_dots */
};
VOX_EXPORT void vox_tree_export (const struct vox_node *tree, struct vox_tree_arrays *out);
VOX_EXPORT struct vox_node* vox_tree_import (const struct vox_tree_arrays *in,
                                             uint64_t n_nodes, uint64_t n_dots);

/** Node handle by breadth-first id; usable wherever a `tree` is accepted. NULL if out of range. */
VOX_EXPORT const struct vox_node* vox_tree_node (const struct vox_node *tree, uint32_t id);

#endif
