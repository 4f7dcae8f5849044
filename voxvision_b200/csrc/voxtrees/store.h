/*
 * store.h -- private layout behind the opaque `struct vox_node`.
 *
 * The reference links heap nodes by pointers (src/voxtrees/tree.h:22-38).
 * Here a tree is one `vox_tree_store`: the canonical flattened arrays
 * (breadth-first, children consecutive in subspace order) plus one small
 * handle per node, so that every node still has an address that can be
 * returned as a leaf and passed back in as a tree (renderer.c:297).
 */
#ifndef VOXTREES_STORE_H
#define VOXTREES_STORE_H

#include <pthread.h>
#include "voxtrees.h"
#include "voxgpu.h"

#define VOX_LEAF 1u
#define VOX_DENSE_LEAF 2u
#define VOX_LEAF_MASK 3u
#define VOX_OVERFLOW 4u
#define VOX_MAX_DOTS 7
#define VOX_NS 8

#define VOX_NODE_MAGIC 0x564f5854u /* "VOXT" */
#define VOX_MAX_DEVICES 64

struct vox_tree_store;

struct vox_node
{
    struct vox_tree_store *store;
    uint32_t id;
    uint32_t magic;
};

struct vox_tree_store
{
    uint64_t n_nodes, n_dots;
    uint64_t n_inner, n_leaf, n_dense, n_overflow;
    uint32_t depth;
    float voxel[3];

    float    *bb_min, *bb_max, *center, *dots;   /* packed 3 floats per entry */
    uint32_t *flags, *dots_num, *first_child, *child_mask;
    uint64_t *dots_off;

    struct vox_node root;                        /* handle of node 0: its address never changes */
    struct vox_node *handles;                    /* handles of nodes 1.. (entry 0 unused), made on first demand
                                                    (vox_store_handle): a renderer never asks for one; dropped by edits */
    uint64_t cap_nodes, cap_dots;                /* allocated entries of the arrays above */

    /* online edits (edit.c): a linked, mutable mirror of the tree; the flat arrays and the
       device copies are refreshed from it on the next use when `dirty` is set */
    struct vox_edit *edit;
    int dirty;

    pthread_mutex_t lock;
    voxgpu_tree *dev[VOX_MAX_DEVICES];           /* device copies, by CUDA ordinal */
};

/* error channel shared by libvoxtrees and libvoxrnd */
VOX_EXPORT void vox_set_error (const char *fmt, ...);

struct vox_tree_store* vox_store_alloc (uint64_t n_nodes, uint64_t n_dots);
void vox_store_finish (struct vox_tree_store *st);   /* handles + statistics */
void vox_store_free (struct vox_tree_store *st);
int vox_node_valid (const struct vox_node *node);
struct vox_node* vox_store_handle (struct vox_tree_store *st, uint64_t id);
/* edit.c */
void vox_edit_free (struct vox_edit *e);
int vox_store_sync (struct vox_tree_store *st);      /* flat arrays <- edited tree, if dirty; 0 on failure */
voxgpu_tree* vox_store_device (struct vox_tree_store *st);

#endif
