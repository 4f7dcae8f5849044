/*
 * build.h -- the recursive builder's intermediate form, shared by tree.c
 * (vox_make_tree) and edit.c (subtrees created by insertions / deletions).
 */
#ifndef VOXTREES_BUILD_H
#define VOXTREES_BUILD_H

#include "store.h"

/* A node as the recursion produces it, before breadth-first numbering. */
struct bnode
{
    float bb_min[3], bb_max[3], center[3];
    uint32_t flags, dots_num;
    uint32_t child[VOX_NS];     /* pool index or NO_NODE */
    size_t set_off;             /* leaves: where the dots sit in the reordered set */
};
#define NO_NODE 0xffffffffu

/* Chunked, append-only pool shared by the builder tasks. */
#define POOL_SHIFT 16
#define POOL_CHUNK (1u << POOL_SHIFT)
#define POOL_MAX_CHUNKS (1u << (31 - POOL_SHIFT))
struct pool
{
    struct bnode *chunk[POOL_MAX_CHUNKS];
    uint32_t used;
    int failed;
};


static inline struct bnode* pool_at (struct pool *p, uint32_t idx)
{
    return &p->chunk[idx >> POOL_SHIFT][idx & (POOL_CHUNK - 1)];
}

uint32_t vox_pool_new (struct pool *p);
void vox_pool_free (struct pool *p);
/* vox_make_tree's recursion (tree.c:249-338) over base[off, off+n); returns the pool index of the
   subtree's root or NO_NODE.  May spawn OpenMP tasks when called inside a parallel region. */
uint32_t vox_build (struct pool *pool, vox_dot *base, size_t off, size_t n, const float voxel[3]);

#endif
