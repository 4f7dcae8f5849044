/*
 * tree.c -- host-side octree construction into the flattened layout.
 *
 * Produces, for the same input array, exactly the tree the reference's
 * vox_make_tree builds (src/voxtrees/tree.c:249-338) -- same bounding boxes,
 * same split centres, same leaf contents in the same order, same reordering
 * of the caller's array -- but emits it as breadth-first arrays instead of
 * heap nodes, and builds disjoint subtrees in parallel (OpenMP tasks).
 *
 * What has to be reproduced exactly, and why:
 *   - the split centre is a SEQUENTIAL fp32 sum in array order
 *     (tree.c:125-136), so the order of the dots inside every sub-range
 *     matters;
 *   - that order is the result of the reference's swap-to-front partition,
 *     run once per subspace index 0..7 (tree.c:166-183, 278-308);
 *   - leaf dot order decides ties in the ray search (search.c:72).
 * What is free: bounding boxes are plain min/max (tree.c:138-155), the
 * subspace of a dot can be computed once per node instead of once per pass,
 * and passes for empty subspaces (and the last non-empty one) move nothing.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <omp.h>
#include "store.h"

vox_dot vox_voxel = {1.0f, 1.0f, 1.0f, 0.0f};   /* tree.c:11 */

void vox_set_voxel_size (float x, float y, float z)
{
    /* tree.c:920-925 */
    vox_dot_set (vox_voxel, x, y, z);
}

/* ------------------------------------------------------------------ store */

struct vox_tree_store* vox_store_alloc (uint64_t n_nodes, uint64_t n_dots)
{
    struct vox_tree_store *st = calloc (1, sizeof *st);
    if (!st) return NULL;
    size_t nn = n_nodes ? n_nodes : 1, nd = n_dots ? n_dots : 1;
    st->n_nodes = n_nodes; st->n_dots = n_dots;
    st->cap_nodes = nn; st->cap_dots = nd;
    st->bb_min = malloc (nn * 3 * sizeof (float));
    st->bb_max = malloc (nn * 3 * sizeof (float));
    st->center = calloc (nn * 3, sizeof (float));
    st->dots   = malloc (nd * 3 * sizeof (float));
    st->flags = malloc (nn * sizeof (uint32_t));
    st->dots_num = malloc (nn * sizeof (uint32_t));
    st->first_child = calloc (nn, sizeof (uint32_t));
    st->child_mask = calloc (nn, sizeof (uint32_t));
    st->dots_off = calloc (nn, sizeof (uint64_t));
    pthread_mutex_init (&st->lock, NULL);
    if (!st->bb_min || !st->bb_max || !st->center || !st->dots || !st->flags || !st->dots_num ||
        !st->first_child || !st->child_mask || !st->dots_off)
    {
        vox_set_error ("out of memory for a tree of %llu nodes", (unsigned long long) n_nodes);
        vox_store_free (st);
        return NULL;
    }
    memcpy (st->voxel, vox_voxel, sizeof st->voxel);
    return st;
}

/* The handle of a node: the root's lives in the store, the others in an array that is only made when
   somebody asks for one (vox_ray_tree_intersection returning a leaf, vox_tree_node). */
struct vox_node* vox_store_handle (struct vox_tree_store *st, uint64_t id)
{
    if (id == 0) return &st->root;
    if (!st->handles)
    {
        pthread_mutex_lock (&st->lock);
        if (!st->handles)
        {
            struct vox_node *h = malloc ((st->n_nodes ? st->n_nodes : 1) * sizeof *h);
            if (h)
                for (uint64_t i=0; i<st->n_nodes; i++) { h[i].store = st; h[i].id = (uint32_t) i; h[i].magic = VOX_NODE_MAGIC; }
            __atomic_store_n (&st->handles, h, __ATOMIC_RELEASE);
        }
        pthread_mutex_unlock (&st->lock);
        if (!st->handles) return NULL;
    }
    return &st->handles[id];
}

/* Root handle, per-kind counts and depth, from the flat arrays. */
void vox_store_finish (struct vox_tree_store *st)
{
    st->root.store = st; st->root.id = 0; st->root.magic = VOX_NODE_MAGIC;
    uint32_t *level = malloc ((st->n_nodes ? st->n_nodes : 1) * sizeof (uint32_t));
    st->n_inner = st->n_leaf = st->n_dense = st->n_overflow = 0;
    st->depth = 0;
    if (st->n_nodes) level[0] = 1;
    for (uint64_t id=0; id<st->n_nodes; id++)
    {
        if (level[id] > st->depth) st->depth = level[id];
        uint32_t fl = st->flags[id];
        if (fl & VOX_DENSE_LEAF) st->n_dense++;
        else if (fl & VOX_LEAF) { st->n_leaf++; if (fl & VOX_OVERFLOW) st->n_overflow++; }
        else
        {
            st->n_inner++;
            int nc = __builtin_popcount (st->child_mask[id]);
            for (int j=0; j<nc; j++) level[st->first_child[id] + j] = level[id] + 1;
        }
    }
    free (level);
}

void vox_store_free (struct vox_tree_store *st)
{
    if (!st) return;
    for (int d=0; d<VOX_MAX_DEVICES; d++) if (st->dev[d]) voxgpu_tree_free (st->dev[d]);
    free (st->bb_min); free (st->bb_max); free (st->center); free (st->dots);
    free (st->flags); free (st->dots_num); free (st->first_child); free (st->child_mask);
    free (st->dots_off);
    if (st->handles) for (uint64_t i=1; i<st->n_nodes; i++) st->handles[i].magic = 0;
    st->root.magic = 0;
    free (st->handles);
    vox_edit_free (st->edit);
    pthread_mutex_destroy (&st->lock);
    free (st);
}

int vox_node_valid (const struct vox_node *node)
{
    if (!node || node->magic != VOX_NODE_MAGIC || !node->store) return 0;
    /* every public entry point passes through here: bring the flat arrays (and, by dropping
       them, the device copies) up to date with pending edits first (edit.c) */
    if (node->store->dirty && !vox_store_sync (node->store)) return 0;
    return node->id < node->store->n_nodes;
}

/* ---------------------------------------------------------------- builder */

#include "build.h"

uint32_t vox_pool_new (struct pool *p)
{
    uint32_t idx;
#pragma omp atomic capture
    idx = p->used++;
    uint32_t c = idx >> POOL_SHIFT;
    if (c >= POOL_MAX_CHUNKS) { p->failed = 1; return NO_NODE; }
    struct bnode *chunk;
#pragma omp atomic read
    chunk = p->chunk[c];
    if (!chunk)
    {
#pragma omp critical (vox_pool_grow)
        {
            if (!p->chunk[c])
            {
                struct bnode *mem = malloc (sizeof (struct bnode) * POOL_CHUNK);
                if (!mem) p->failed = 1;
#pragma omp atomic write
                p->chunk[c] = mem;
            }
        }
        if (!p->chunk[c]) return NO_NODE;
    }
    return idx;
}


static inline void swap_dots (vox_dot a, vox_dot b)
{
    vox_dot t;
    memcpy (t, a, sizeof (vox_dot));
    memcpy (a, b, sizeof (vox_dot));
    memcpy (b, t, sizeof (vox_dot));
}

#define TASK_MIN_DOTS 32768

uint32_t vox_build (struct pool *pool, vox_dot *base, size_t off, size_t n, const float voxel[3])
{
    vox_dot *set = base + off;
    uint32_t self = vox_pool_new (pool);
    if (self == NO_NODE) return NO_NODE;
    struct bnode *node = pool_at (pool, self);
    memset (node, 0, sizeof *node);
    for (int i=0; i<VOX_NS; i++) node->child[i] = NO_NODE;
    node->dots_num = (uint32_t) n;
    node->set_off = off;

    /* tree.c:138-155: box = [min over dots, max over dots + voxel] */
    float mn[3] = { set[0][0], set[0][1], set[0][2] };
    float mx[3] = { set[0][0], set[0][1], set[0][2] };
    for (size_t i=1; i<n; i++)
        for (int j=0; j<3; j++)
        {
            float v = set[i][j];
            if (v < mn[j]) mn[j] = v;
            if (v > mx[j]) mx[j] = v;
        }
    for (int j=0; j<3; j++) { mx[j] = mx[j] + voxel[j]; node->bb_min[j] = mn[j]; node->bb_max[j] = mx[j]; }

    /* geom.c:157-167 dense_set_p: |n*vvol - bbvol| < vvol */
    {
        float bbvol = (mx[0] - mn[0]) * (mx[1] - mn[1]) * (mx[2] - mn[2]);
        float vvol = voxel[0] * voxel[1] * voxel[2];
        if (fabsf (n*vvol - bbvol) < vvol) { node->flags = VOX_DENSE_LEAF; return self; }
    }
    if (n <= VOX_MAX_DOTS) { node->flags = VOX_LEAF; return self; }

    /* tree.c:125-136 find_center: running fp32 sum in array order */
    float sum[3] = { 0, 0, 0 }, center[3];
    for (size_t i=0; i<n; i++)
        for (int j=0; j<3; j++)
        {
            float t = set[i][j] - mn[j];
            sum[j] = t + sum[j];
        }
    for (int j=0; j<3; j++)
    {
        center[j] = ceilf ((sum[j]/n + mn[j]) / voxel[j]) * voxel[j];
        node->center[j] = center[j];
    }

    /* subspace of every dot, once (geom.c:44-51: bit j <=> center[j] > dot[j]) */
    uint8_t *code = malloc (n);
    if (!code) { pool->failed = 1; return NO_NODE; }
    size_t hist[VOX_NS] = { 0 };
    for (size_t i=0; i<n; i++)
    {
        uint8_t c = (uint8_t) ((center[0] > set[i][0]) | ((center[1] > set[i][1]) << 1) |
                               ((center[2] > set[i][2]) << 2));
        code[i] = c;
        hist[c]++;
    }

    /* tree.c:281-305: nothing separates -> an overflowed leaf with all n dots */
    for (int idx=0; idx<VOX_NS; idx++)
        if (hist[idx] == n)
        {
            free (code);
            node->flags = VOX_LEAF | VOX_OVERFLOW;
            memset (node->center, 0, sizeof node->center);
            return self;
        }

    /* tree.c:166-183 sort_set, idx = 0..7: swap-to-front partition of [offset, n).
     * A pass only moves something if some later dot is in another subspace. */
    size_t offset = 0, ranges[VOX_NS][2];
    for (int idx=0; idx<VOX_NS; idx++)
    {
        size_t cnt = hist[idx];
        ranges[idx][0] = offset; ranges[idx][1] = cnt;
        if (cnt && offset + cnt < n)
        {
            size_t counter = offset;
            for (size_t i=offset; counter < offset + cnt; i++)
            {
                if (code[i] == idx)
                {
                    if (i != counter)
                    {
                        swap_dots (set[counter], set[i]);
                        code[i] = code[counter];
                        code[counter] = (uint8_t) idx;
                    }
                    counter++;
                }
            }
        }
        offset += cnt;
    }
    free (code);

    for (int idx=0; idx<VOX_NS; idx++)
    {
        size_t c_off = ranges[idx][0], c_n = ranges[idx][1];
        if (!c_n) continue;
        if (c_n >= TASK_MIN_DOTS)
        {
#pragma omp task default(none) firstprivate(pool, base, off, c_off, c_n, voxel, self, idx)
            {
                uint32_t ch = vox_build (pool, base, off + c_off, c_n, voxel);
                pool_at (pool, self)->child[idx] = ch;
                if (ch == NO_NODE) pool->failed = 1;
            }
        }
        else
        {
            uint32_t ch = vox_build (pool, base, off + c_off, c_n, voxel);
            pool_at (pool, self)->child[idx] = ch;
            if (ch == NO_NODE) pool->failed = 1;
        }
    }
#pragma omp taskwait
    return self;
}

void vox_pool_free (struct pool *pool)
{
    if (!pool) return;
    /* chunks are handed out in order: nothing lives beyond the one `used` reached */
    uint32_t last = pool->used >> POOL_SHIFT;
    if (last >= POOL_MAX_CHUNKS) last = POOL_MAX_CHUNKS - 1;
    for (uint32_t c=0; c<=last; c++) free (pool->chunk[c]);
    free (pool);
}

/* Breadth-first numbering of the pool into a store. */
static struct vox_tree_store* flatten_pool (struct pool *pool, uint32_t root, const vox_dot *set)
{
    uint64_t n_nodes = pool->used, n_dots = 0;
    for (uint32_t i=0; i<n_nodes; i++)
    {
        struct bnode *b = pool_at (pool, i);
        if ((b->flags & VOX_LEAF) && !(b->flags & VOX_DENSE_LEAF)) n_dots += b->dots_num;
    }
    struct vox_tree_store *st = vox_store_alloc (n_nodes, n_dots);
    uint32_t *order = malloc (n_nodes * sizeof (uint32_t));   /* bfs position -> pool index */
    if (!st || !order) { vox_store_free (st); free (order); return NULL; }

    uint64_t head = 0, tail = 0, dots = 0;
    order[tail++] = root;
    while (head < tail)
    {
        const struct bnode *b = pool_at (pool, order[head]);
        uint64_t id = head++;
        memcpy (st->bb_min + 3*id, b->bb_min, 3*sizeof (float));
        memcpy (st->bb_max + 3*id, b->bb_max, 3*sizeof (float));
        memcpy (st->center + 3*id, b->center, 3*sizeof (float));
        st->flags[id] = b->flags;
        st->dots_num[id] = b->dots_num;
        st->dots_off[id] = dots;
        if (b->flags & VOX_DENSE_LEAF) {}
        else if (b->flags & VOX_LEAF)
        {
            for (uint32_t k=0; k<b->dots_num; k++)
                memcpy (st->dots + 3*(dots + k), set[b->set_off + k], 3*sizeof (float));
            dots += b->dots_num;
        }
        else
        {
            uint32_t mask = 0;
            st->first_child[id] = (uint32_t) tail;
            for (int s=0; s<VOX_NS; s++)
                if (b->child[s] != NO_NODE) { mask |= 1u << s; order[tail++] = b->child[s]; }
            st->child_mask[id] = mask;
        }
    }
    free (order);
    vox_store_finish (st);
    return st;
}

struct vox_node* vox_make_tree (vox_dot set[], size_t n)
{
    if (n == 0) return NULL;   /* tree.c:251-254: an empty set is the NULL tree */
    if (n >= 0x7fffffffull) { vox_set_error ("vox_make_tree: %zu voxels exceed 31-bit indexing", n); return NULL; }

    struct pool *pool = calloc (1, sizeof *pool);
    if (!pool) { vox_set_error ("vox_make_tree: out of memory"); return NULL; }
    float voxel[3] = { vox_voxel[0], vox_voxel[1], vox_voxel[2] };
    uint32_t root = NO_NODE;
#pragma omp parallel default(none) shared(pool, set, n, voxel, root)
#pragma omp single nowait
    root = vox_build (pool, set, 0, n, voxel);

    struct vox_tree_store *st = NULL;
    if (root != NO_NODE && !pool->failed) st = flatten_pool (pool, root, (const vox_dot*) set);
    else vox_set_error ("vox_make_tree: out of memory while building");
    vox_pool_free (pool);
    return st ? &st->root : NULL;
}

struct vox_node* vox_make_dense_leaf (const struct vox_box *box)
{
    /* tree.c:229-240; dimensions as geom.c:192-196 get_dimensions */
    struct vox_tree_store *st = vox_store_alloc (1, 0);
    if (!st) return NULL;
    size_t dim[3];
    for (int i=0; i<3; i++)
    {
        st->bb_min[i] = box->min[i];
        st->bb_max[i] = box->max[i];
        dim[i] = (size_t) ((box->max[i] - box->min[i]) / vox_voxel[i]);
    }
    st->flags[0] = VOX_DENSE_LEAF;
    st->dots_num[0] = (uint32_t) (dim[0]*dim[1]*dim[2]);
    vox_store_finish (st);
    return &st->root;
}

void vox_destroy_tree (struct vox_node *tree)
{
    if (!tree) return;
    if (!vox_node_valid (tree)) { vox_set_error ("vox_destroy_tree: not a live tree"); return; }
    if (tree->id != 0) { vox_set_error ("vox_destroy_tree: only the root can be destroyed"); return; }
    vox_store_free (tree->store);
}

size_t vox_voxels_in_tree (const struct vox_node *tree)
{
    /* tree.c:340-343 */
    return vox_node_valid (tree) ? tree->store->dots_num[tree->id] : 0;
}

void vox_bounding_box (const struct vox_node *tree, struct vox_box *box)
{
    /* tree.c:357-360 */
    if (!vox_node_valid (tree)) return;
    const struct vox_tree_store *st = tree->store;
    vox_dot_set (box->min, st->bb_min[3*tree->id], st->bb_min[3*tree->id+1], st->bb_min[3*tree->id+2]);
    vox_dot_set (box->max, st->bb_max[3*tree->id], st->bb_max[3*tree->id+1], st->bb_max[3*tree->id+2]);
}

void vox_dump_tree (const struct vox_node *tree)
{
    /* tree.c:868-918: debugging aid; ids are breadth-first positions instead of addresses */
    if (!vox_node_valid (tree)) return;
    const struct vox_tree_store *st = tree->store;
    uint64_t id = tree->id;
    uint32_t fl = st->flags[id];
    const char *kind = (fl & VOX_LEAF) ? "LEAF" : (fl & VOX_DENSE_LEAF) ? "DENSE LEAF" : "INNER";
    printf ("Node %llu %s\n", (unsigned long long) id, kind);
    printf ("Bounding box min <%f, %f, %f>\n", st->bb_min[3*id], st->bb_min[3*id+1], st->bb_min[3*id+2]);
    printf ("Bounding box max <%f, %f, %f>\n", st->bb_max[3*id], st->bb_max[3*id+1], st->bb_max[3*id+2]);
    printf ("Number of voxels %u\n", st->dots_num[id]);
    if (fl & VOX_LEAF)
    {
        printf ("Dots:\n");
        for (uint32_t i=0; i<st->dots_num[id]; i++)
        {
            const float *d = st->dots + 3*(st->dots_off[id] + i);
            printf ("<%f, %f, %f>\n", d[0], d[1], d[2]);
        }
    }
    else if (!(fl & VOX_LEAF_MASK))
    {
        printf ("Center <%f, %f, %f>\n", st->center[3*id], st->center[3*id+1], st->center[3*id+2]);
        printf ("Children:\n");
        uint32_t next = st->first_child[id];
        for (int s=0; s<VOX_NS; s++)
        {
            if (st->child_mask[id] & (1u << s)) printf ("%u\n", next++);
            else printf ("(nil)\n");
        }
    }
    printf ("------Node ends here------\n");
    if (!(fl & VOX_LEAF_MASK))
    {
        int nc = __builtin_popcount (st->child_mask[id]);
        for (int j=0; j<nc; j++) vox_dump_tree (vox_store_handle ((struct vox_tree_store*) st, st->first_child[id] + j));
        printf ("=======Node and children end here=======\n");
    }
}

/* ------------------------------------------------- inspection and import */

void vox_tree_get_info (const struct vox_node *tree, struct vox_tree_info *info)
{
    memset (info, 0, sizeof *info);
    if (!vox_node_valid (tree)) return;
    const struct vox_tree_store *st = tree->store;
    info->n_nodes = st->n_nodes; info->n_inner = st->n_inner; info->n_leaf = st->n_leaf;
    info->n_dense = st->n_dense; info->n_overflow = st->n_overflow; info->n_dots = st->n_dots;
    info->depth = st->depth; info->node_id = tree->id;
}

void vox_tree_export (const struct vox_node *tree, struct vox_tree_arrays *out)
{
    if (!vox_node_valid (tree)) return;
    const struct vox_tree_store *st = tree->store;
    size_t n = st->n_nodes;
    memcpy (out->bb_min, st->bb_min, n*3*sizeof (float));
    memcpy (out->bb_max, st->bb_max, n*3*sizeof (float));
    memcpy (out->center, st->center, n*3*sizeof (float));
    memcpy (out->flags, st->flags, n*sizeof (uint32_t));
    memcpy (out->dots_num, st->dots_num, n*sizeof (uint32_t));
    memcpy (out->first_child, st->first_child, n*sizeof (uint32_t));
    memcpy (out->child_mask, st->child_mask, n*sizeof (uint32_t));
    memcpy (out->dots_off, st->dots_off, n*sizeof (uint64_t));
    memcpy (out->dots, st->dots, st->n_dots*3*sizeof (float));
}

struct vox_node* vox_tree_import (const struct vox_tree_arrays *in, uint64_t n_nodes, uint64_t n_dots)
{
    if (n_nodes == 0) return NULL;
    /* validate the links before trusting them on the device */
    uint64_t expect_child = 1;
    for (uint64_t id=0; id<n_nodes; id++)
    {
        uint32_t fl = in->flags[id];
        if (fl & VOX_DENSE_LEAF) continue;
        if (fl & VOX_LEAF)
        {
            if (in->dots_off[id] + in->dots_num[id] > n_dots)
            { vox_set_error ("vox_tree_import: leaf %llu points outside dots[]", (unsigned long long) id); return NULL; }
            continue;
        }
        /* mask == 0 is legal: deleting the last voxel of an octant leaves the reference's childless
           inner node behind (tree.c:695-756), and edited trees are exported / replicated like any other */
        uint32_t mask = in->child_mask[id];
        if (mask > 0xff || in->first_child[id] != expect_child)
        { vox_set_error ("vox_tree_import: node %llu breaks the breadth-first child layout", (unsigned long long) id); return NULL; }
        expect_child += __builtin_popcount (mask);
    }
    if (expect_child != n_nodes) { vox_set_error ("vox_tree_import: %llu nodes unreachable", (unsigned long long) (n_nodes - expect_child)); return NULL; }

    struct vox_tree_store *st = vox_store_alloc (n_nodes, n_dots);
    if (!st) return NULL;
    memcpy (st->bb_min, in->bb_min, n_nodes*3*sizeof (float));
    memcpy (st->bb_max, in->bb_max, n_nodes*3*sizeof (float));
    memcpy (st->center, in->center, n_nodes*3*sizeof (float));
    memcpy (st->flags, in->flags, n_nodes*sizeof (uint32_t));
    memcpy (st->dots_num, in->dots_num, n_nodes*sizeof (uint32_t));
    memcpy (st->first_child, in->first_child, n_nodes*sizeof (uint32_t));
    memcpy (st->child_mask, in->child_mask, n_nodes*sizeof (uint32_t));
    memcpy (st->dots_off, in->dots_off, n_nodes*sizeof (uint64_t));
    memcpy (st->dots, in->dots, n_dots*3*sizeof (float));
    vox_store_finish (st);
    return &st->root;
}

const struct vox_node* vox_tree_node (const struct vox_node *tree, uint32_t id)
{
    if (!vox_node_valid (tree) || id >= tree->store->n_nodes) return NULL;
    return vox_store_handle (tree->store, id);
}
