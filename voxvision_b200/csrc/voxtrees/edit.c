/*
 * edit.c -- online tree edits and what they mean for the device copy
 * (SURVEY.md 8 f-3).
 *
 *   vox_insert_voxel[_coord]   reference: src/voxtrees/tree.c:437-633
 *   vox_delete_voxel[_coord]   reference: src/voxtrees/tree.c:635-866
 *   vox_rebuild_tree           reference: src/voxtrees/tree.c:366-435
 *
 * The searchable form of a tree here is a flat, breadth-first array set that is
 * mirrored in GPU memory; it has no room for a node to grow.  So edits go to a
 * linked, mutable mirror of the tree (`struct enode`) that is made LAZILY: it
 * starts as one stub that stands for "flat node 0 and everything below it", and a
 * stub turns into a real node (with stubs for its children) only when an edit has
 * to look inside it -- an edit touches one root-to-leaf path, so the first edit
 * of a 40 M-node tree costs a handful of small allocations, not a copy of the
 * tree.  Edits are applied to the mirror with the reference's rules -- the SHAPE of the edited
 * tree decides later search results (which leaf a ray ends in, tie order
 * inside leaves, never-shrunk bounding boxes), so the rules are followed to the
 * letter, quirks included -- and the store is only marked dirty.  The next
 * query / render / inspection call re-numbers the mirror breadth-first into NEW
 * flat arrays (vox_store_sync; untouched subtrees are copied from the old arrays
 * with their ids shifted), discards the mirror, drops the device copies, and the
 * usual lazy upload brings the new arrays to the GPU: a burst of edits between two
 * frames costs one re-flatten (O(nodes) of plain copying) and one upload.
 *
 * The root handle (`&store->root`) keeps its address across edits, so a context
 * whose scene is the root keeps rendering the edited tree, as it would with the
 * reference whenever the reference's root node survives the edit.  Handles of
 * other nodes die with the first edit after they were obtained (the reference
 * frees / reallocates nodes on edits as well).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "build.h"

struct enode
{
    float bb_min[3], bb_max[3];
    uint32_t flags, dots_num;
    float center[3];                 /* inner */
    struct enode *child[VOX_NS];     /* inner */
    float (*dots)[3];                /* leaf; room for max (VOX_MAX_DOTS, dots_num) */
    int64_t ref;                     /* >= 0: a STUB for flat node `ref` and its whole subtree: box, flags and
                                        dots_num are valid, centre / children / dots are still in the flat arrays */
};

struct vox_edit
{
    struct enode *root;
    float voxel[3];
    const struct vox_tree_store *st; /* the flat arrays the stubs point into */
};

/* ------------------------------------------------------------ small parts */

static struct enode* enode_new (uint32_t flags)
{
    struct enode *n = calloc (1, sizeof *n);
    if (n) { n->flags = flags; n->ref = -1; }
    return n;
}

/* a stub for flat node `id` */
static struct enode* enode_stub (const struct vox_tree_store *st, uint64_t id)
{
    struct enode *n = enode_new (st->flags[id]);
    if (!n) return NULL;
    memcpy (n->bb_min, st->bb_min + 3*id, sizeof n->bb_min);
    memcpy (n->bb_max, st->bb_max + 3*id, sizeof n->bb_max);
    n->dots_num = st->dots_num[id];
    n->ref = (int64_t) id;
    return n;
}

/* look inside a stub: it becomes a real node whose children are stubs */
static void realize (const struct vox_edit *e, struct enode *n)
{
    if (!n || n->ref < 0) return;
    const struct vox_tree_store *st = e->st;
    const uint64_t id = (uint64_t) n->ref;
    n->ref = -1;
    if (n->flags & VOX_DENSE_LEAF) return;
    if (n->flags & VOX_LEAF)
    {
        uint32_t cap = n->dots_num > VOX_MAX_DOTS ? n->dots_num : VOX_MAX_DOTS;
        n->dots = malloc (cap * sizeof *n->dots);
        if (!n->dots) { n->dots_num = 0; vox_set_error ("out of memory while editing a tree"); return; }
        memcpy (n->dots, st->dots + 3*st->dots_off[id], (size_t) n->dots_num * 3 * sizeof (float));
        return;
    }
    memcpy (n->center, st->center + 3*id, sizeof n->center);
    uint32_t mask = st->child_mask[id], next = st->first_child[id];
    for (int s=0; s<VOX_NS; s++) if (mask & (1u << s)) n->child[s] = enode_stub (st, next++);
}

static void enode_free_tree (struct enode *n)
{
    if (!n) return;
    if (!(n->flags & VOX_LEAF_MASK)) for (int i=0; i<VOX_NS; i++) enode_free_tree (n->child[i]);
    free (n->dots);
    free (n);
}

void vox_edit_free (struct vox_edit *e)
{
    if (!e) return;
    enode_free_tree (e->root);
    free (e);
}

static inline int same_dot (const float a[3], const float b[3])
{
    return memcmp (a, b, 3*sizeof (float)) == 0;     /* vox_dot_equalp, voxvision.h.in */
}

/* geom.c:44-51 */
static inline int subspace (const float center[3], const float dot[3])
{
    return (center[0] > dot[0]) | ((center[1] > dot[1]) << 1) | ((center[2] > dot[2]) << 2);
}

/* geom.c:5-11: min <= dot < max */
static inline int in_box (const struct enode *n, const float dot[3])
{
    for (int i=0; i<3; i++) if (dot[i] < n->bb_min[i] || dot[i] >= n->bb_max[i]) return 0;
    return 1;
}

/* tree.c:197-208 update_bounding_box: grow the box so that it holds the voxel at `dot` */
static void grow_box (float mn[3], float mx[3], const float dot[3], const float voxel[3])
{
    for (int i=0; i<3; i++)
    {
        float top = dot[i] + voxel[i];
        mn[i] = (dot[i] < mn[i]) ? dot[i] : mn[i];
        mx[i] = (top > mx[i]) ? top : mx[i];
    }
}

/* geom.c:157-167 */
static int dense_box (const float mn[3], const float mx[3], size_t n, const float voxel[3])
{
    float bbvol = (mx[0] - mn[0]) * (mx[1] - mn[1]) * (mx[2] - mn[2]);
    float vvol = voxel[0] * voxel[1] * voxel[2];
    return fabsf (n*vvol - bbvol) < vvol;
}

/* geom.c:192-196 */
static void box_dims (const float mn[3], const float mx[3], const float voxel[3], size_t dim[3])
{
    for (int i=0; i<3; i++) dim[i] = (size_t) ((mx[i] - mn[i]) / voxel[i]);
}

/* tree.c:229-240 */
static struct enode* dense_leaf (const float mn[3], const float mx[3], const float voxel[3])
{
    struct enode *n = enode_new (VOX_DENSE_LEAF);
    size_t dim[3];
    memcpy (n->bb_min, mn, sizeof n->bb_min);
    memcpy (n->bb_max, mx, sizeof n->bb_max);
    box_dims (mn, mx, voxel, dim);
    n->dots_num = (uint32_t) (dim[0]*dim[1]*dim[2]);
    return n;
}

/* geom.c:169-182: the octant `idx` of the box around `center`; 0 if it is flat.
   The result box is only meaningful when 1 is returned. */
static int octant (const struct enode *box, const float center[3], int idx, float mn[3], float mx[3])
{
    memcpy (mn, box->bb_min, 3*sizeof (float));
    memcpy (mx, box->bb_max, 3*sizeof (float));
    for (int i=0; i<3; i++)
    {
        if (idx & (1<<i)) mx[i] = center[i];
        else mn[i] = center[i];
        if (mn[i] == mx[i]) return 0;
    }
    return 1;
}

/* geom.c:199-210: exactly two unit dimensions; *axis = the long one */
static int is_stripe (const struct enode *n, const float voxel[3], int *axis)
{
    size_t dim[3];
    int count = 0;
    box_dims (n->bb_min, n->bb_max, voxel, dim);
    for (int i=0; i<3; i++)
    {
        if (dim[i] == 1) count++;
        else *axis = i;
    }
    return count == 2;
}

/* Voxels of a subtree in the reference's order (flatten_tree, tree.c:366-422): children 0..7 depth-first,
   leaf dots as stored, a dense box x-outer / z-inner by repeated addition. */
static size_t collect_box (const float mn[3], const float mx[3], const float voxel[3], vox_dot *out)
{
    size_t dim[3], count = 0;
    box_dims (mn, mx, voxel, dim);
    float cur[3];
    cur[0] = mn[0];
    for (size_t i=0; i<dim[0]; i++)
    {
        cur[1] = mn[1];
        for (size_t j=0; j<dim[1]; j++)
        {
            cur[2] = mn[2];
            for (size_t k=0; k<dim[2]; k++)
            {
                vox_dot_set (out[count], cur[0], cur[1], cur[2]);
                cur[2] += voxel[2];
                count++;
            }
            cur[1] += voxel[1];
        }
        cur[0] += voxel[0];
    }
    return count;
}

/* the same walk over the flat arrays (a subtree no edit has looked into) */
static size_t collect_flat (const struct vox_tree_store *st, uint64_t id, const float voxel[3], vox_dot *out)
{
    const uint32_t fl = st->flags[id];
    if (fl & VOX_LEAF)
    {
        const float *d = st->dots + 3*st->dots_off[id];
        for (uint32_t i=0; i<st->dots_num[id]; i++) vox_dot_set (out[i], d[3*i], d[3*i+1], d[3*i+2]);
        return st->dots_num[id];
    }
    if (fl & VOX_DENSE_LEAF) return collect_box (st->bb_min + 3*id, st->bb_max + 3*id, voxel, out);
    size_t count = 0;
    const int nc = __builtin_popcount (st->child_mask[id]);
    for (int j=0; j<nc; j++) count += collect_flat (st, st->first_child[id] + j, voxel, out + count);
    return count;
}

static size_t collect (const struct vox_edit *e, const struct enode *n, const float voxel[3], vox_dot *out)
{
    if (!n) return 0;
    if (n->ref >= 0) return collect_flat (e->st, (uint64_t) n->ref, voxel, out);
    if (n->flags & VOX_LEAF)
    {
        for (uint32_t i=0; i<n->dots_num; i++) vox_dot_set (out[i], n->dots[i][0], n->dots[i][1], n->dots[i][2]);
        return n->dots_num;
    }
    if (n->flags & VOX_DENSE_LEAF) return collect_box (n->bb_min, n->bb_max, voxel, out);
    size_t count = 0;
    for (int i=0; i<VOX_NS; i++) count += collect (e, n->child[i], voxel, out + count);
    return count;
}

/* ------------------------------------------- builder output -> linked form */

static struct enode* from_pool (struct pool *pool, uint32_t idx, const vox_dot *set)
{
    const struct bnode *b = pool_at (pool, idx);
    struct enode *n = enode_new (b->flags);
    if (!n) return NULL;
    memcpy (n->bb_min, b->bb_min, sizeof n->bb_min);
    memcpy (n->bb_max, b->bb_max, sizeof n->bb_max);
    n->dots_num = b->dots_num;
    if (b->flags & VOX_DENSE_LEAF) return n;
    if (b->flags & VOX_LEAF)
    {
        uint32_t cap = b->dots_num > VOX_MAX_DOTS ? b->dots_num : VOX_MAX_DOTS;
        n->dots = malloc (cap * sizeof *n->dots);
        if (!n->dots) { n->dots_num = 0; return n; }
        for (uint32_t k=0; k<b->dots_num; k++) memcpy (n->dots[k], set[b->set_off + k], 3*sizeof (float));
        return n;
    }
    memcpy (n->center, b->center, sizeof n->center);
    for (int s=0; s<VOX_NS; s++)
        if (b->child[s] != NO_NODE) n->child[s] = from_pool (pool, b->child[s], set);
    return n;
}

/* vox_make_tree (tree.c:249-338) of a handful of voxels, as a linked subtree. */
static struct enode* subtree_of (vox_dot *set, size_t n, const float voxel[3])
{
    if (!n) return NULL;
    struct pool *pool = calloc (1, sizeof *pool);
    if (!pool) return NULL;
    uint32_t root = vox_build (pool, set, 0, n, voxel);
    struct enode *res = (root != NO_NODE && !pool->failed) ? from_pool (pool, root, (const vox_dot*) set) : NULL;
    vox_pool_free (pool);
    return res;
}

/* flat arrays -> linked form: one stub for the root; everything else on demand */
static struct vox_edit* edit_from_store (const struct vox_tree_store *st)
{
    struct vox_edit *e = calloc (1, sizeof *e);
    if (!e) return NULL;
    memcpy (e->voxel, st->voxel, sizeof e->voxel);
    e->st = st;
    e->root = enode_stub (st, 0);
    if (!e->root) { free (e); return NULL; }
    return e;
}

/* ------------------------------------------------ linked form -> flat arrays */

/* One entry of the breadth-first queue: a real node of the mirror, or a node of the OLD flat arrays
   (inside a subtree no edit has looked into). */
struct qitem
{
    const struct enode *node; uint64_t old_id;
    uint64_t first;          /* pass 2: new id of the first child (inner nodes) */
    uint64_t doff;           /* pass 2: dots stored before this node (every node records it, as flatten_pool does) */
};

struct queue { struct qitem *it; uint64_t n, cap; };

static int push (struct queue *q, const struct enode *node, uint64_t old_id)
{
    if (q->n == q->cap)
    {
        uint64_t cap = q->cap ? 2*q->cap : 1024;
        struct qitem *it = realloc (q->it, cap * sizeof *it);
        if (!it) return 0;
        q->it = it; q->cap = cap;
    }
    q->it[q->n].node = node; q->it[q->n].old_id = old_id;
    q->n++;
    return 1;
}

int vox_store_sync (struct vox_tree_store *st)
{
    if (!st->dirty) return 1;
    pthread_mutex_lock (&st->lock);
    if (!st->dirty) { pthread_mutex_unlock (&st->lock); return 1; }
    const struct vox_edit *e = st->edit;

    /* breadth-first order of the edited tree, children consecutive in subspace order (as flatten_pool,
       tree.c); a stub contributes its flat subtree in that subtree's own (already breadth-first) child order */
    struct queue q = { NULL, 0, 0 };
    uint64_t n_dots = 0;
    q.cap = st->n_nodes + 4096;                      /* edits change the node count by little */
    q.it = malloc (q.cap * sizeof *q.it);
    if (!q.it) q.cap = 0;
    int ok = push (&q, e->root, 0);
    for (uint64_t head = 0; ok && head < q.n; head++)
    {
        const struct enode *n = q.it[head].node;
        if (n && n->ref < 0)
        {
            if (n->flags & VOX_DENSE_LEAF) continue;
            if (n->flags & VOX_LEAF) { n_dots += n->dots_num; continue; }
            for (int s=0; s<VOX_NS && ok; s++) if (n->child[s]) ok = push (&q, n->child[s], 0);
            continue;
        }
        const uint64_t id = n ? (uint64_t) n->ref : q.it[head].old_id;
        q.it[head].node = NULL; q.it[head].old_id = id;           /* from here on: an old flat node */
        const uint32_t fl = st->flags[id];
        if (fl & VOX_DENSE_LEAF) continue;
        if (fl & VOX_LEAF) { n_dots += st->dots_num[id]; continue; }
        const int nc = __builtin_popcount (st->child_mask[id]);
        for (int j=0; j<nc && ok; j++) ok = push (&q, NULL, st->first_child[id] + j);
    }
    const uint64_t n_nodes = q.n;

    /* new arrays (the old ones are still being read) */
    struct vox_tree_store *nw = ok ? vox_store_alloc (n_nodes + n_nodes/8 + 64, n_dots + n_dots/8 + 64) : NULL;
    if (!nw)
    {
        free (q.it);
        pthread_mutex_unlock (&st->lock);
        vox_set_error ("out of memory refreshing an edited tree");
        return 0;
    }
    /* where every node's children / dots go: a running count in queue order ... */
    {
        uint64_t next_child = 1, dots = 0;
        for (uint64_t id = 0; id < n_nodes; id++)
        {
            const struct enode *n = q.it[id].node;
            const uint32_t fl = n ? n->flags : st->flags[q.it[id].old_id];
            q.it[id].first = 0;
            q.it[id].doff = dots;
            if (fl & VOX_DENSE_LEAF) {}
            else if (fl & VOX_LEAF) dots += n ? n->dots_num : st->dots_num[q.it[id].old_id];
            else
            {
                q.it[id].first = next_child;
                if (n) { for (int s=0; s<VOX_NS; s++) next_child += n->child[s] != NULL; }
                else next_child += (uint64_t) __builtin_popcount (st->child_mask[q.it[id].old_id]);
            }
        }
    }
    /* ... then the copying itself is independent per node */
    #pragma omp parallel for schedule(static)
    for (uint64_t id = 0; id < n_nodes; id++)
    {
        const struct enode *n = q.it[id].node;
        const uint64_t first = q.it[id].first, doff = q.it[id].doff;
        nw->dots_off[id] = doff;
        nw->first_child[id] = 0;
        nw->child_mask[id] = 0;
        memset (nw->center + 3*id, 0, 3*sizeof (float));
        if (n)
        {
            memcpy (nw->bb_min + 3*id, n->bb_min, 3*sizeof (float));
            memcpy (nw->bb_max + 3*id, n->bb_max, 3*sizeof (float));
            nw->flags[id] = n->flags;
            nw->dots_num[id] = n->dots_num;
            if (n->flags & VOX_DENSE_LEAF) {}
            else if (n->flags & VOX_LEAF)
            {
                memcpy (nw->dots + 3*doff, n->dots, (size_t) n->dots_num * 3 * sizeof (float));
            }
            else
            {
                uint32_t mask = 0;
                memcpy (nw->center + 3*id, n->center, 3*sizeof (float));
                for (int s=0; s<VOX_NS; s++) if (n->child[s]) mask |= 1u << s;
                nw->first_child[id] = (uint32_t) first;
                nw->child_mask[id] = mask;
            }
            continue;
        }
        const uint64_t old = q.it[id].old_id;
        memcpy (nw->bb_min + 3*id, st->bb_min + 3*old, 3*sizeof (float));
        memcpy (nw->bb_max + 3*id, st->bb_max + 3*old, 3*sizeof (float));
        nw->flags[id] = st->flags[old];
        nw->dots_num[id] = st->dots_num[old];
        if (st->flags[old] & VOX_DENSE_LEAF) {}
        else if (st->flags[old] & VOX_LEAF)
        {
            memcpy (nw->dots + 3*doff, st->dots + 3*st->dots_off[old], (size_t) st->dots_num[old] * 3 * sizeof (float));
        }
        else
        {
            memcpy (nw->center + 3*id, st->center + 3*old, 3*sizeof (float));
            nw->first_child[id] = (uint32_t) first;
            nw->child_mask[id] = st->child_mask[old];
        }
    }
    free (q.it);

    /* the store keeps its address (and with it the root handle); it takes over the new arrays */
    if (st->handles) for (uint64_t i=1; i<st->n_nodes; i++) st->handles[i].magic = 0;   /* old handles die */
    free (st->bb_min); free (st->bb_max); free (st->center); free (st->dots);
    free (st->flags); free (st->dots_num); free (st->first_child); free (st->child_mask);
    free (st->dots_off); free (st->handles);
    st->bb_min = nw->bb_min; st->bb_max = nw->bb_max; st->center = nw->center; st->dots = nw->dots;
    st->flags = nw->flags; st->dots_num = nw->dots_num; st->first_child = nw->first_child;
    st->child_mask = nw->child_mask; st->dots_off = nw->dots_off; st->handles = NULL;
    st->cap_nodes = nw->cap_nodes; st->cap_dots = nw->cap_dots;
    st->n_nodes = n_nodes; st->n_dots = n_dots;
    pthread_mutex_destroy (&nw->lock);
    free (nw);
    vox_store_finish (st);

    /* the mirror pointed into the old arrays: the next edit starts a new one */
    vox_edit_free (st->edit);
    st->edit = NULL;

    /* the device copies describe the tree before the edits */
    for (int d=0; d<VOX_MAX_DEVICES; d++)
        if (st->dev[d]) { voxgpu_tree_free (st->dev[d]); st->dev[d] = NULL; }
    st->dirty = 0;
    pthread_mutex_unlock (&st->lock);
    return 1;
}

/* ------------------------------------------------------------- membership */

/* voxel_in_tree, tree.c:437-458 */
static int contains_flat (const struct vox_tree_store *st, uint64_t id, const float v[3])
{
    for (;;)
    {
        const float *mn = st->bb_min + 3*id, *mx = st->bb_max + 3*id;
        for (int i=0; i<3; i++) if (v[i] < mn[i] || v[i] >= mx[i]) return 0;
        const uint32_t fl = st->flags[id];
        if (fl & VOX_DENSE_LEAF) return 1;
        if (fl & VOX_LEAF)
        {
            const float *d = st->dots + 3*st->dots_off[id];
            for (uint32_t i=0; i<st->dots_num[id]; i++) if (same_dot (v, d + 3*i)) return 1;
            return 0;
        }
        const uint32_t sub = (uint32_t) subspace (st->center + 3*id, v), mask = st->child_mask[id];
        if (!(mask & (1u << sub))) return 0;
        id = st->first_child[id] + (uint32_t) __builtin_popcount (mask & ((1u << sub) - 1u));
    }
}

static int contains (const struct vox_edit *e, const struct enode *n, const float v[3])
{
    while (n && in_box (n, v))
    {
        if (n->ref >= 0) return contains_flat (e->st, (uint64_t) n->ref, v);
        if (n->flags & VOX_DENSE_LEAF) return 1;
        if (n->flags & VOX_LEAF)
        {
            for (uint32_t i=0; i<n->dots_num; i++) if (same_dot (v, n->dots[i])) return 1;
            return 0;
        }
        n = n->child[subspace (n->center, v)];
    }
    return 0;
}

/* -------------------------------------------------------------- insertion */

static struct enode* insert (const struct vox_edit *e, struct enode *t, const float v[3]);

/* insert_in_big_dense, tree.c:461-503: a solid box and one voxel outside it become the two children of a new node
   split at the box vertex nearest to the voxel */
static struct enode* wrap_dense (const struct vox_edit *e, struct enode *t, const float v[3])
{
    const float *voxel = e->voxel;
    struct enode *node = enode_new (0);
    memcpy (node->bb_min, t->bb_min, sizeof node->bb_min);
    memcpy (node->bb_max, t->bb_max, sizeof node->bb_max);
    grow_box (node->bb_min, node->bb_max, v, voxel);
    node->dots_num = t->dots_num + 1;
    float mid[3];
    for (int i=0; i<3; i++)
    {
        /* geom.c:33-42 closest_vertex */
        float d1 = fabsf (t->bb_min[i] - v[i]);
        float d2 = fabsf (t->bb_max[i] - v[i]);
        node->center[i] = (d1 < d2) ? t->bb_min[i] : t->bb_max[i];
        mid[i] = (t->bb_min[i] + t->bb_max[i]) / 2;
    }
    int at_box = subspace (node->center, mid), at_voxel = subspace (node->center, v);
    node->child[at_box] = t;
    node->child[at_voxel] = insert (e, node->child[at_voxel], v);
    return node;
}

/* vox_insert_voxel_, tree.c:506-611; the voxel is known to be absent.  Returns what takes t's place. */
static struct enode* insert (const struct vox_edit *e, struct enode *t, const float v[3])
{
    const float *voxel = e->voxel;
    realize (e, t);
    if (!t)
    {
        vox_dot one;
        vox_dot_set (one, v[0], v[1], v[2]);
        return subtree_of (&one, 1, voxel);
    }

    /* does the voxel complete a solid box? (tree.c:522-538) */
    float old_min[3], old_max[3];
    memcpy (old_min, t->bb_min, sizeof old_min);
    memcpy (old_max, t->bb_max, sizeof old_max);
    grow_box (t->bb_min, t->bb_max, v, voxel);
    if (dense_box (t->bb_min, t->bb_max, (size_t) t->dots_num + 1, voxel))
    {
        if (t->flags & VOX_DENSE_LEAF) { t->dots_num++; return t; }
        struct enode *solid = dense_leaf (t->bb_min, t->bb_max, voxel);
        enode_free_tree (t);
        return solid;
    }
    if (t->flags & VOX_DENSE_LEAF)
    {
        /* a solid box keeps its exact extent */
        memcpy (t->bb_min, old_min, sizeof old_min);
        memcpy (t->bb_max, old_max, sizeof old_max);
        if (t->dots_num >= VOX_MAX_DOTS) return wrap_dense (e, t, v);
        /* small: turn it into an ordinary leaf (tree.c:556-568) */
        vox_dot set[VOX_MAX_DOTS + 1];
        size_t n = collect (e, t, voxel, set);
        vox_dot_set (set[n], v[0], v[1], v[2]);
        struct enode *node = subtree_of (set, n + 1, voxel);
        enode_free_tree (t);
        return node;
    }
    if (t->flags & VOX_LEAF)
    {
        if (t->dots_num < VOX_MAX_DOTS)
        {
            memcpy (t->dots[t->dots_num], v, 3*sizeof (float));
            t->dots_num++;
            return t;
        }
        /* full (tree.c:581-590; an overflowed leaf, which the reference only asserts against,
           is re-split together with the new voxel) */
        size_t n = t->dots_num;
        vox_dot *set = malloc ((n + 1) * sizeof (vox_dot));
        collect (e, t, voxel, set);
        vox_dot_set (set[n], v[0], v[1], v[2]);
        struct enode *node = subtree_of (set, n + 1, voxel);
        free (set);
        enode_free_tree (t);
        return node;
    }
    /* inner node: count and descend (tree.c:593-607) */
    int idx = subspace (t->center, v);
    t->dots_num++;
    t->child[idx] = insert (e, t->child[idx], v);
    return t;
}

/* --------------------------------------------------------------- deletion */

/* delete_from_dense_stripe, tree.c:641-693: a stripe loses an end voxel by shrinking, an inner voxel by splitting in two.
   NULL if t is not a stripe. */
static struct enode* cut_stripe (struct enode *t, const float v[3], const float voxel[3])
{
    int axis = 0;
    if (!is_stripe (t, voxel, &axis)) return NULL;
    if (same_dot (t->bb_min, v))
    {
        t->dots_num--;
        t->bb_min[axis] += voxel[axis];
        return t;
    }
    float far_corner[3] = { v[0] + voxel[0], v[1] + voxel[1], v[2] + voxel[2] };
    if (same_dot (t->bb_max, far_corner))
    {
        t->dots_num--;
        t->bb_max[axis] -= voxel[axis];
        return t;
    }
    struct enode *node = enode_new (0);
    memcpy (node->bb_min, t->bb_min, sizeof node->bb_min);
    memcpy (node->bb_max, t->bb_max, sizeof node->bb_max);
    node->dots_num = t->dots_num - 1;
    memcpy (node->center, v, sizeof node->center);
    float mn[3], mx[3];
    octant (t, v, 1 << axis, mn, mx);
    node->child[1 << axis] = dense_leaf (mn, mx, voxel);
    octant (t, v, 0, mn, mx);
    struct enode *upper = node->child[0] = dense_leaf (mn, mx, voxel);
    upper->dots_num--;
    upper->bb_min[axis] += voxel[axis];
    enode_free_tree (t);
    return node;
}

/* delete_from_big_dense, tree.c:695-756: split a solid box around the voxel into up to 8 solid boxes; the octant that
   starts AT the voxel then loses its min-corner voxel by the same routine. */
static struct enode* cut_dense (struct enode *t, const float v[3], const float voxel[3])
{
    struct enode *node = cut_stripe (t, v, voxel);
    if (node) return node;

    float mn[3], mx[3];
    node = enode_new (0);
    memcpy (node->bb_min, t->bb_min, sizeof node->bb_min);
    memcpy (node->bb_max, t->bb_max, sizeof node->bb_max);
    node->dots_num = t->dots_num - 1;
    if (same_dot (t->bb_min, v))
    {
        /* min-corner voxel: split at its far corner, octant 7 (the voxel itself) stays empty */
        for (int i=0; i<3; i++) node->center[i] = v[i] + voxel[i];
        for (int i=0; i<VOX_NS-1; i++)
            if (octant (t, node->center, i, mn, mx)) node->child[i] = dense_leaf (mn, mx, voxel);
        enode_free_tree (t);
        return node;
    }
    memcpy (node->center, v, sizeof node->center);
    for (int i=0; i<VOX_NS; i++)
        if (octant (t, v, i, mn, mx)) node->child[i] = dense_leaf (mn, mx, voxel);
    node->child[0] = cut_dense (node->child[0], v, voxel);
    enode_free_tree (t);
    return node;
}

/* vox_delete_voxel_, tree.c:759-845; the voxel is known to be present.  Returns what takes t's place. */
static struct enode* delete (const struct vox_edit *e, struct enode *t, const float v[3])
{
    const float *voxel = e->voxel;
    realize (e, t);
    if (t->dots_num == 1) { enode_free_tree (t); return NULL; }
    if (t->flags & VOX_LEAF)
    {
        /* the bounding box is left as it is (tree.c:772) */
        uint32_t i = 0;
        while (i < t->dots_num && !same_dot (t->dots[i], v)) i++;
        if (i == t->dots_num) return t;
        t->dots_num--;
        memmove (t->dots + i, t->dots + i + 1, (size_t) (t->dots_num - i) * sizeof *t->dots);
        return t;
    }
    if (t->flags & VOX_DENSE_LEAF)
    {
        if (t->dots_num > VOX_MAX_DOTS) return cut_dense (t, v, voxel);
        struct enode *node = cut_stripe (t, v, voxel);
        if (node) return node;
        vox_dot set[VOX_MAX_DOTS];
        size_t n = collect (e, t, voxel, set), i = 0;
        while (i < n && !same_dot (set[i], v)) i++;
        if (i < n) { memmove (set + i, set + i + 1, (n - i - 1) * sizeof (vox_dot)); n--; }
        node = subtree_of (set, n, voxel);
        enode_free_tree (t);
        return node;
    }
    int idx = subspace (t->center, v);
    struct enode *child = t->child[idx];
    if (!child) return t;    /* unreachable for a voxel `contains` found */
    if (child->dots_num == t->dots_num)
    {
        /* an only child takes its parent's place (tree.c:828-839) */
        free (t);
        return delete (e, child, v);
    }
    t->dots_num--;
    t->child[idx] = delete (e, child, v);
    return t;
}

/* ------------------------------------------------------------------- API */

/* vox_align, tree.c:185-195 */
static void align_voxel (vox_dot dot)
{
    for (int i=0; i<3; i++) dot[i] = vox_voxel[i] * floorf (dot[i] / vox_voxel[i]);
}

static struct vox_tree_store* editable (struct vox_node **tree_ptr, const char *who)
{
    struct vox_node *tree = *tree_ptr;
    /* not vox_node_valid(): that would re-flatten after every single edit of a burst */
    if (!tree || tree->magic != VOX_NODE_MAGIC || !tree->store || tree->id != 0)
    { vox_set_error ("%s: not the root of a live tree", who); return NULL; }
    struct vox_tree_store *st = tree->store;
    if (!st->edit) st->edit = edit_from_store (st);
    if (!st->edit) { vox_set_error ("%s: out of memory", who); return NULL; }
    return st;
}

int vox_insert_voxel (struct vox_node **tree_ptr, vox_dot voxel)
{
    align_voxel (voxel);
    if (!*tree_ptr)
    {
        /* an empty tree: the new voxel is the whole tree (tree.c:540-551) */
        vox_dot one;
        vox_dot_copy (one, voxel);
        *tree_ptr = vox_make_tree (&one, 1);
        return *tree_ptr != NULL;
    }
    struct vox_tree_store *st = editable (tree_ptr, "vox_insert_voxel");
    if (!st) return 0;
    pthread_mutex_lock (&st->lock);
    int absent = !contains (st->edit, st->edit->root, voxel);
    if (absent)
    {
        st->edit->root = insert (st->edit, st->edit->root, voxel);
        st->dirty = 1;
    }
    pthread_mutex_unlock (&st->lock);
    return absent;
}

int vox_insert_voxel_coord (struct vox_node **tree_ptr, float x, float y, float z)
{
    vox_dot dot;
    vox_dot_set (dot, x, y, z);
    return vox_insert_voxel (tree_ptr, dot);
}

int vox_delete_voxel (struct vox_node **tree_ptr, vox_dot voxel)
{
    align_voxel (voxel);
    if (!*tree_ptr) return 0;
    struct vox_tree_store *st = editable (tree_ptr, "vox_delete_voxel");
    if (!st) return 0;
    pthread_mutex_lock (&st->lock);
    int present = contains (st->edit, st->edit->root, voxel);
    if (present)
    {
        st->edit->root = delete (st->edit, st->edit->root, voxel);
        st->dirty = 1;
    }
    int emptied = present && !st->edit->root;
    pthread_mutex_unlock (&st->lock);
    if (emptied)
    {
        /* the last voxel went: the tree is the NULL tree again (tree.c:765-770) */
        st->dirty = 0;
        vox_store_free (st);
        *tree_ptr = NULL;
    }
    return present;
}

int vox_delete_voxel_coord (struct vox_node **tree_ptr, float x, float y, float z)
{
    vox_dot dot;
    vox_dot_set (dot, x, y, z);
    return vox_delete_voxel (tree_ptr, dot);
}

struct vox_node* vox_rebuild_tree (const struct vox_node *tree)
{
    /* tree.c:424-435: list the voxels in tree order, build anew; the old tree stays valid */
    if (!tree) return NULL;
    if (!vox_node_valid (tree)) { vox_set_error ("vox_rebuild_tree: not a live tree"); return NULL; }
    if (tree->id != 0) { vox_set_error ("vox_rebuild_tree: pass the root"); return NULL; }
    struct vox_tree_store *st = tree->store;
    struct vox_edit *e = st->edit, *tmp = NULL;
    if (!e) e = tmp = edit_from_store (st);
    if (!e) { vox_set_error ("vox_rebuild_tree: out of memory"); return NULL; }
    size_t n = e->root ? e->root->dots_num : 0;
    vox_dot *set = vox_alloc (sizeof (vox_dot) * (n ? n : 1));
    struct vox_node *res = NULL;
    if (set)
    {
        size_t got = collect (e, e->root, e->voxel, set);
        res = vox_make_tree (set, got);
        free (set);
    }
    vox_edit_free (tmp);
    return res;
}
