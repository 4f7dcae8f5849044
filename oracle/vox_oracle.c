/*
 * vox_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle_api.h).
 *
 * Plain-C CPU restatement ("port") of the reference's hot path, SCALAR
 * variant (SSE_INTRIN off), to be compiled with -O2 -ffp-contract=off so that
 * every fp32 operation is the one the reference's scalar build performs.
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference/src).  Parity pin: tests/test_oracle_pin.py checks this
 * port bit-for-bit against the reference itself (oracle/_ref, built from the
 * reference's own sources) and against the golden fixtures under
 * tests/golden/ generated from it.
 *
 * The product never calls into this file.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <omp.h>

#include "oracle_api.h"
#include "render_loop.h"

typedef float vec3[3];

static vec3 g_voxel = {1.0f, 1.0f, 1.0f};   /* voxtrees/tree.c:11 */

#define MAX_DOTS 7   /* voxtrees/params.h:29 */
#define NS 8         /* voxtrees/params.h:24 */

typedef struct onode
{
    vec3 bb_min, bb_max;        /* voxtrees/tree.h:30-32 */
    unsigned flags, dots_num;
    vec3 *dots;                 /* leaf */
    vec3 center;                /* inner */
    struct onode *children[NS];
} onode;

/* ------------------------------------------------------------------ geometry */

/* voxtrees/geom.c:13-31 fit_into_box */
static int clamp_into_box (const float *mn, const float *mx, const float *p, float *out)
{
    int same = 1;
    for (int i=0; i<3; i++)
    {
        if (p[i] < mn[i])      { out[i] = mn[i]; same = 0; }
        else if (p[i] > mx[i]) { out[i] = mx[i]; same = 0; }
        else out[i] = p[i];
    }
    return same;
}

/* voxtrees/geom.c:69-99 hit_box */
static int ray_box (const float *mn, const float *mx, const float *o, const float *d, float *res)
{
    if (clamp_into_box (mn, mx, o, res)) return 1;

    int axis = 0;
    float tmax = -1;
    for (int i=0; i<3; i++)
    {
        float t = (res[i] - o[i]) / d[i];
        if (t > tmax) { axis = i; tmax = t; }
    }
    if (tmax <= 0) return 0;

    for (int i=0; i<3; i++)
    {
        if (i == axis) continue;
        float c = o[i] + tmax*d[i];
        if (c < mn[i] || c > mx[i]) return 0;
        res[i] = c;
    }
    return 1;
}

/* voxtrees/geom.c:101-123 hit_plane_within_box */
static int ray_plane_in_box (const float *o, const float *d, const float *plane, int axis,
                             float *res, const float *mn, const float *mx)
{
    float k = plane[axis] - o[axis];
    if (k == 0 || ((d[axis] < 0) != (k < 0))) return 0;
    k = k / d[axis];
    for (int i=0; i<3; i++)
    {
        res[i] = o[i] + k*d[i];
        if (res[i] < mn[i] || res[i] > mx[i]) return 0;
    }
    return 1;
}

/* voxtrees/geom.c:44-51 get_subspace_idx */
static int subspace_of (const float *center, const float *p)
{
    int s = 0;
    for (int i=0; i<3; i++) s |= (center[i] > p[i]) << i;
    return s;
}

/* voxtrees/geom.c:53-65 get_corrected_subspace_idx */
static int subspace_corrected (const float *center, const float *p, const float *dir)
{
    int s = 0;
    for (int i=0; i<3; i++)
    {
        int bit = (center[i] == p[i]) ? (dir[i] < 0) : (center[i] > p[i]);
        s |= bit << i;
    }
    return s;
}

/* voxtrees/geom.c:125-131 vox_abs_metric */
static float l1_metric (const float *a, const float *b)
{
    float r = 0;
    for (int i=0; i<3; i++) r += fabsf (a[i] - b[i]);
    return r;
}

/* voxtrees/geom.c:133-139 vox_sqr_metric; powf(x, 2.0) == x*x exactly (gcc
 * folds it at -O2; glibc's powf is exact for squares that are representable
 * and correctly rounded otherwise -- checked against _ref by the pin tests) */
static float sqr_metric (const float *a, const float *b)
{
    float r = 0;
    for (int i=0; i<3; i++) { float t = a[i] - b[i]; r += t*t; }
    return r;
}

/* voxtrees/geom.c:141-147 vox_sqr_norm */
static float sqr_norm (const float *a)
{
    float r = 0;
    for (int i=0; i<3; i++) r += a[i]*a[i];
    return r;
}

/* voxtrees/geom.c:149-154 box_ball_interp */
static int box_ball (const float *mn, const float *mx, const float *c, float radius)
{
    vec3 fitted;
    clamp_into_box (mn, mx, c, fitted);
    return sqr_metric (fitted, c) < (radius*radius);
}

/* voxtrees/geom.c:157-167 dense_set_p */
static int dense_set (const float *mn, const float *mx, size_t n)
{
    vec3 size;
    for (int i=0; i<3; i++) size[i] = mx[i] - mn[i];
    float bb_volume = size[0]*size[1]*size[2];
    float vox_volume = g_voxel[0]*g_voxel[1]*g_voxel[2];
    return fabsf (n*vox_volume - bb_volume) < vox_volume;
}

/* ------------------------------------------------------------ tree building */

/* voxtrees/tree.c:138-155 calc_bounding_box (scalar; note the else-if) */
static void bounding_box (vec3 *set, size_t n, float *mn, float *mx)
{
    memcpy (mn, set[0], sizeof (vec3));
    memcpy (mx, set[0], sizeof (vec3));
    for (size_t i=0; i<n; i++)
        for (int j=0; j<3; j++)
        {
            if (set[i][j] < mn[j]) mn[j] = set[i][j];
            else if (set[i][j] > mx[j]) mx[j] = set[i][j];
        }
    for (int j=0; j<3; j++) mx[j] = mx[j] + g_voxel[j];
}

/* voxtrees/tree.c:125-136 find_center (scalar): sequential fp32 sum of
 * (p - bbox.min) in array order, mean, + min, snapped UP to the voxel grid */
static void split_center (vec3 *set, size_t n, const float *mn, float *res)
{
    res[0] = res[1] = res[2] = 0;
    for (size_t i=0; i<n; i++)
        for (int j=0; j<3; j++)
        {
            float t = set[i][j] - mn[j];
            res[j] = t + res[j];
        }
    for (int j=0; j<3; j++)
        res[j] = ceilf ((res[j]/n + mn[j]) / g_voxel[j]) * g_voxel[j];
}

/* voxtrees/tree.c:166-183 sort_set (scalar): swap-to-front partition */
static size_t partition (vec3 *set, size_t n, size_t offset, int subspace, const float *center)
{
    size_t counter = offset;
    for (size_t i=offset; i<n; i++)
    {
        if (subspace_of (center, set[i]) == subspace)
        {
            vec3 tmp;
            memcpy (tmp, set[counter], sizeof tmp);
            memcpy (set[counter], set[i], sizeof tmp);
            memcpy (set[i], tmp, sizeof tmp);
            counter++;
        }
    }
    return counter;
}

/* voxtrees/tree.c:249-338 vox_make_tree */
static onode* make_tree (vec3 *set, size_t n)
{
    if (n == 0) return NULL;
    onode *node = calloc (1, sizeof *node);
    bounding_box (set, n, node->bb_min, node->bb_max);
    int densep = dense_set (node->bb_min, node->bb_max, n);
    int leafp = n <= MAX_DOTS;
    node->dots_num = (unsigned) n;
    if (densep) node->flags = ORC_DENSE_LEAF;
    else if (leafp)
    {
        node->dots = malloc (MAX_DOTS * sizeof (vec3));
        memcpy (node->dots, set, n * sizeof (vec3));
        node->flags = ORC_LEAF;
    }
    else
    {
        size_t offset = 0;
        node->flags = 0;
        split_center (set, n, node->bb_min, node->center);
        for (int idx=0; idx<NS; idx++)
        {
            size_t new_offset = partition (set, n, offset, idx, node->center);
            if (offset == 0 && new_offset == n)
            {
                /* tree.c:281-305: nothing separates -> one overflowed leaf */
                node->flags = ORC_LEAF | ORC_OVERFLOW;
                node->dots = malloc (n * sizeof (vec3));
                memcpy (node->dots, set, n * sizeof (vec3));
                memset (node->center, 0, sizeof node->center);
                break;
            }
            node->children[idx] = make_tree (set + offset, new_offset - offset);
            offset = new_offset;
        }
    }
    return node;
}

static void free_tree (onode *node)
{
    if (!node) return;
    if (node->flags & ORC_LEAF) free (node->dots);
    else if (!(node->flags & 3u)) for (int i=0; i<NS; i++) free_tree (node->children[i]);
    free (node);
}

/* ------------------------------------------------------------------- search */

/* voxtrees/search.c:8-153 vox_ray_tree_intersection.  win = winning dot index. */
static const onode* ray_tree (const onode *tree, const float *origin, const float *dir,
                              float *res, int *win, orc_stats *st)
{
    vec3 entry;
    const onode *leaf = NULL;

    if (st) { st->calls++; if (!tree) st->null_calls++; }
    if (!tree || !ray_box (tree->bb_min, tree->bb_max, origin, dir, entry)) return NULL;

    if (tree->flags & ORC_DENSE_LEAF)
    {
        /* search.c:31-37 */
        if (st) st->dense++;
        memcpy (res, entry, sizeof entry);
        *win = -1;
        return tree;
    }

    if (tree->flags & ORC_LEAF)
    {
        /* search.c:39-81 */
        float closest = INFINITY;
        if (st) { st->leaf++; st->dots_tested += tree->dots_num; }
        for (unsigned i=0; i<tree->dots_num; i++)
        {
            vec3 vmax, far;
            const float *vmin = tree->dots[i];
            for (int k=0; k<3; k++) vmax[k] = vmin[k] + g_voxel[k];
            if (ray_box (vmin, vmax, entry, dir, far))
            {
                float dist = l1_metric (entry, far);
                if (dist == 0)
                {
                    if (st) st->dots_tested -= tree->dots_num - i - 1;
                    memcpy (res, far, sizeof far);
                    *win = (int) i;
                    return tree;
                }
                if (dist < closest)
                {
                    closest = dist;
                    memcpy (res, far, sizeof far);
                    *win = (int) i;
                    leaf = tree;
                }
            }
        }
        return leaf;
    }

    /* inner node, search.c:83-150 */
    if (st) st->inner++;
    int subspace = subspace_corrected (tree->center, entry, dir);
    if ((leaf = ray_tree (tree->children[subspace], entry, dir, res, win, st))) return leaf;

    vec3 plane_hit[3];
    int plane_axis[3];
    unsigned count = 0;
    for (int i=0; i<3; i++)
    {
        plane_axis[count] = i;
        if (st) st->plane_tests++;
        if (ray_plane_in_box (entry, dir, tree->center, i, plane_hit[count],
                              tree->bb_min, tree->bb_max)) count++;
    }

    for (unsigned i=0; i<count; i++)
    {
        /* search.c:115-134: one selection-sort pass, strict <, swap */
        for (unsigned j=i+1; j<count; j++)
        {
            if (l1_metric (entry, plane_hit[j]) < l1_metric (entry, plane_hit[i]))
            {
                vec3 tmp; int t2;
                memcpy (tmp, plane_hit[j], sizeof tmp);
                memcpy (plane_hit[j], plane_hit[i], sizeof tmp);
                memcpy (plane_hit[i], tmp, sizeof tmp);
                t2 = plane_axis[j]; plane_axis[j] = plane_axis[i]; plane_axis[i] = t2;
            }
        }
        subspace ^= 1 << plane_axis[i];
        if ((leaf = ray_tree (tree->children[subspace], plane_hit[i], dir, res, win, st)))
            return leaf;
    }
    return NULL;
}

/* voxtrees/search.c:155-183 vox_tree_ball_collidep */
static int ball_tree (const onode *tree, const float *center, float radius)
{
    if (!tree) return 0;
    if (!box_ball (tree->bb_min, tree->bb_max, center, radius)) return 0;
    if (tree->flags & ORC_DENSE_LEAF) return 1;
    if (tree->flags & ORC_LEAF)
    {
        for (unsigned i=0; i<tree->dots_num; i++)
        {
            vec3 vmax;
            for (int k=0; k<3; k++) vmax[k] = tree->dots[i][k] + g_voxel[k];
            if (box_ball (tree->dots[i], vmax, center, radius)) return 1;
        }
        return 0;
    }
    for (int i=0; i<NS; i++)
        if (ball_tree (tree->children[i], center, radius)) return 1;
    return 0;
}

/* -------------------------------------------------------------- API: trees */

struct id_pair { const onode *node; int32_t id; };

struct orc_tree
{
    onode *root;
    vec3 *set;
    size_t n;
    const onode **bfs;
    uint64_t *dots_off;
    struct id_pair *ids;
    orc_tree_info info;
    int indexed;
};

const char* orc_impl_name (void) { return "port"; }
void orc_set_voxel (float x, float y, float z) { g_voxel[0] = x; g_voxel[1] = y; g_voxel[2] = z; }
void orc_get_voxel (float out[3]) { memcpy (out, g_voxel, sizeof g_voxel); }

orc_tree* orc_tree_build (const float *xyz, size_t n)
{
    orc_tree *t = calloc (1, sizeof *t);
    t->n = n;
    t->set = malloc ((n ? n : 1) * sizeof (vec3));
    memcpy (t->set, xyz, n * sizeof (vec3));
    t->root = make_tree (t->set, n);
    return t;
}

orc_tree* orc_tree_build_by_insertion (const float *xyz, size_t n)
{
    (void) xyz; (void) n;
    return NULL; /* tree mutation is out of scope */
}

/* A tree given as flat arrays (orc_flat layout), e.g. one that was edited elsewhere. */
orc_tree* orc_tree_from_flat (const orc_flat *f, uint64_t n_nodes)
{
    orc_tree *t = calloc (1, sizeof *t);
    if (!n_nodes) return t;
    onode **byid = malloc (n_nodes * sizeof *byid);
    for (uint64_t id=0; id<n_nodes; id++)
    {
        onode *n = byid[id] = calloc (1, sizeof *n);
        memcpy (n->bb_min, f->bb_min + 3*id, sizeof (vec3));
        memcpy (n->bb_max, f->bb_max + 3*id, sizeof (vec3));
        n->flags = f->flags[id];
        n->dots_num = f->dots_num[id];
        if (n->flags & ORC_DENSE_LEAF) continue;
        if (n->flags & ORC_LEAF)
        {
            size_t cap = n->dots_num > MAX_DOTS ? n->dots_num : MAX_DOTS;
            n->dots = malloc (cap * sizeof (vec3));
            memcpy (n->dots, f->dots + 3*f->dots_off[id], n->dots_num * sizeof (vec3));
        }
        else memcpy (n->center, f->center + 3*id, sizeof (vec3));
    }
    for (uint64_t id=0; id<n_nodes; id++)
    {
        if (f->flags[id] & 3u) continue;
        uint32_t next = f->first_child[id];
        for (int s=0; s<NS; s++) if (f->child_mask[id] & (1u << s)) byid[id]->children[s] = byid[next++];
    }
    t->root = byid[0];
    free (byid);
    return t;
}

int orc_tree_edit (orc_tree *t, const uint8_t *insert, const float *xyz, size_t n, uint8_t *result)
{
    (void) t; (void) insert; (void) xyz; (void) n; (void) result;
    return 0; /* reference glue only: the port keeps to the search / render path */
}

orc_tree* orc_tree_rebuild (orc_tree *t)
{
    (void) t;
    return NULL;
}

void orc_tree_free (orc_tree *t)
{
    if (!t) return;
    free_tree (t->root);
    free (t->set); free (t->bfs); free (t->dots_off); free (t->ids);
    free (t);
}

void orc_tree_reordered_set (orc_tree *t, float *xyz) { memcpy (xyz, t->set, t->n * sizeof (vec3)); }

static int cmp_id_pair (const void *a, const void *b)
{
    const struct id_pair *x = a, *y = b;
    return (x->node > y->node) - (x->node < y->node);
}

static void count_nodes (const onode *node, uint32_t depth, orc_tree_info *info)
{
    if (!node) { info->n_null++; return; }
    info->n_nodes++;
    if (depth > info->depth) info->depth = depth;
    if (node->flags & ORC_DENSE_LEAF) info->n_dense++;
    else if (node->flags & ORC_LEAF)
    {
        info->n_leaf++;
        if (node->flags & ORC_OVERFLOW) info->n_overflow++;
        info->n_dots += node->dots_num;
    }
    else
    {
        info->n_inner++;
        for (int i=0; i<NS; i++) count_nodes (node->children[i], depth+1, info);
    }
}

static void build_index (orc_tree *t)
{
    if (t->indexed) return;
    memset (&t->info, 0, sizeof t->info);
    if (t->root) count_nodes (t->root, 1, &t->info);
    uint64_t n = t->info.n_nodes, head = 0, tail = 0, dots = 0;
    t->bfs = malloc ((n ? n : 1) * sizeof *t->bfs);
    t->dots_off = malloc ((n ? n : 1) * sizeof *t->dots_off);
    t->ids = malloc ((n ? n : 1) * sizeof *t->ids);
    if (t->root) t->bfs[tail++] = t->root;
    while (head < tail)
    {
        const onode *node = t->bfs[head];
        t->ids[head].node = node; t->ids[head].id = (int32_t) head;
        t->dots_off[head] = dots;
        if (node->flags & ORC_DENSE_LEAF) {}
        else if (node->flags & ORC_LEAF) dots += node->dots_num;
        else for (int i=0; i<NS; i++) if (node->children[i]) t->bfs[tail++] = node->children[i];
        head++;
    }
    qsort (t->ids, n, sizeof *t->ids, cmp_id_pair);
    t->indexed = 1;
}

static int32_t node_to_id (void *ctx, const void *node)
{
    orc_tree *t = ctx;
    struct id_pair key = { node, 0 };
    struct id_pair *p = bsearch (&key, t->ids, t->info.n_nodes, sizeof key, cmp_id_pair);
    return p ? p->id : -2;
}

void orc_tree_info_get (orc_tree *t, orc_tree_info *info) { build_index (t); *info = t->info; }

void orc_tree_flatten (orc_tree *t, orc_flat *out)
{
    build_index (t);
    uint64_t n = t->info.n_nodes, next_child = 1;
    for (uint64_t id=0; id<n; id++)
    {
        const onode *node = t->bfs[id];
        memcpy (out->bb_min + 3*id, node->bb_min, sizeof (vec3));
        memcpy (out->bb_max + 3*id, node->bb_max, sizeof (vec3));
        out->flags[id] = node->flags;
        out->dots_num[id] = node->dots_num;
        memset (out->center + 3*id, 0, sizeof (vec3));
        out->first_child[id] = 0;
        out->child_mask[id] = 0;
        out->dots_off[id] = t->dots_off[id];
        if (node->flags & ORC_DENSE_LEAF) {}
        else if (node->flags & ORC_LEAF)
            memcpy (out->dots + 3*t->dots_off[id], node->dots, node->dots_num * sizeof (vec3));
        else
        {
            uint32_t mask = 0;
            memcpy (out->center + 3*id, node->center, sizeof (vec3));
            out->first_child[id] = (uint32_t) next_child;
            for (int i=0; i<NS; i++) if (node->children[i]) { mask |= 1u << i; next_child++; }
            out->child_mask[id] = mask;
        }
    }
}

void orc_ray_batch (orc_tree *t, uint32_t start_node,
                    const float *origins, const float *dirs, size_t n,
                    int32_t *leaf_id, float *hit, int64_t *dot_idx,
                    orc_stats *stats, int threads)
{
    build_index (t);
    const onode *start = t->root ? t->bfs[start_node] : NULL;
    orc_stats total;
    memset (&total, 0, sizeof total);
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads)
    {
        orc_stats local;
        memset (&local, 0, sizeof local);
#pragma omp for schedule(static)
        for (size_t i=0; i<n; i++)
        {
            vec3 res;
            int win = -1;
            const onode *leaf = ray_tree (start, origins + 3*i, dirs + 3*i, res, &win,
                                          stats ? &local : NULL);
            int32_t id = leaf ? node_to_id (t, leaf) : -1;
            if (leaf_id) leaf_id[i] = id;
            if (leaf && hit) memcpy (hit + 3*i, res, sizeof res);
            if (dot_idx) dot_idx[i] = (leaf && win >= 0) ? (int64_t)(t->dots_off[id] + win) : -1;
        }
#pragma omp critical
        {
            total.calls += local.calls; total.null_calls += local.null_calls;
            total.inner += local.inner; total.leaf += local.leaf; total.dense += local.dense;
            total.dots_tested += local.dots_tested; total.plane_tests += local.plane_tests;
        }
    }
    if (stats) *stats = total;
}

void orc_ball_batch (orc_tree *t, const float *centers, const float *radii,
                     size_t n, uint8_t *out, int threads)
{
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i=0; i<n; i++)
        out[i] = (uint8_t) ball_tree (t->root, centers + 3*i, radii[i]);
}

/* ---------------------------------------------------------------- cameras */

struct orc_camera
{
    int kind;
    vec3 position;
    float fov, mul, xsub, ysub;
    float quat[4];                 /* simple: (w, i, j, k), voxvision.h.in:33-41 */
    float phi, sinphi, cosphi, k;  /* doom */
};

/* voxrnd/vect-ops.c:8-16 */
static void cross (const float *a, const float *b, float *res)
{
    float i1 = a[0], j1 = a[1], k1 = a[2];
    float i2 = b[0], j2 = b[1], k2 = b[2];
    res[0] =  j1*k2 - k1*j2;
    res[1] = -i1*k2 + k1*i2;
    res[2] =  i1*j2 - j1*i2;
}

/* voxrnd/vect-ops.c:18-26 vox_quat_mul */
static void quat_mul (const float *q1, const float *q2, float *res)
{
    float e1 = q1[0], i1 = q1[1], j1 = q1[2], k1 = q1[3];
    float e2 = q2[0], i2 = q2[1], j2 = q2[2], k2 = q2[3];
    float r0 = e1*e2 - i1*i2 - j1*j2 - k1*k2;
    float r1 = j1*k2 - k1*j2 + i1*e2 + e1*i2;
    float r2 = k1*i2 - i1*k2 + j1*e2 + e1*j2;
    float r3 = i1*j2 - j1*i2 + k1*e2 + e1*k2;
    res[0] = r0; res[1] = r1; res[2] = r2; res[3] = r3;
}

/* voxrnd/vect-ops.c:28-36 vox_rotate_vector (and :38-46 for the _q form) */
static void rotate_vec (const float *q, const float *v, float *res)
{
    vec3 t, t2, in;
    memcpy (in, v, sizeof in);
    cross (q+1, in, t);
    for (int i=0; i<3; i++) t[i] *= 2.0;
    cross (q+1, t, t2);
    for (int i=0; i<3; i++) res[i] = in[i] + q[0]*t[i] + t2[i];
}

/* voxrnd/vect-ops.c:48-57 vox_quat_normalize */
static void quat_normalize (float *q)
{
    float sum = 0;
    for (int i=0; i<4; i++) sum += q[i]*q[i];
    float norm = sqrtf (sum);
    for (int i=0; i<4; i++) q[i] /= norm;
}

orc_camera* orc_camera_new (int kind)
{
    /* simple-camera.c:138-158, doom-camera.c:138-158 */
    orc_camera *c = calloc (1, sizeof *c);
    c->kind = kind;
    c->fov = 1.0;
    c->quat[0] = 1;
    c->cosphi = 1.0;
    return c;
}

void orc_camera_free (orc_camera *c) { free (c); }
void orc_camera_set_position (orc_camera *c, const float p[3]) { memcpy (c->position, p, sizeof (vec3)); }
void orc_camera_set_fov (orc_camera *c, float fov) { c->fov = fov; }
void orc_camera_get_position (orc_camera *c, float p[3]) { memcpy (p, c->position, sizeof (vec3)); }

void orc_camera_set_rotation (orc_camera *c, const float a[3])
{
    if (c->kind == ORC_CAMERA_DOOM)
    {
        /* doom-camera.c:77-82 */
        c->k = a[0];
        c->phi = a[2];
        c->sinphi = sinf (c->phi);
        c->cosphi = cosf (c->phi);
        return;
    }
    /* simple-camera.c:171-184 */
    float r0[4] = { cosf (a[0]/2), sinf (a[0]/2), 0, 0 };
    float r1[4] = { cosf (a[1]/2), 0, sinf (a[1]/2), 0 };
    float r2[4] = { cosf (a[2]/2), 0, 0, sinf (a[2]/2) };
    quat_mul (r1, r0, c->quat);
    quat_mul (r2, c->quat, c->quat);
}

void orc_camera_rotate (orc_camera *c, const float delta[3])
{
    if (c->kind == ORC_CAMERA_DOOM)
    {
        /* doom-camera.c:94-102 */
        c->k += delta[0];
        c->phi += delta[2];
        c->sinphi = sinf (c->phi);
        c->cosphi = cosf (c->phi);
        return;
    }
    /* simple-camera.c:49-103 */
    static const float orts[3][4] = { {0,1,0,0}, {0,0,1,0}, {0,0,0,1} };
    for (int i=2; i>=0; i--)
    {
        float ort[4];
        ort[0] = 0;
        rotate_vec (c->quat, orts[i]+1, ort+1);
        float sinang = sinf (delta[i]/2);
        float cosang = cosf (delta[i]/2);
        for (int j=0; j<3; j++) ort[j+1] = sinang*ort[j+1];
        ort[0] = cosang;
        quat_mul (ort, c->quat, c->quat);
    }
    quat_normalize (c->quat);
}

void orc_camera_look_at (orc_camera *c, const float p[3])
{
    vec3 sub;
    for (int i=0; i<3; i++) sub[i] = p[i] - c->position[i];
    if (c->kind == ORC_CAMERA_DOOM)
    {
        /* doom-camera.c:104-115 */
        c->k = sub[2] / hypotf (sub[0], sub[1]);
        c->phi = -atan2f (sub[0], sub[1]);
        c->cosphi = cosf (c->phi);
        c->sinphi = sinf (c->phi);
        return;
    }
    /* simple-camera.c:105-118 */
    c->quat[0] = 1; c->quat[1] = c->quat[2] = c->quat[3] = 0;
    vec3 rot = { atan2f (sub[2], hypotf (sub[0], sub[1])), 0, -atan2f (sub[0], sub[1]) };
    orc_camera_rotate (c, rot);
}

void orc_camera_move (orc_camera *c, const float delta[3])
{
    vec3 tr;
    if (c->kind == ORC_CAMERA_DOOM)
    {
        /* doom-camera.c:24-34, 84-92 */
        tr[0] = delta[0]*c->cosphi - delta[1]*c->sinphi;
        tr[1] = delta[1]*c->cosphi + delta[0]*c->sinphi;
        tr[2] = delta[2];
    }
    else rotate_vec (c->quat, delta, tr);  /* simple-camera.c:39-46 */
    for (int i=0; i<3; i++) c->position[i] = c->position[i] + tr[i];
}

void orc_camera_set_window_size (orc_camera *c, int w, int h)
{
    /* simple-camera.c:120-136 == doom-camera.c:117-136 */
    if (w > h)
    {
        c->xsub = c->fov;
        c->ysub = c->fov*h/w;
        c->mul  = c->fov*2/w;
    }
    else
    {
        c->ysub = c->fov;
        c->xsub = c->fov*w/h;
        c->mul  = c->fov*2/h;
    }
}

void orc_camera_screen2world (orc_camera *c, float dir[3], int sx, int sy)
{
    if (c->kind == ORC_CAMERA_DOOM)
    {
        /* doom-camera.c:36-57 */
        vec3 d;
        d[0] = c->mul*sx - c->xsub;
        d[1] = 1;
        d[2] = c->ysub - c->mul*sy;
        d[2] += c->k;
        float xtr = d[0]*c->cosphi - d[1]*c->sinphi;
        float ytr = d[1]*c->cosphi + d[0]*c->sinphi;
        dir[0] = xtr; dir[1] = ytr; dir[2] = d[2];
        return;
    }
    /* simple-camera.c:21-31 */
    vec3 ray = { c->mul*sx - c->xsub, 1, c->ysub - c->mul*sy };
    rotate_vec (c->quat, ray, dir);
}

void orc_camera_rays (orc_camera *c, int w, int h, float *origins, float *dirs)
{
    float pos[3];
    orc_camera_get_position (c, pos);
#pragma omp parallel for schedule(static)
    for (int y=0; y<h; y++)
        for (int x=0; x<w; x++)
        {
            size_t i = (size_t) y*w + x;
            memcpy (origins + 3*i, pos, sizeof pos);
            orc_camera_screen2world (c, dirs + 3*i, x, y);
        }
}

void orc_camera_get_state (orc_camera *c, float s[11])
{
    memset (s, 0, 11*sizeof(float));
    memcpy (s, c->position, sizeof (vec3));
    if (c->kind == ORC_CAMERA_DOOM)
    {
        s[3] = c->sinphi; s[4] = c->cosphi; s[5] = c->k; s[6] = c->phi;
    }
    else memcpy (s+3, c->quat, 4*sizeof(float));
    s[7] = c->mul; s[8] = c->xsub; s[9] = c->ysub; s[10] = c->fov;
}

/* ----------------------------------------------------------------- render */

static const void* hook_ray (const void *node, const float *o, const float *d, float *res)
{
    int win;
    return ray_tree (node, o, d, res, &win, NULL);
}
static void hook_s2w (void *cam, float *dir, int sx, int sy) { orc_camera_screen2world (cam, dir, sx, sy); }
static void hook_pos (void *cam, float *pos) { orc_camera_get_position (cam, pos); }

/* ------------------------------------------------------------------- lights */

struct orc_lights
{
    orc_lights_flat flat;    /* owns copies of the arrays */
};

static void* dup_bytes (const void *p, size_t n)
{
    void *q = malloc (n ? n : 1);
    if (n) memcpy (q, p, n);
    return q;
}

orc_lights* orc_lights_new (void) { return NULL; /* the M-tree's construction is reference-only */ }

orc_lights* orc_lights_from_flat (const orc_lights_flat *f)
{
    orc_lights *l = calloc (1, sizeof *l);
    l->flat = *f;
    l->flat.node_bounds = dup_bytes (f->node_bounds, (size_t) f->n_nodes * 4 * sizeof (float));
    l->flat.node_link = dup_bytes (f->node_link, (size_t) f->n_nodes * 4 * sizeof (uint32_t));
    l->flat.sphere_geom = dup_bytes (f->sphere_geom, (size_t) f->n_spheres * 4 * sizeof (float));
    l->flat.sphere_light = dup_bytes (f->sphere_light, (size_t) f->n_spheres * 4 * sizeof (float));
    return l;
}

void orc_lights_free (orc_lights *l)
{
    if (!l) return;
    free (l->flat.node_bounds); free (l->flat.node_link); free (l->flat.sphere_geom); free (l->flat.sphere_light);
    free (l);
}

int orc_lights_insert (orc_lights *l, const float center[3], float radius, const float color[3])
{ (void) l; (void) center; (void) radius; (void) color; return -1; }
int orc_lights_delete (orc_lights *l, const float center[3], float radius)
{ (void) l; (void) center; (void) radius; return -1; }
void orc_lights_clear (orc_lights *l) { (void) l; }
int orc_lights_count (orc_lights *l) { return (int) l->flat.n_spheres; }
int orc_lights_set_ambient (orc_lights *l, const float color[3])
{
    for (int i=0; i<3; i++) if (color[i] < 0 || color[i] > 1) return 0;   /* lights.c:81-89 */
    memcpy (l->flat.ambient, color, sizeof l->flat.ambient);
    return 1;
}

void orc_lights_flatten (orc_lights *l, orc_lights_flat *out)
{
    const orc_lights_flat *f = &l->flat;
    memcpy (out->ambient, f->ambient, sizeof out->ambient);
    out->n_nodes = f->n_nodes; out->n_spheres = f->n_spheres;
    if (!out->node_bounds) return;
    memcpy (out->node_bounds, f->node_bounds, (size_t) f->n_nodes * 4 * sizeof (float));
    memcpy (out->node_link, f->node_link, (size_t) f->n_nodes * 4 * sizeof (uint32_t));
    memcpy (out->sphere_geom, f->sphere_geom, (size_t) f->n_spheres * 4 * sizeof (float));
    memcpy (out->sphere_light, f->sphere_light, (size_t) f->n_spheres * 4 * sizeof (float));
}

static void hook_light (void *lights, const float *inter, float *light)
{
    orc_flat_light (&((orc_lights*) lights)->flat, sqr_metric, inter, light);
}

void orc_lights_get (orc_lights *l, const float *points, size_t n, float *rgb)
{
    for (size_t i=0; i<n; i++) hook_light (l, points + 3*i, rgb + 3*i);
}

int orc_render_lit (orc_tree *t, orc_camera *c, int w, int h, unsigned quality,
                    const uint8_t *texture, const orc_pixfmt *fmt, orc_lights *lights,
                    uint32_t *pixels, int threads)
{
    build_index (t);
    orc_render_hooks hk = {
        .scene = t->root, .camera = c,
        .ray = hook_ray, .screen2world = hook_s2w, .get_position = hook_pos,
        .sqr_metric = sqr_metric, .sqr_norm = sqr_norm,
        .get_light = lights ? hook_light : NULL, .lights = lights,
    };
    memcpy (hk.voxel, g_voxel, sizeof hk.voxel);
    return orc_render_loop (&hk, w, h, quality, texture, fmt, pixels, NULL, NULL, threads);
}

int orc_render (orc_tree *t, orc_camera *c, int w, int h, unsigned quality,
                const uint8_t *texture, const orc_pixfmt *fmt,
                uint32_t *pixels, int32_t *leaf_id, float *hit, int threads)
{
    build_index (t);
    orc_render_hooks hk = {
        .scene = t->root, .camera = c,
        .ray = hook_ray, .screen2world = hook_s2w, .get_position = hook_pos,
        .sqr_metric = sqr_metric, .sqr_norm = sqr_norm,
        .node_id = leaf_id ? node_to_id : NULL, .idctx = t,
    };
    memcpy (hk.voxel, g_voxel, sizeof hk.voxel);
    return orc_render_loop (&hk, w, h, quality, texture, fmt, pixels, leaf_id, hit, threads);
}
