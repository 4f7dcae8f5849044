/*
 * ref_glue.c -- TEST INFRASTRUCTURE ONLY (see oracle_api.h).
 *
 * Glue that exposes the UNMODIFIED reference sources (compiled where they lie
 * under /root/reference by oracle/Makefile, scalar build = SSE_INTRIN off,
 * gcc -std=gnu11 -O2 -ffp-contract=off) through the oracle API.  Nothing here
 * re-implements the search, the tree builder or the cameras: every such call
 * lands in the reference's own object code.  What IS restated here, because
 * the reference file cannot be compiled in this image (Blocks/GCD/SDL2/vn3d):
 *   - the vox_render block loop, via render_loop.h (renderer.c:200-327);
 * and what is new:
 *   - breadth-first flattening of the reference's pointer tree, so the GPU
 *     traversal can be tested on exactly the tree the reference built.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <omp.h>

/* reference headers (need -DVOXVISION_BUILDING -DVOXTREES_SOURCE -DVOXRND_SOURCE) */
#include "voxtrees/tree.h"
#include "voxtrees/search.h"
#include "voxtrees/geom.h"
#include "voxrnd/camera.h"
#include "voxrnd/modules.h"
/* lights.h pulls in mtree.h, whose Block-typed declaration (mtree.h:124-126) gcc cannot parse: the
   Makefile puts copies of both headers with those three lines deleted first on the include path */
#include "voxrnd/lights.h"

#include "oracle_api.h"
#include "render_loop.h"

/* The oracle API speaks packed float[3]; the reference's vox_dot is float[3]
 * (scalar build) or 16-byte aligned float[4] (SSE build).  DOT_IN/DOT_OUT move
 * the three coordinates across. */
#define DOT_IN(dst, src)  do { memset ((dst), 0, sizeof (vox_dot)); memcpy ((dst), (src), 3*sizeof (float)); } while (0)
#define DOT_OUT(dst, src) memcpy ((dst), (src), 3*sizeof (float))

/* the two camera files are compiled with -Dmodule_init=<kind>_module_init */
struct vox_module* simple_module_init (void);
struct vox_module* doom_module_init (void);

/* Layout mirrors of the private camera structs (simple-camera.c:10-17,
 * doom-camera.c:8-15), used only to read the POD state out. */
struct simple_camera_mirror
{
    struct vox_camera_interface *iface;
    float xsub, ysub;
    vox_dot position;
    vox_quat rotation;
    float mul, fov;
};
struct doom_camera_mirror
{
    struct vox_camera_interface *iface;
    float xsub, ysub;
    vox_dot position;
    float phi, sinphi, cosphi, k;
    float mul, fov;
};

struct id_pair { const struct vox_node *node; int32_t id; };

struct orc_tree
{
    struct vox_node *root;
    vox_dot *set;     /* the (reordered) input copy */
    size_t n;
    /* breadth-first order, built lazily */
    const struct vox_node **bfs;
    uint64_t *dots_off;
    struct id_pair *ids;  /* sorted by pointer */
    orc_tree_info info;
    int indexed;
};

struct orc_camera
{
    int kind;
    struct vox_camera *cam;
};

#ifdef SSE_INTRIN
const char* orc_impl_name (void) { return "reference-sse"; }
#else
const char* orc_impl_name (void) { return "reference-scalar"; }
#endif

void orc_set_voxel (float x, float y, float z) { vox_set_voxel_size (x, y, z); }
void orc_get_voxel (float out[3]) { DOT_OUT (out, vox_voxel); }

orc_tree* orc_tree_build (const float *xyz, size_t n)
{
    orc_tree *t = calloc (1, sizeof *t);
    t->n = n;
    t->set = aligned_alloc (16, (((n ? n : 1) * sizeof (vox_dot)) + 15) & ~(size_t) 15);
    for (size_t i=0; i<n; i++) DOT_IN (t->set[i], xyz + 3*i);
    t->root = vox_make_tree (t->set, n);
    return t;
}

orc_tree* orc_tree_build_by_insertion (const float *xyz, size_t n)
{
    orc_tree *t = calloc (1, sizeof *t);
    t->n = n;
    t->set = aligned_alloc (16, (((n ? n : 1) * sizeof (vox_dot)) + 15) & ~(size_t) 15);
    for (size_t i=0; i<n; i++) DOT_IN (t->set[i], xyz + 3*i);
    for (size_t i=0; i<n; i++)
    {
        vox_dot d;
        vox_dot_copy (d, t->set[i]);
        vox_insert_voxel (&t->root, d);
    }
    return t;
}

/* A tree given as flat arrays (orc_flat layout), rebuilt as the reference's heap nodes
   (tree.h:22-38; allocation sizes as node_alloc, tree.c:212-227). */
orc_tree* orc_tree_from_flat (const orc_flat *f, uint64_t n_nodes)
{
    orc_tree *t = calloc (1, sizeof *t);
    if (!n_nodes) return t;
    struct vox_node **byid = malloc (n_nodes * sizeof *byid);
    for (uint64_t id=0; id<n_nodes; id++)
    {
        struct vox_node *n = byid[id] = vox_alloc (sizeof (struct vox_node));
        memset (n, 0, sizeof *n);
        DOT_IN (n->bounding_box.min, f->bb_min + 3*id);
        DOT_IN (n->bounding_box.max, f->bb_max + 3*id);
        n->flags = f->flags[id];
        n->dots_num = f->dots_num[id];
        if (n->flags & VOX_DENSE_LEAF) continue;
        if (n->flags & VOX_LEAF)
        {
            size_t cap = n->dots_num > VOX_MAX_DOTS ? n->dots_num : VOX_MAX_DOTS;
            n->data.dots = vox_alloc (cap * sizeof (vox_dot));
            for (unsigned k=0; k<n->dots_num; k++) DOT_IN (n->data.dots[k], f->dots + 3*(f->dots_off[id] + k));
        }
        else DOT_IN (n->data.inner.center, f->center + 3*id);
    }
    for (uint64_t id=0; id<n_nodes; id++)
    {
        if (f->flags[id] & VOX_LEAF_MASK) continue;
        uint32_t next = f->first_child[id];
        for (int s=0; s<VOX_NS; s++) if (f->child_mask[id] & (1u << s)) byid[id]->data.inner.children[s] = byid[next++];
    }
    t->root = byid[0];
    free (byid);
    return t;
}

static void drop_index (orc_tree *t)
{
    free (t->bfs); free (t->dots_off); free (t->ids);
    t->bfs = NULL; t->dots_off = NULL; t->ids = NULL;
    t->indexed = 0;
}

/* vox_insert_voxel / vox_delete_voxel (tree.c:613-621, 847-855), one call per entry. */
int orc_tree_edit (orc_tree *t, const uint8_t *insert, const float *xyz, size_t n, uint8_t *result)
{
    drop_index (t);
    for (size_t i=0; i<n; i++)
    {
        vox_dot d;
        DOT_IN (d, xyz + 3*i);
        int r = insert[i] ? vox_insert_voxel (&t->root, d) : vox_delete_voxel (&t->root, d);
        if (result) result[i] = (uint8_t) r;
    }
    return 1;
}

/* vox_rebuild_tree (tree.c:424-435) */
orc_tree* orc_tree_rebuild (orc_tree *t)
{
    orc_tree *r = calloc (1, sizeof *r);
    r->root = vox_rebuild_tree (t->root);
    r->n = 0;
    r->set = NULL;
    return r;
}

void orc_tree_free (orc_tree *t)
{
    if (!t) return;
    vox_destroy_tree (t->root);
    free (t->set); free (t->bfs); free (t->dots_off); free (t->ids);
    free (t);
}

void orc_tree_reordered_set (orc_tree *t, float *xyz)
{
    for (size_t i=0; i<t->n; i++) DOT_OUT (xyz + 3*i, t->set[i]);
}

static int cmp_id_pair (const void *a, const void *b)
{
    const struct id_pair *x = a, *y = b;
    return (x->node > y->node) - (x->node < y->node);
}

static void count_nodes (const struct vox_node *node, uint32_t depth, orc_tree_info *info)
{
    if (!node) { info->n_null++; return; }
    info->n_nodes++;
    if (depth > info->depth) info->depth = depth;
    if (node->flags & VOX_DENSE_LEAF) info->n_dense++;
    else if (node->flags & VOX_LEAF)
    {
        info->n_leaf++;
        if (node->flags & VOX_OVERFLOW) info->n_overflow++;
        info->n_dots += node->dots_num;
    }
    else
    {
        info->n_inner++;
        for (int i=0; i<VOX_NS; i++) count_nodes (node->data.inner.children[i], depth+1, info);
    }
}

static void build_index (orc_tree *t)
{
    if (t->indexed) return;
    memset (&t->info, 0, sizeof t->info);
    if (t->root) count_nodes (t->root, 1, &t->info);
    uint64_t n = t->info.n_nodes;
    t->bfs = malloc ((n ? n : 1) * sizeof *t->bfs);
    t->dots_off = malloc ((n ? n : 1) * sizeof *t->dots_off);
    t->ids = malloc ((n ? n : 1) * sizeof *t->ids);
    uint64_t head = 0, tail = 0, dots = 0;
    if (t->root) t->bfs[tail++] = t->root;
    while (head < tail)
    {
        const struct vox_node *node = t->bfs[head];
        t->ids[head].node = node; t->ids[head].id = (int32_t)head;
        t->dots_off[head] = dots;
        if (node->flags & VOX_DENSE_LEAF) {}
        else if (node->flags & VOX_LEAF) dots += node->dots_num;
        else for (int i=0; i<VOX_NS; i++)
        {
            const struct vox_node *c = node->data.inner.children[i];
            if (c) t->bfs[tail++] = c;
        }
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

void orc_tree_info_get (orc_tree *t, orc_tree_info *info)
{
    build_index (t);
    *info = t->info;
}

void orc_tree_flatten (orc_tree *t, orc_flat *out)
{
    build_index (t);
    uint64_t n = t->info.n_nodes, next_child = 1;
    for (uint64_t id=0; id<n; id++)
    {
        const struct vox_node *node = t->bfs[id];
        DOT_OUT (out->bb_min + 3*id, node->bounding_box.min);
        DOT_OUT (out->bb_max + 3*id, node->bounding_box.max);
        out->flags[id] = node->flags;
        out->dots_num[id] = node->dots_num;
        out->center[3*id] = out->center[3*id+1] = out->center[3*id+2] = 0;
        out->first_child[id] = 0;
        out->child_mask[id] = 0;
        out->dots_off[id] = t->dots_off[id];
        if (node->flags & VOX_DENSE_LEAF) {}
        else if (node->flags & VOX_LEAF)
            for (unsigned k=0; k<node->dots_num; k++)
                DOT_OUT (out->dots + 3*(t->dots_off[id] + k), node->data.dots[k]);
        else
        {
            DOT_OUT (out->center + 3*id, node->data.inner.center);
            out->first_child[id] = (uint32_t)next_child;
            uint32_t mask = 0;
            for (int i=0; i<VOX_NS; i++)
                if (node->data.inner.children[i]) { mask |= 1u << i; next_child++; }
            out->child_mask[id] = mask;
        }
    }
}

/* "Hit voxel" derived from (leaf, hit point): first dot of the leaf whose
 * closed box [dot, dot+vox_voxel] contains the point. */
static int64_t derive_dot (orc_tree *t, int32_t id, const float *p)
{
    const struct vox_node *node = t->bfs[id];
    if (node->flags & VOX_DENSE_LEAF) return -1;
    for (unsigned i=0; i<node->dots_num; i++)
    {
        const float *d = node->data.dots[i];
        int in = 1;
        for (int k=0; k<3; k++)
        {
            float mx = d[k] + vox_voxel[k];
            if (p[k] < d[k] || p[k] > mx) in = 0;
        }
        if (in) return (int64_t)(t->dots_off[id] + i);
    }
    return -2;
}

void orc_ray_batch (orc_tree *t, uint32_t start_node,
                    const float *origins, const float *dirs, size_t n,
                    int32_t *leaf_id, float *hit, int64_t *dot_idx,
                    orc_stats *stats, int threads)
{
    build_index (t);
    if (stats) memset (stats, 0, sizeof *stats);
    const struct vox_node *start = t->root ? t->bfs[start_node] : NULL;
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i=0; i<n; i++)
    {
        vox_dot o, d, res;
        DOT_IN (o, origins + 3*i);
        DOT_IN (d, dirs + 3*i);
        const struct vox_node *leaf = vox_ray_tree_intersection (start, o, d, res);
        int32_t id = leaf ? node_to_id (t, leaf) : -1;
        if (leaf_id) leaf_id[i] = id;
        if (leaf && hit) DOT_OUT (hit + 3*i, res);
        if (dot_idx) dot_idx[i] = leaf ? derive_dot (t, id, res) : -1;
    }
}

void orc_ball_batch (orc_tree *t, const float *centers, const float *radii,
                     size_t n, uint8_t *out, int threads)
{
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i=0; i<n; i++)
    {
        vox_dot c;
        DOT_IN (c, centers + 3*i);
        out[i] = (uint8_t) vox_tree_ball_collidep (t->root, c, radii[i]);
    }
}

/* ---- cameras: calls go through the reference vtable ---- */

orc_camera* orc_camera_new (int kind)
{
    struct vox_module *m = (kind == ORC_CAMERA_DOOM) ? doom_module_init () : simple_module_init ();
    struct vox_camera_interface *iface = (struct vox_camera_interface*) m->methods;
    orc_camera *c = calloc (1, sizeof *c);
    c->kind = kind;
    c->cam = iface->construct_camera (NULL);
    return c;
}

void orc_camera_free (orc_camera *c)
{
    if (!c) return;
    c->cam->iface->destroy_camera (c->cam);
    free (c);
}

void orc_camera_set_position (orc_camera *c, const float p[3])
{ vox_dot d; DOT_IN (d, p); c->cam->iface->set_property_dot (c->cam, "position", d); }
void orc_camera_set_fov (orc_camera *c, float fov)
{ c->cam->iface->set_property_number (c->cam, "fov", fov); }
void orc_camera_set_rotation (orc_camera *c, const float a[3])
{ vox_dot d; DOT_IN (d, a); c->cam->iface->set_property_dot (c->cam, "rotation", d); }
void orc_camera_look_at (orc_camera *c, const float p[3])
{ vox_dot d; DOT_IN (d, p); c->cam->iface->look_at (c->cam, d); }
void orc_camera_rotate (orc_camera *c, const float delta[3])
{ vox_dot d; DOT_IN (d, delta); c->cam->iface->rotate_camera (c->cam, d); }
void orc_camera_move (orc_camera *c, const float delta[3])
{ vox_dot d; DOT_IN (d, delta); c->cam->iface->move_camera (c->cam, d); }
void orc_camera_set_window_size (orc_camera *c, int w, int h)
{ c->cam->iface->set_window_size (c->cam, w, h); }
void orc_camera_get_position (orc_camera *c, float p[3])
{ vox_dot d; c->cam->iface->get_position (c->cam, d); DOT_OUT (p, d); }
void orc_camera_screen2world (orc_camera *c, float dir[3], int sx, int sy)
{ vox_dot d; c->cam->iface->screen2world (c->cam, d, sx, sy); DOT_OUT (dir, d); }

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
    if (c->kind == ORC_CAMERA_DOOM)
    {
        const struct doom_camera_mirror *m = (const void*) c->cam;
        DOT_OUT (s, m->position);
        s[3] = m->sinphi; s[4] = m->cosphi; s[5] = m->k; s[6] = m->phi;
        s[7] = m->mul; s[8] = m->xsub; s[9] = m->ysub; s[10] = m->fov;
    }
    else
    {
        const struct simple_camera_mirror *m = (const void*) c->cam;
        DOT_OUT (s, m->position);
        memcpy (s+3, m->rotation, 4*sizeof(float));
        s[7] = m->mul; s[8] = m->xsub; s[9] = m->ysub; s[10] = m->fov;
    }
}

/* ---- render: restated block loop, reference search + camera underneath ---- */

static const void* hook_ray (const void *node, const float *o, const float *d, float *res)
{
    vox_dot o_, d_, r_;
    DOT_IN (o_, o); DOT_IN (d_, d);
    const struct vox_node *leaf = vox_ray_tree_intersection (node, o_, d_, r_);
    if (leaf) DOT_OUT (res, r_);
    return leaf;
}
static void hook_s2w (void *cam, float *dir, int sx, int sy)
{ struct vox_camera *c = cam; vox_dot d; c->iface->screen2world (c, d, sx, sy); DOT_OUT (dir, d); }
static void hook_pos (void *cam, float *pos)
{ struct vox_camera *c = cam; vox_dot d; c->iface->get_position (c, d); DOT_OUT (pos, d); }
static float hook_sqr_metric (const float *a, const float *b)
{ vox_dot a_, b_; DOT_IN (a_, a); DOT_IN (b_, b); return vox_sqr_metric (a_, b_); }
static float hook_sqr_norm (const float *a) { vox_dot a_; DOT_IN (a_, a); return vox_sqr_norm (a_); }

/* ---- lights: the reference's light manager and M-tree (lights.c, mtree.c object code) ---- */

/* layout mirror of the private struct (lights.c:11-14), to reach the M-tree root */
struct light_manager_mirror
{
    vox_dot ambient_light;
    struct vox_mtree_node *bound_lights;
};

struct orc_lights
{
    struct vox_light_manager *lm;   /* reference-driven, or */
    orc_lights_flat *flat;          /* given as flat arrays */
};

orc_lights* orc_lights_new (void)
{
    orc_lights *l = calloc (1, sizeof *l);
    l->lm = vox_make_light_manager ();
    return l;
}

static void* dup_bytes (const void *p, size_t n)
{
    void *q = malloc (n ? n : 1);
    if (n) memcpy (q, p, n);
    return q;
}

orc_lights* orc_lights_from_flat (const orc_lights_flat *f)
{
    orc_lights *l = calloc (1, sizeof *l);
    l->flat = malloc (sizeof *l->flat);
    *l->flat = *f;
    l->flat->node_bounds = dup_bytes (f->node_bounds, (size_t) f->n_nodes * 4 * sizeof (float));
    l->flat->node_link = dup_bytes (f->node_link, (size_t) f->n_nodes * 4 * sizeof (uint32_t));
    l->flat->sphere_geom = dup_bytes (f->sphere_geom, (size_t) f->n_spheres * 4 * sizeof (float));
    l->flat->sphere_light = dup_bytes (f->sphere_light, (size_t) f->n_spheres * 4 * sizeof (float));
    return l;
}

void orc_lights_free (orc_lights *l)
{
    if (!l) return;
    if (l->lm) vox_destroy_light_manager (l->lm);
    if (l->flat)
    {
        free (l->flat->node_bounds); free (l->flat->node_link); free (l->flat->sphere_geom); free (l->flat->sphere_light);
        free (l->flat);
    }
    free (l);
}

int orc_lights_insert (orc_lights *l, const float center[3], float radius, const float color[3])
{
    if (!l->lm) return -1;
    vox_dot c_, col_;
    DOT_IN (c_, center); DOT_IN (col_, color);
    return vox_insert_shadowless_light (l->lm, c_, radius, col_);
}

int orc_lights_delete (orc_lights *l, const float center[3], float radius)
{
    if (!l->lm) return -1;
    vox_dot c_;
    DOT_IN (c_, center);
    return vox_delete_shadowless_light (l->lm, c_, radius);
}

void orc_lights_clear (orc_lights *l) { if (l->lm) vox_delete_shadowless_lights (l->lm); }
int orc_lights_count (orc_lights *l) { return l->lm ? vox_shadowless_lights_number (l->lm) : (int) l->flat->n_spheres; }

int orc_lights_set_ambient (orc_lights *l, const float color[3])
{
    if (l->lm) { vox_dot c_; DOT_IN (c_, color); return vox_set_ambient_light (l->lm, c_); }
    for (int i=0; i<3; i++) if (color[i] < 0 || color[i] > 1) return 0;
    memcpy (l->flat->ambient, color, sizeof l->flat->ambient);
    return 1;
}

/* pre-order walk of the reference's M-tree into the flat layout */
static void flatten_mtree (const struct vox_mtree_node *n, orc_lights_flat *o)
{
    uint32_t me = o->n_nodes++, first = o->n_spheres;
    if (o->node_bounds)
    {
        DOT_OUT (o->node_bounds + 4*me, n->bounding_sphere.center);
        o->node_bounds[4*me + 3] = n->bounding_sphere.sqr_radius;
    }
    if (n->leaf)
        for (unsigned i=0; i<n->num; i++)
        {
            uint32_t k = o->n_spheres++;
            if (!o->sphere_geom) continue;
            DOT_OUT (o->sphere_geom + 4*k, n->data.spheres[i].center);
            o->sphere_geom[4*k + 3] = n->data.spheres[i].sqr_radius;
            DOT_OUT (o->sphere_light + 4*k, n->data.spheres[i].color);
            o->sphere_light[4*k + 3] = n->data.spheres[i].radius;
        }
    else for (unsigned i=0; i<n->num; i++) flatten_mtree (n->data.children[i], o);
    if (o->node_link)
    {
        o->node_link[4*me] = o->n_nodes;
        o->node_link[4*me + 1] = n->leaf ? 1u : 0u;
        o->node_link[4*me + 2] = first;
        o->node_link[4*me + 3] = n->leaf ? n->num : 0u;
    }
}

void orc_lights_flatten (orc_lights *l, orc_lights_flat *out)
{
    if (l->flat)
    {
        const orc_lights_flat *f = l->flat;
        memcpy (out->ambient, f->ambient, sizeof out->ambient);
        out->n_nodes = f->n_nodes; out->n_spheres = f->n_spheres;
        if (!out->node_bounds) return;
        memcpy (out->node_bounds, f->node_bounds, (size_t) f->n_nodes * 4 * sizeof (float));
        memcpy (out->node_link, f->node_link, (size_t) f->n_nodes * 4 * sizeof (uint32_t));
        memcpy (out->sphere_geom, f->sphere_geom, (size_t) f->n_spheres * 4 * sizeof (float));
        memcpy (out->sphere_light, f->sphere_light, (size_t) f->n_spheres * 4 * sizeof (float));
        return;
    }
    const struct light_manager_mirror *m = (const void*) l->lm;
    DOT_OUT (out->ambient, m->ambient_light);
    out->n_nodes = out->n_spheres = 0;
    if (m->bound_lights) flatten_mtree (m->bound_lights, out);
}

static void hook_light (void *lights, const float *inter, float *light)
{
    orc_lights *l = lights;
    if (l->lm)
    {
        vox_dot p_, out_;
        DOT_IN (p_, inter);
        vox_get_light (l->lm, p_, out_);       /* lights.c:104-116 */
        DOT_OUT (light, out_);
    }
    else orc_flat_light (l->flat, hook_sqr_metric, inter, light);
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
        .scene = t->root, .camera = c->cam,
        .ray = hook_ray, .screen2world = hook_s2w, .get_position = hook_pos,
        .sqr_metric = hook_sqr_metric, .sqr_norm = hook_sqr_norm,
        .get_light = lights ? hook_light : NULL, .lights = lights,
    };
    DOT_OUT (hk.voxel, vox_voxel);
    return orc_render_loop (&hk, w, h, quality, texture, fmt, pixels, NULL, NULL, threads);
}

int orc_render (orc_tree *t, orc_camera *c, int w, int h, unsigned quality,
                const uint8_t *texture, const orc_pixfmt *fmt,
                uint32_t *pixels, int32_t *leaf_id, float *hit, int threads)
{
    build_index (t);
    orc_render_hooks hk = {
        .scene = t->root, .camera = c->cam,
        .ray = hook_ray, .screen2world = hook_s2w, .get_position = hook_pos,
        .sqr_metric = hook_sqr_metric, .sqr_norm = hook_sqr_norm,
        .node_id = leaf_id ? node_to_id : NULL, .idctx = t,
    };
    DOT_OUT (hk.voxel, vox_voxel);
    return orc_render_loop (&hk, w, h, quality, texture, fmt, pixels, leaf_id, hit, threads);
}
