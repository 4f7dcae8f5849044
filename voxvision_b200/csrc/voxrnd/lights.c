/*
 * lights.c -- light manager (reference: src/voxrnd/lights.c) and its device copy.
 *
 * The manager itself is host state: an ambient colour and an M-tree of spheres
 * (voxtrees/mtree.c).  What the render kernel reads is a flattened snapshot:
 * M-tree nodes in pre-order with a "skip" link each, spheres in leaf order (see
 * voxgpu_lights_flat).  The snapshot is rebuilt and uploaded by the first
 * vox_render after a change (a version counter), never per frame.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "private.h"
#include "mtree-node.h"

#define VOX_LIGHT_DEVICES 64

struct vox_light_manager
{
    vox_dot ambient_light;
    struct vox_mtree_node *bound_lights;
    unsigned long version;            /* bumped by every change */
    /* snapshots by CUDA ordinal, each made from `dev_version[d]` (a context may draw on several devices) */
    voxgpu_lights *dev[VOX_LIGHT_DEVICES];
    unsigned long dev_version[VOX_LIGHT_DEVICES];
};

/* lights.c:81-89 */
static int valid_colour (const vox_dot c)
{
    for (int i=0; i<3; i++) if (c[i] < 0 || c[i] > 1) return 0;
    return 1;
}

struct vox_light_manager* vox_make_light_manager (void)
{
    struct vox_light_manager *lm = calloc (1, sizeof *lm);
    if (!lm) return NULL;
    vox_dot_set (lm->ambient_light, 1, 1, 1);
    lm->version = 1;
    return lm;
}

void vox_destroy_light_manager (struct vox_light_manager *lm)
{
    if (!lm) return;
    vox_mtree_destroy (lm->bound_lights);
    int cur = voxgpu_get_device ();
    for (int d=0; d<VOX_LIGHT_DEVICES; d++)
        if (lm->dev[d]) { voxgpu_set_device (d); voxgpu_lights_free (lm->dev[d]); }
    if (cur >= 0) voxgpu_set_device (cur);
    free (lm);
}

int vox_set_ambient_light (struct vox_light_manager *lm, const vox_dot color)
{
    if (!valid_colour (color)) return 0;
    vox_dot_copy (lm->ambient_light, color);
    lm->version++;
    return 1;
}

int vox_insert_shadowless_light (struct vox_light_manager *lm, const vox_dot center, float radius, const vox_dot color)
{
    if (!valid_colour (color)) return 0;
    struct vox_sphere s;
    memset (&s, 0, sizeof s);
    vox_dot_copy (s.center, center);
    vox_dot_copy (s.color, color);
    s.radius = radius;
    int ok = vox_mtree_add_sphere (&lm->bound_lights, &s);
    if (ok) lm->version++;
    return ok;
}

int vox_delete_shadowless_light (struct vox_light_manager *lm, const vox_dot center, float radius)
{
    struct vox_sphere s;
    memset (&s, 0, sizeof s);
    vox_dot_copy (s.center, center);
    s.radius = radius;
    int ok = vox_mtree_remove_sphere (&lm->bound_lights, &s);
    if (ok) lm->version++;
    return ok;
}

void vox_delete_shadowless_lights (struct vox_light_manager *lm)
{
    vox_mtree_destroy (lm->bound_lights);
    lm->bound_lights = NULL;
    lm->version++;
}

int vox_shadowless_lights_number (const struct vox_light_manager *lm)
{
    return (int) vox_mtree_items (lm->bound_lights);
}

/* lights.c:91-102 */
struct light_sum { vox_dot at, colour; };

static void add_light (const struct vox_sphere *s, void *arg)
{
    struct light_sum *sum = arg;
    float dist = fminf (s->radius, sqrtf (vox_sqr_metric (s->center, sum->at)));
    float add = 1 - dist / s->radius;
    float r = sum->colour[0] + add * s->color[0];
    float g = sum->colour[1] + add * s->color[1];
    float b = sum->colour[2] + add * s->color[2];
    vox_dot_set (sum->colour, r, g, b);
}

void vox_get_light (const struct vox_light_manager *lm, const vox_dot intersection, vox_dot light)
{
    /* lights.c:104-116 */
    struct light_sum sum;
    vox_dot_copy (sum.colour, lm->ambient_light);
    vox_dot_copy (sum.at, intersection);
    vox_mtree_spheres_containing_f (lm->bound_lights, intersection, add_light, &sum);
    vox_dot_copy (light, sum.colour);
}

/* ---------------------------------------------------- snapshot for the GPU */

struct flat_out
{
    uint32_t nodes, spheres;
    float *node_bounds; uint32_t *node_link; float *sphere_geom; float *sphere_light;
};

static void flatten (const struct vox_mtree_node *n, struct flat_out *o)
{
    uint32_t me = o->nodes++;
    uint32_t first = o->spheres;
    if (o->node_bounds)
    {
        memcpy (o->node_bounds + 4*me, n->bounding_sphere.center, 3*sizeof (float));
        o->node_bounds[4*me + 3] = n->bounding_sphere.sqr_radius;
    }
    if (n->leaf)
    {
        for (unsigned i=0; i<n->num; i++)
        {
            uint32_t k = o->spheres++;
            if (!o->sphere_geom) continue;
            memcpy (o->sphere_geom + 4*k, n->spheres[i].center, 3*sizeof (float));
            o->sphere_geom[4*k + 3] = n->spheres[i].sqr_radius;
            memcpy (o->sphere_light + 4*k, n->spheres[i].color, 3*sizeof (float));
            o->sphere_light[4*k + 3] = n->spheres[i].radius;
        }
    }
    else for (unsigned i=0; i<n->num; i++) flatten (n->children[i], o);
    if (o->node_link)
    {
        o->node_link[4*me] = o->nodes;              /* first node after this subtree */
        o->node_link[4*me + 1] = n->leaf ? 1u : 0u;
        o->node_link[4*me + 2] = first;
        o->node_link[4*me + 3] = n->leaf ? n->num : 0u;
    }
}

void vox_light_manager_flatten (const struct vox_light_manager *lm, uint32_t *n_nodes, uint32_t *n_spheres,
                                float *node_bounds, uint32_t *node_link, float *sphere_geom, float *sphere_light)
{
    struct flat_out o = { 0, 0, node_bounds, node_link, sphere_geom, sphere_light };
    if (lm->bound_lights) flatten (lm->bound_lights, &o);
    if (n_nodes) *n_nodes = o.nodes;
    if (n_spheres) *n_spheres = o.spheres;
}

/* The snapshot for the current device, rebuilt if the lights changed since it was made. */
const voxgpu_lights* vox_light_manager_device (struct vox_light_manager *lm)
{
    int dev = voxgpu_get_device ();
    if (dev < 0 || dev >= VOX_LIGHT_DEVICES) { vox_set_error ("bad CUDA device %d", dev); return NULL; }
    if (lm->dev[dev] && lm->dev_version[dev] == lm->version) return lm->dev[dev];
    voxgpu_synchronize (NULL);                       /* frames in flight may still read the old snapshot */
    voxgpu_lights_free (lm->dev[dev]);
    lm->dev[dev] = NULL;

    uint32_t nn = 0, ns = 0;
    vox_light_manager_flatten (lm, &nn, &ns, NULL, NULL, NULL, NULL);
    voxgpu_lights_flat f;
    memset (&f, 0, sizeof f);
    float *bounds = malloc ((nn + 1) * 4 * sizeof (float)), *geom = malloc ((ns + 1) * 4 * sizeof (float));
    float *light = malloc ((ns + 1) * 4 * sizeof (float));
    uint32_t *link = malloc ((nn + 1) * 4 * sizeof (uint32_t));
    vox_light_manager_flatten (lm, NULL, NULL, bounds, link, geom, light);
    memcpy (f.ambient, lm->ambient_light, sizeof f.ambient);
    f.n_nodes = nn; f.n_spheres = ns;
    f.node_bounds = bounds; f.node_link = link; f.sphere_geom = geom; f.sphere_light = light;
    lm->dev[dev] = voxgpu_lights_upload (&f);
    free (bounds); free (geom); free (light); free (link);
    if (!lm->dev[dev]) { vox_set_error ("light upload failed: %s", voxgpu_last_error ()); return NULL; }
    lm->dev_version[dev] = lm->version;
    return lm->dev[dev];
}
