/*
 * search.c -- tree queries.  All of them run on the GPU through the thin
 * C-ABI layer (voxgpu.h); there is no CPU implementation behind them.
 *
 *   vox_ray_tree_intersection   reference: src/voxtrees/search.c:8-153
 *   vox_tree_ball_collidep      reference: src/voxtrees/search.c:155-183
 * and their batched forms (new).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "store.h"

/* Device copy of the tree on the current CUDA device: made on first use,
 * "uploaded once" (cudaMemcpyAsync from pinned staging, voxgpu_tree_upload). */
voxgpu_tree* vox_store_device (struct vox_tree_store *st)
{
    if (voxgpu_device_count () <= 0)
    {
        vox_set_error ("no CUDA device available (%s); voxvision-b200 has no CPU search path",
                       voxgpu_last_error ());
        return NULL;
    }
    int dev = voxgpu_get_device ();
    if (dev < 0 || dev >= VOX_MAX_DEVICES) { vox_set_error ("bad CUDA device %d", dev); return NULL; }
    pthread_mutex_lock (&st->lock);
    voxgpu_tree *t = st->dev[dev];
    if (!t)
    {
        voxgpu_flat f;
        memset (&f, 0, sizeof f);
        f.n_nodes = st->n_nodes; f.n_dots = st->n_dots; f.depth = st->depth;
        f.bb_min = st->bb_min; f.bb_max = st->bb_max; f.flags = st->flags; f.dots_num = st->dots_num;
        f.center = st->center; f.first_child = st->first_child; f.child_mask = st->child_mask;
        f.dots_off = st->dots_off; f.dots = st->dots;
        memcpy (f.voxel, st->voxel, sizeof f.voxel);
        t = st->dev[dev] = voxgpu_tree_upload (&f);
        if (!t) vox_set_error ("tree upload failed: %s", voxgpu_last_error ());
    }
    pthread_mutex_unlock (&st->lock);
    return t;
}

void* vox_tree_device_handle (const struct vox_node *tree)
{
    if (!vox_node_valid (tree)) { vox_set_error ("vox_tree_device_handle: not a live tree"); return NULL; }
    return vox_store_device (tree->store);
}

const struct vox_node*
vox_ray_tree_intersection (const struct vox_node *tree, const vox_dot origin,
                           const vox_dot dir, vox_dot res)
{
    vox_set_error ("");
    if (!tree) return NULL;   /* search.c:22: an empty tree is never hit */
    if (!vox_node_valid (tree)) { vox_set_error ("vox_ray_tree_intersection: not a live tree"); return NULL; }
    voxgpu_tree *dev = vox_store_device (tree->store);
    if (!dev) return NULL;
    vox_dot o, d, p;
    int32_t leaf = -1;
    vox_dot_copy (o, origin); vox_dot_copy (d, dir);
    memset (p, 0, sizeof p);
    if (voxgpu_rays_host (dev, tree->id, o, d, 1, &leaf, p, NULL, NULL))
    {
        vox_set_error ("vox_ray_tree_intersection: %s", voxgpu_last_error ());
        return NULL;
    }
    if (leaf < 0) return NULL;
    res[0] = p[0]; res[1] = p[1]; res[2] = p[2];
    return vox_store_handle (tree->store, (uint64_t) leaf);
}

int vox_tree_ball_collidep (const struct vox_node *tree, const vox_dot center, float radius)
{
    vox_set_error ("");
    if (!tree) return 0;      /* search.c:158 */
    if (!vox_node_valid (tree)) { vox_set_error ("vox_tree_ball_collidep: not a live tree"); return 0; }
    if (tree->id != 0) { vox_set_error ("vox_tree_ball_collidep: only whole trees are supported"); return 0; }
    voxgpu_tree *dev = vox_store_device (tree->store);
    if (!dev) return 0;
    vox_dot c;
    uint8_t out = 0;
    vox_dot_copy (c, center);
    if (voxgpu_balls_host (dev, c, &radius, 1, &out, NULL))
    {
        vox_set_error ("vox_tree_ball_collidep: %s", voxgpu_last_error ());
        return 0;
    }
    return out;
}

long vox_ray_tree_intersection_batch (const struct vox_node *tree,
                                      const vox_dot *origins, const vox_dot *dirs,
                                      size_t n, int32_t *leaf_ids, vox_dot *points, int32_t *voxels)
{
    vox_set_error ("");
    if (!tree)
    {
        if (leaf_ids) for (size_t i=0; i<n; i++) leaf_ids[i] = -1;
        if (voxels) for (size_t i=0; i<n; i++) voxels[i] = -1;
        return 0;
    }
    if (!vox_node_valid (tree)) { vox_set_error ("vox_ray_tree_intersection_batch: not a live tree"); return -1; }
    voxgpu_tree *dev = vox_store_device (tree->store);
    if (!dev) return -1;
    long hits = 0;
    if (voxgpu_rays_host (dev, tree->id, (const float*) origins, (const float*) dirs, n,
                          leaf_ids, (float*) points, voxels, &hits))
    {
        vox_set_error ("vox_ray_tree_intersection_batch: %s", voxgpu_last_error ());
        return -1;
    }
    return hits;
}

long vox_tree_ball_collidep_batch (const struct vox_node *tree, const vox_dot *centers,
                                   const float *radii, size_t n, uint8_t *out)
{
    vox_set_error ("");
    if (!tree) { memset (out, 0, n); return 0; }
    if (!vox_node_valid (tree) || tree->id != 0)
    { vox_set_error ("vox_tree_ball_collidep_batch: not a live whole tree"); return -1; }
    voxgpu_tree *dev = vox_store_device (tree->store);
    if (!dev) return -1;
    long hits = 0;
    if (voxgpu_balls_host (dev, (const float*) centers, radii, n, out, &hits))
    {
        vox_set_error ("vox_tree_ball_collidep_batch: %s", voxgpu_last_error ());
        return -1;
    }
    return hits;
}
