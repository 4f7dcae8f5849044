/*
 * oracle_api.h -- TEST INFRASTRUCTURE ONLY.
 *
 * One C interface, implemented twice:
 *   (1) oracle/vox_oracle.c  -- a plain-C CPU restatement ("port") of the
 *       reference's hot path (voxtrees search/geom/tree + voxrnd cameras and
 *       the vox_render block loop).  Built into oracle/_build/libvoxoracle.so.
 *   (2) oracle/ref_glue.c    -- thin glue around the UNMODIFIED reference
 *       sources, compiled where they lie under /root/reference (scalar build,
 *       gcc -O2 -ffp-contract=off = the oracle of record).
 *       Built into oracle/_ref/libvoxref_scalar.so (git-ignored).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load either library.  The product
 * (voxvision_b200/) never links, imports or calls anything declared here.
 *
 * All points are packed float[3] (the scalar reference's vox_dot,
 * src/voxvision.h.in:57); node ids are positions in the breadth-first
 * flattening defined by orc_tree_flatten().
 */
#ifndef VOX_ORACLE_API_H
#define VOX_ORACLE_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_EXPORT __attribute__((visibility("default")))

/* Node kinds as the reference encodes them in vox_node.flags (tree.h:15-18). */
#define ORC_LEAF        1u
#define ORC_DENSE_LEAF  2u
#define ORC_OVERFLOW    4u

/* Quality bits, renderer.h:81-127. */
#define ORC_QUALITY_BEST      0u
#define ORC_QUALITY_FAST      1u
#define ORC_QUALITY_ADAPTIVE  2u
#define ORC_QUALITY_RAY_MERGE           4u
#define ORC_QUALITY_RAY_MERGE_ACCURATE  8u

#define ORC_CAMERA_SIMPLE 0
#define ORC_CAMERA_DOOM   1

typedef struct orc_tree orc_tree;
typedef struct orc_camera orc_camera;

/* Sizes of a flattened tree. */
typedef struct
{
    uint64_t n_nodes;      /* non-NULL nodes */
    uint64_t n_inner, n_leaf, n_dense, n_overflow;
    uint64_t n_null;       /* NULL child slots of inner nodes */
    uint64_t n_dots;       /* dots stored in (non-dense) leaves */
    uint32_t depth;        /* root alone = 1 */
} orc_tree_info;

/* Flattened arrays, caller-allocated from orc_tree_info.  Breadth-first; the
 * non-NULL children of an inner node are consecutive, in subspace order. */
typedef struct
{
    float    *bb_min;      /* 3*n_nodes */
    float    *bb_max;      /* 3*n_nodes */
    uint32_t *flags;       /* n_nodes, reference flags (0 = inner) */
    uint32_t *dots_num;    /* n_nodes, reference dots_num */
    float    *center;      /* 3*n_nodes, zeros for leaves */
    uint32_t *first_child; /* n_nodes, id of first non-NULL child (inner) */
    uint32_t *child_mask;  /* n_nodes, bit s set <=> children[s] != NULL */
    uint64_t *dots_off;    /* n_nodes, offset (in dots) into dots[] (leaf) */
    float    *dots;        /* 3*n_dots */
} orc_flat;

/* Per-ray-batch traversal counters (node visits by kind) used for the
 * "algorithmic bytes per ray" of the roofline.  Only the port
 * fills them; the reference glue returns zeros. */
typedef struct
{
    uint64_t calls;        /* vox_ray_tree_intersection invocations */
    uint64_t null_calls;   /* ... with tree == NULL */
    uint64_t inner, leaf, dense; /* non-NULL nodes entered, by kind */
    uint64_t dots_tested;  /* leaf voxels tested */
    uint64_t plane_tests;  /* hit_plane_within_box calls */
} orc_stats;

/* Pixel format: the part of SDL_PixelFormat that SDL_MapRGB reads. */
typedef struct
{
    uint32_t amask;
    uint8_t rloss, gloss, bloss, pad0;
    uint8_t rshift, gshift, bshift, pad1;
} orc_pixfmt;

/* A light manager's lights as flat arrays (the layout the product uploads to the GPU,
 * include/voxgpu.h voxgpu_lights_flat): M-tree nodes in pre-order. */
typedef struct
{
    float ambient[3];
    uint32_t n_nodes, n_spheres;
    float    *node_bounds;   /* 4*n_nodes: bounding centre xyz, squared radius */
    uint32_t *node_link;     /* 4*n_nodes: skip, leaf flag, first sphere, count */
    float    *sphere_geom;   /* 4*n_spheres: centre xyz, squared radius */
    float    *sphere_light;  /* 4*n_spheres: colour rgb, radius */
} orc_lights_flat;
typedef struct orc_lights orc_lights;

ORC_EXPORT const char* orc_impl_name (void);  /* "port" or "reference-scalar" ... */

/* Global voxel size (params.h:38, tree.c:11). */
ORC_EXPORT void orc_set_voxel (float x, float y, float z);
ORC_EXPORT void orc_get_voxel (float out[3]);

/* vox_make_tree (tree.c:249-338).  The input is copied; n==0 gives a valid
 * handle wrapping a NULL tree. */
ORC_EXPORT orc_tree* orc_tree_build (const float *xyz, size_t n);
/* Tree built by successive vox_insert_voxel calls (tree.c:506-650); reference
 * glue only -- the port returns NULL (insertion is out of scope). */
ORC_EXPORT orc_tree* orc_tree_build_by_insertion (const float *xyz, size_t n);
/* Online edits: entry i inserts (insert[i] != 0) or deletes voxel xyz[3i..] with
 * vox_insert_voxel / vox_delete_voxel (tree.c:613-621, 847-855); result[i] = their return
 * value.  Returns 0 if unsupported (the port).  orc_tree_rebuild = vox_rebuild_tree
 * (tree.c:424-435), NULL if unsupported. */
/* A tree from flat arrays in the orc_flat layout (copied). */
ORC_EXPORT orc_tree* orc_tree_from_flat (const orc_flat *flat, uint64_t n_nodes);
ORC_EXPORT int orc_tree_edit (orc_tree *t, const uint8_t *insert, const float *xyz, size_t n, uint8_t *result);
ORC_EXPORT orc_tree* orc_tree_rebuild (orc_tree *t);
ORC_EXPORT void orc_tree_free (orc_tree *t);
ORC_EXPORT void orc_tree_info_get (orc_tree *t, orc_tree_info *info);
ORC_EXPORT void orc_tree_flatten (orc_tree *t, orc_flat *out);
/* The (destructively reordered) copy of the input set after building. */
ORC_EXPORT void orc_tree_reordered_set (orc_tree *t, float *xyz);

/* vox_ray_tree_intersection (search.c:8-153) for n rays, searched from node
 * start_node (0 = root).  leaf_id[i] = -1 on miss; hit[] untouched on miss.
 * dot_idx[i] = index of the winning dot in flat dots[] (leaf) or -1 (dense).
 * Any output pointer may be NULL. */
ORC_EXPORT void orc_ray_batch (orc_tree *t, uint32_t start_node,
                               const float *origins, const float *dirs, size_t n,
                               int32_t *leaf_id, float *hit, int64_t *dot_idx,
                               orc_stats *stats, int threads);

/* vox_tree_ball_collidep (search.c:155-183) for n balls. */
ORC_EXPORT void orc_ball_batch (orc_tree *t, const float *centers, const float *radii,
                                size_t n, uint8_t *out, int threads);

/* Cameras (simple-camera.c, doom-camera.c) through the reference vtable order
 * of calls; kind = ORC_CAMERA_*. */
ORC_EXPORT orc_camera* orc_camera_new (int kind);
ORC_EXPORT void orc_camera_free (orc_camera *c);
ORC_EXPORT void orc_camera_set_position (orc_camera *c, const float p[3]);
ORC_EXPORT void orc_camera_set_fov (orc_camera *c, float fov);
ORC_EXPORT void orc_camera_set_rotation (orc_camera *c, const float angles[3]);
ORC_EXPORT void orc_camera_look_at (orc_camera *c, const float p[3]);
ORC_EXPORT void orc_camera_rotate (orc_camera *c, const float delta[3]);
ORC_EXPORT void orc_camera_move (orc_camera *c, const float delta[3]);
ORC_EXPORT void orc_camera_set_window_size (orc_camera *c, int w, int h);
ORC_EXPORT void orc_camera_get_position (orc_camera *c, float p[3]);
ORC_EXPORT void orc_camera_screen2world (orc_camera *c, float dir[3], int sx, int sy);
/* All primary rays of a w x h frame, row-major: origins[3*(y*w+x)] = get_position,
 * dirs[...] = screen2world (x, y) -- the rays vox_render casts (renderer.c:236,292). */
ORC_EXPORT void orc_camera_rays (orc_camera *c, int w, int h, float *origins, float *dirs);
/* POD state: simple -> {pos[3], quat[4] (w,i,j,k), mul, xsub, ysub, fov};
 *            doom   -> {pos[3], sinphi, cosphi, k, phi, mul, xsub, ysub, fov}. */
ORC_EXPORT void orc_camera_get_state (orc_camera *c, float state[11]);

/* vox_render + vox_redraw (renderer.c:200-327, copy-helper.c:5-40): renders a
 * w x h frame into the LINEAR pixel array (missed pixels = 0).  Optional
 * per-pixel outputs leaf_id / hit as in orc_ray_batch (only meaningful in
 * QUALITY_BEST, where each pixel is an independent root search). */
ORC_EXPORT int orc_render (orc_tree *t, orc_camera *c, int w, int h, unsigned quality,
                           const uint8_t *texture, const orc_pixfmt *fmt,
                           uint32_t *pixels, int32_t *leaf_id, float *hit,
                           int threads);

/* ---- lights (voxrnd/lights.c over voxtrees/mtree.c) ----
 * orc_lights_new: a light manager driven through the reference's own API -- reference glue only
 * (mtree.c / lights.c object code), the port returns NULL.  orc_lights_from_flat: a light set
 * given as flat arrays, evaluated by the restated traversal (both implementations). */
ORC_EXPORT orc_lights* orc_lights_new (void);
ORC_EXPORT orc_lights* orc_lights_from_flat (const orc_lights_flat *flat);
ORC_EXPORT void orc_lights_free (orc_lights *l);
ORC_EXPORT int orc_lights_insert (orc_lights *l, const float center[3], float radius, const float color[3]);
ORC_EXPORT int orc_lights_delete (orc_lights *l, const float center[3], float radius);
ORC_EXPORT void orc_lights_clear (orc_lights *l);
ORC_EXPORT int orc_lights_count (orc_lights *l);
ORC_EXPORT int orc_lights_set_ambient (orc_lights *l, const float color[3]);
/* sizes, then contents (arrays caller-allocated; NULL arrays = sizes + ambient only) */
ORC_EXPORT void orc_lights_flatten (orc_lights *l, orc_lights_flat *out);
/* vox_get_light (lights.c:104-116) for n points -> rgb[3n] */
ORC_EXPORT void orc_lights_get (orc_lights *l, const float *points, size_t n, float *rgb);
/* orc_render with a light manager attached (renderer.c:41-49) */
ORC_EXPORT int orc_render_lit (orc_tree *t, orc_camera *c, int w, int h, unsigned quality,
                               const uint8_t *texture, const orc_pixfmt *fmt, orc_lights *lights,
                               uint32_t *pixels, int threads);

#ifdef __cplusplus
}
#endif
#endif
