/**
   @file voxgpu.h
   @brief The thin C-ABI layer between the host C libraries and the CUDA
   kernels (libvoxgpu.so).  Plain pointers and sizes only.

   This is the drop-in boundary for the device side of the hot path:
     - voxgpu_tree_upload       <- the pointer tree built by vox_make_tree
                                   (reference: src/voxtrees/tree.c:249-338, node
                                   layout tree.h:22-38), flattened
     - voxgpu_rays_*            <- vox_ray_tree_intersection, src/voxtrees/search.c:8-153
     - voxgpu_balls_*           <- vox_tree_ball_collidep,    src/voxtrees/search.c:155-183
     - voxgpu_render            <- the vox_render block loop,  src/voxrnd/renderer.c:200-319
                                   with screen2world (simple-camera.c:21-31,
                                   doom-camera.c:36-57), get_color (renderer.c:27-52)
                                   and the de-tiling of vox_redraw (copy-helper.c:5-40)
                                   folded into one kernel.
   libvoxtrees / libvoxrnd call it; bench.py and multi-GPU drivers may call it
   directly with device buffers and their own streams.

   Every function returns 0 on success and a negative value on failure unless
   stated otherwise; voxgpu_last_error() then describes the failure.  There is
   no CPU fallback anywhere behind this interface.
**/
#ifndef VOXGPU_H
#define VOXGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VOXGPU_EXPORT __attribute__((visibility ("default")))

typedef struct voxgpu_tree voxgpu_tree;
typedef void* voxgpu_stream;   /* a cudaStream_t, or NULL for the default stream */

/* Node flags as in the reference (tree.h:15-18). */
#define VOXGPU_LEAF       1u
#define VOXGPU_DENSE_LEAF 2u

/** Host view of a flattened tree: breadth-first, the non-NULL children of an
    inner node consecutive in subspace order. Points are packed float[3]. */
typedef struct
{
    uint64_t n_nodes, n_dots;
    uint32_t depth;
    const float    *bb_min;      /* 3*n_nodes */
    const float    *bb_max;      /* 3*n_nodes */
    const uint32_t *flags;       /* n_nodes */
    const uint32_t *dots_num;    /* n_nodes */
    const float    *center;      /* 3*n_nodes */
    const uint32_t *first_child; /* n_nodes */
    const uint32_t *child_mask;  /* n_nodes */
    const uint64_t *dots_off;    /* n_nodes */
    const float    *dots;        /* 3*n_dots */
    float voxel[3];              /* vox_voxel at build time */
} voxgpu_flat;

/** Camera for device-side ray generation (see vox_camera_desc in voxrnd/camera.h). */
#define VOXGPU_CAMERA_SIMPLE 0
#define VOXGPU_CAMERA_DOOM   1
#define VOXGPU_CAMERA_TABLE  2   /* directions come from a device buffer */
typedef struct
{
    int kind;
    float position[3];
    float mul, xsub, ysub;
    float rotation[4];
    float sinphi, cosphi, k;
    const float *dir_table;      /* TABLE: device pointer, float4 per pixel, w*h entries */
} voxgpu_camera;

/** What SDL_MapRGB needs from a 32-bpp pixel format. */
typedef struct
{
    uint32_t amask;
    uint8_t rloss, gloss, bloss, pad0;
    uint8_t rshift, gshift, bshift, pad1;
} voxgpu_pixfmt;

/** Shadowless lights (reference: src/voxrnd/lights.c:91-116 over the M-tree of
    src/voxtrees/mtree.c), flattened by the host: M-tree nodes in PRE-ORDER, so that walking the
    array front to back, jumping to `skip` when a node's bounding sphere does not contain the
    point, visits the spheres in exactly the order of the reference's recursion
    (mtree.c:463-493) -- the order the fp32 light sum is taken in. */
typedef struct
{
    float ambient[3];
    uint32_t n_nodes, n_spheres;
    const float    *node_bounds;   /* 4*n_nodes: bounding sphere centre xyz, squared radius */
    const uint32_t *node_link;     /* 4*n_nodes: skip (index of the next node outside this subtree),
                                      leaf flag, first sphere (leaf), number of spheres (leaf) */
    const float    *sphere_geom;   /* 4*n_spheres: centre xyz, squared radius */
    const float    *sphere_light;  /* 4*n_spheres: colour rgb, radius */
} voxgpu_lights_flat;
typedef struct voxgpu_lights voxgpu_lights;

/** One render request: rows [y0, y1) of a w x h frame. */
typedef struct
{
    int w, h;
    int y0, y1;
    int pitch_pixels;            /* distance between rows of `pixels`, in pixels */
    unsigned quality;            /* VOX_QUALITY_* bits, renderer.h:81-127 */
    voxgpu_pixfmt fmt;
    const voxgpu_lights *lights; /* NULL: unlit grey (renderer.c:38-39); else renderer.c:41-49 */
    /* Multi-GPU screen split at the finest grain: with interleave_world = N > 1 the launch draws
       only the 4-row tile rows r (counted from y0) with r % N == interleave_rank and leaves every
       other pixel of the band untouched, so N GPUs writing into ONE frame buffer get statistically
       identical shares of the image in one launch each.  0 / 1: the whole band.  Quality BEST only. */
    int interleave_world, interleave_rank;
    /* The node the searches start from: 0 = the whole tree; the breadth-first id of an inner node or leaf
       renders that subtree only (the reference accepts any node as a scene: vox_context_set_scene with
       what vox_ray_tree_intersection returned, renderer.c:297). */
    uint32_t start_node;
} voxgpu_frame;

/** Counters of the instrumented ("stats") variant of the ray kernel. */
typedef struct
{
    unsigned long long rays, hits;
    unsigned long long calls;        /* node records fetched (non-NULL calls) */
    unsigned long long inner, leaf, dense;
    unsigned long long dots_tested;
    unsigned long long expansions;   /* inner nodes whose dividing planes were evaluated */
    unsigned long long max_stack;
} voxgpu_stats;

/* ---- device management ---- */
VOXGPU_EXPORT const char* voxgpu_last_error (void);
VOXGPU_EXPORT int  voxgpu_device_count (void);            /* 0 if none / no driver */
VOXGPU_EXPORT int  voxgpu_set_device (int device);
VOXGPU_EXPORT int  voxgpu_get_device (void);
VOXGPU_EXPORT int  voxgpu_sm_count (void);
VOXGPU_EXPORT int  voxgpu_synchronize (voxgpu_stream stream);

/* ---- memory helpers (so host C never needs cuda_runtime.h) ---- */
VOXGPU_EXPORT void* voxgpu_malloc (size_t bytes);                 /* device */
VOXGPU_EXPORT void  voxgpu_free (void *dptr);
VOXGPU_EXPORT void* voxgpu_host_alloc (size_t bytes);             /* pinned + mapped */
VOXGPU_EXPORT void  voxgpu_host_free (void *hptr);
VOXGPU_EXPORT int   voxgpu_host_register (void *hptr, size_t bytes);
VOXGPU_EXPORT int   voxgpu_host_unregister (void *hptr);
VOXGPU_EXPORT void* voxgpu_host_device_pointer (void *hptr);      /* device alias of mapped host memory */
VOXGPU_EXPORT int   voxgpu_memcpy_h2d (void *dptr, const void *hptr, size_t bytes, voxgpu_stream s);
VOXGPU_EXPORT int   voxgpu_memcpy_d2h (void *hptr, const void *dptr, size_t bytes, voxgpu_stream s);
VOXGPU_EXPORT int   voxgpu_memset (void *dptr, int value, size_t bytes, voxgpu_stream s);
VOXGPU_EXPORT voxgpu_stream voxgpu_stream_create (void);
VOXGPU_EXPORT void  voxgpu_stream_destroy (voxgpu_stream s);
/** Work queued on `waiter` after this call starts only when everything queued on `other` so far is done. */
VOXGPU_EXPORT int   voxgpu_stream_join (voxgpu_stream waiter, voxgpu_stream other);

/* ---- peer memory: another process's device buffer as a render target ----
 * export on the owner (memory from voxgpu_malloc), ship the 64 bytes to the other
 * process by any means, open there; the returned address is valid as d_pixels. */
#define VOXGPU_IPC_HANDLE_BYTES 64
VOXGPU_EXPORT int   voxgpu_ipc_export (const void *dptr, unsigned char handle[VOXGPU_IPC_HANDLE_BYTES]);
VOXGPU_EXPORT void* voxgpu_ipc_open (const unsigned char handle[VOXGPU_IPC_HANDLE_BYTES]);
VOXGPU_EXPORT int   voxgpu_ipc_close (void *dptr);
/** One process, several devices: kernels of the current device may then store into `peer_device`'s memory. */
VOXGPU_EXPORT int   voxgpu_enable_peer (int peer_device);

/* ---- tree ---- */
/** Convert to the device layout and upload (cudaMemcpyAsync from pinned staging) on the current device. */
VOXGPU_EXPORT voxgpu_tree* voxgpu_tree_upload (const voxgpu_flat *flat);
VOXGPU_EXPORT void voxgpu_tree_free (voxgpu_tree *tree);
VOXGPU_EXPORT size_t voxgpu_tree_device_bytes (const voxgpu_tree *tree);
VOXGPU_EXPORT int voxgpu_tree_device (const voxgpu_tree *tree);

/* ---- lights ---- */
/** Copy a flattened light set to the current device (one small synchronous upload). */
VOXGPU_EXPORT voxgpu_lights* voxgpu_lights_upload (const voxgpu_lights_flat *flat);
VOXGPU_EXPORT void voxgpu_lights_free (voxgpu_lights *lights);

/* ---- batched ray queries (K2) ----
 * origins/dirs: float4 per ray (xyz + unused lane = vox_dot).
 * leaf_ids: int32 per ray (-1 = miss); points: float4 per ray (written on a hit);
 * voxels: int32 per ray (index into the flat dots, -1 = dense leaf or miss).
 * Any output may be NULL. */
VOXGPU_EXPORT int voxgpu_rays_device (const voxgpu_tree *tree, uint32_t start_node,
                                      const float *d_origins, const float *d_dirs, size_t n,
                                      int32_t *d_leaf_ids, float *d_points, int32_t *d_voxels,
                                      voxgpu_stream stream);
/** Order in which voxgpu_rays_device searches a batch: -1 (default; $VOXGPU_RAYS_SORT sets the start-up value)
    = batches of 262 144 rays and more in the order of the Morton code of the point where each ray enters the
    start node's box (two-pass radix sort of the ray indices on the device, results scattered back to the
    caller's positions), smaller ones as given; 0 = always as given; 1 = always in entry order.  Results do
    not depend on it: the search of a ray is a function of the ray. */
VOXGPU_EXPORT void voxgpu_set_rays_sort (int mode);
/** Batches searched in entry order run on persistent warps whose lanes take a new ray as soon as this many
    lanes of the warp are idle (1..32; default 8, $VOXGPU_RAYS_REFILL sets the start-up value); 0 = one ray per
    thread for the lifetime of the warp.  Results do not depend on it. */
VOXGPU_EXPORT void voxgpu_set_rays_refill (int lanes);
/** Host buffers: copies in, runs, copies out, synchronises. n_hits may be NULL. */
VOXGPU_EXPORT int voxgpu_rays_host (const voxgpu_tree *tree, uint32_t start_node,
                                    const float *origins, const float *dirs, size_t n,
                                    int32_t *leaf_ids, float *points, int32_t *voxels,
                                    long *n_hits);
/** Instrumented variant (device buffers, synchronous): same results + counters. */
VOXGPU_EXPORT int voxgpu_rays_stats (const voxgpu_tree *tree, uint32_t start_node,
                                     const float *d_origins, const float *d_dirs, size_t n,
                                     int32_t *d_leaf_ids, float *d_points, int32_t *d_voxels,
                                     voxgpu_stats *stats);

/* ---- raw volume -> voxel list (reference: src/voxtrees/datareader.c:51-76) ----
 * `volume`: dim[0]*dim[1]*dim[2] samples of `samplesize` (1..4) bytes, little-endian, x outer / z inner, in HOST
 * memory.  Keeps the samples with min <= sample < max_excl (datareader.c's Block reduced to the range
 * predicate of the Lua binding, src/voxengine/modules/voxtrees.c:306-342; "sample > t" is [t+1, 2^32)).
 * The volume is uploaded, thresholded and stream-compacted on the device in the reference's order; on
 * success *count voxels {i*voxel[0], j*voxel[1], k*voxel[2], 0} (float4 = vox_dot) are in *dots, page-locked
 * host memory the caller releases with voxgpu_host_free (NULL when *count == 0). */
VOXGPU_EXPORT int voxgpu_volume_extract (const void *volume, const unsigned dim[3], unsigned samplesize,
                                         unsigned min, unsigned long long max_excl, const float voxel[3],
                                         float **dots, size_t *count);

/* ---- batched ball queries (K3) ---- centers: float4 per ball; out: 1 byte per ball. */
VOXGPU_EXPORT int voxgpu_balls_device (const voxgpu_tree *tree, const float *d_centers,
                                       const float *d_radii, size_t n, uint8_t *d_out,
                                       voxgpu_stream stream);
VOXGPU_EXPORT int voxgpu_balls_host (const voxgpu_tree *tree, const float *centers,
                                     const float *radii, size_t n, uint8_t *out, long *n_hits);

/* ---- primary-ray render (K1) ----
 * d_texture: 32*32*32 bytes on the device; d_pixels: device-accessible pointer
 * (device memory, or the device alias of mapped pinned host memory) to pixel
 * (0, y0) -- i.e. the first row of the requested band.  Every pixel of the
 * band is written (0 where nothing is hit).  d_leaf_ids / d_points (optional,
 * same indexing as pixels but dense w-pitch, int32 / float4 per pixel) expose the
 * per-pixel search result for parity tests. */
VOXGPU_EXPORT int voxgpu_render (const voxgpu_tree *tree, const voxgpu_camera *camera,
                                 const voxgpu_frame *frame, const uint8_t *d_texture,
                                 uint32_t *d_pixels, int32_t *d_leaf_ids, float *d_points,
                                 voxgpu_stream stream);

/* ---- frames to host memory ----
 * What vox_render + vox_redraw (renderer.c:200-327) use when the surface is in
 * host memory: the frame is rendered into a device framebuffer in `bands`
 * horizontal bands (0 = default, or $VOXGPU_BANDS), each on its own prioritised
 * stream and each followed by a DMA copy into `host_pixels`, which must be
 * page-locked (voxgpu_host_alloc / voxgpu_host_register).  _render returns as
 * soon as the work is queued (and the margin of the surface outside the root
 * box's screen bounds, which no voxel can reach, is zeroed by the calling
 * thread: only the visible rectangle is rendered and copied); _wait returns
 * once every band has landed.
 * frame->y0/y1/pitch_pixels are ignored (the whole frame is drawn) -- or, with
 * frame->interleave_world = N > 1, this device's share of a frame that N devices
 * split: the tile rows r with r % N == interleave_rank, one launch, one strided
 * DMA copy of exactly those rows (the columns of the visible rectangle) into
 * `host_pixels` (needs host_pitch_bytes == 4*w; the margin is the caller's to
 * zero, once for all shares: voxgpu_multi_render does). */
typedef struct voxgpu_present voxgpu_present;
VOXGPU_EXPORT voxgpu_present* voxgpu_present_create (int w, int h, int bands);
VOXGPU_EXPORT void voxgpu_present_destroy (voxgpu_present *p);
VOXGPU_EXPORT int voxgpu_present_bands (const voxgpu_present *p);
VOXGPU_EXPORT int voxgpu_present_render (voxgpu_present *p, const voxgpu_tree *tree, const voxgpu_camera *camera,
                                         const voxgpu_frame *frame, const uint8_t *d_texture,
                                         void *host_pixels, size_t host_pitch_bytes);
VOXGPU_EXPORT int voxgpu_present_wait (voxgpu_present *p);
/** Device time from the first band's launch to the last band's arrival in host memory, ms (waits). */
VOXGPU_EXPORT float voxgpu_present_last_ms (voxgpu_present *p);
/** Bytes of frames all presenters of this process have queued for device-to-host copies so far (only the visible
    rectangle of a frame crosses PCIe; the margin is zeroed on the host). */
VOXGPU_EXPORT unsigned long long voxgpu_present_bytes_copied (void);
/** The presenter's host arithmetic for the visible spans, on its own (no device is touched): `boxes` is six
    arrays of n floats (min x, y, z, max x, y, z; n <= 512) that together hold every voxel; on return 1,
    a voxel can only show in row y at columns lo[y/4] <= x < hi[y/4] (hi <= lo: the rows show nothing); 0: no
    claim (a box beside or behind the camera).  lo, hi: h/4 ints each.  The presenter copies only those columns
    (in strips, on 16-pixel boundaries) and zeroes the rest of the surface on the host. */
/** Knobs of the above, process-wide (a negative / zero argument leaves that knob alone): spans on or off (off: the
    root box's rectangle only; default on, $VOXGPU_SPANS), the most copies one frame is cut into (default 16,
    $VOXGPU_COPY_STRIPS), and what one more copy must save in bytes to be worth it (default 98304, $VOXGPU_COPY_COST). */
VOXGPU_EXPORT void voxgpu_set_present_spans (int on, int max_strips, long long copy_cost);
/** What voxgpu_present_render would copy for a frame of a tree whose coarse cut is `boxes`, drawn in `nbands` bands
    (host arithmetic only, for tests): per 4-row unit u the columns [keep_lo[u], keep_hi[u]) the copies write -- the
    rest of the surface is zeroed on the host.  -> bytes copied; -1: the camera cannot see the boxes (everything is
    zeroed); -2: bad argument. */
VOXGPU_EXPORT long long voxgpu_present_plan (const float *boxes, unsigned n, const voxgpu_camera *c, int w, int h, int nbands,
                                             int *keep_lo, int *keep_hi);
VOXGPU_EXPORT int voxgpu_visible_spans (const float *boxes, unsigned n, const voxgpu_camera *c, int w, int h, int *lo, int *hi);
/** The device framebuffer the last frame was rendered into (w*h pixels, pitch w). */
VOXGPU_EXPORT const uint32_t* voxgpu_present_device_frame (const voxgpu_present *p);

/* ---- one frame on all GPUs of the process ----
 * The in-process counterpart of the reference's fan-out inside vox_render (dispatch_apply over the
 * blocks, renderer.c:214-226): device i of `devices` draws the tile rows r with r % ndev == i and
 * DMA-copies them into the page-locked `host_pixels` over its own PCIe link; one worker thread per
 * device queues its share.  trees[i] / d_tex[i] / lights[i] are the replicas on devices[i] (lights may
 * be NULL).  Quality BEST only.  _render returns when every share is queued, _wait when all have landed. */
typedef struct voxgpu_multi voxgpu_multi;
VOXGPU_EXPORT voxgpu_multi* voxgpu_multi_create (int w, int h, int ndev, const int *devices);
VOXGPU_EXPORT void voxgpu_multi_destroy (voxgpu_multi *m);
VOXGPU_EXPORT int voxgpu_multi_devices (const voxgpu_multi *m);
VOXGPU_EXPORT int voxgpu_multi_device (const voxgpu_multi *m, int i);
VOXGPU_EXPORT int voxgpu_multi_render (voxgpu_multi *m, const voxgpu_tree *const *trees, const voxgpu_camera *camera,
                                       const voxgpu_frame *frame, const uint8_t *const *d_tex,
                                       const voxgpu_lights *const *lights, void *host_pixels, size_t host_pitch_bytes);
VOXGPU_EXPORT int voxgpu_multi_wait (voxgpu_multi *m);
/** Slowest device's time from its launch to the arrival of its share in host memory, ms (waits). */
VOXGPU_EXPORT float voxgpu_multi_last_ms (voxgpu_multi *m);

/** Development knob: 1 = strip-interleaved tile order where a launch has no screen mask (default), 2 = raster
    order.  Kept for A/B measurements; results are identical. */
VOXGPU_EXPORT void voxgpu_set_variant (int variant);
/** Tile cost feedback: every traversed 8x4 tile of a BEST frame records the node visits of its slowest lane in a map
    kept with the device tree; the next frame of the same geometry hands out first the marked tiles in whose 3x3
    neighbourhood that map holds at least `heavy_min` visits (default 48; 0 = off; $VOXGPU_TILE_FEEDBACK sets the
    start-up value).  A pure scheduling hint for frame sequences: results do not depend on it. */
VOXGPU_EXPORT void voxgpu_set_tile_feedback (int heavy_min);

/** Screen-space culling of 8x4 tiles that cannot show a voxel.  1 (default): the bounds of the tree's
    root box AND a per-launch screen mask marked from the boxes of a cut through the tree (k_mark);
    2: the root-box bounds only; 0: none ($VOXGPU_NO_CULL=1, $VOXGPU_NO_MASK=1 set the start-up value).
    Results do not depend on it: it only skips rays whose search would come back empty.  Kept
    switchable for A/B runs and for the tests that prove the independence. */
VOXGPU_EXPORT void voxgpu_set_cull (int on);
/** The bounds that culling uses (host arithmetic, no device needed): out = {x0, x1, y0, y1}, pixels outside
    [x0,x1) x [y0,y1) of a w x h frame cannot see the box; the whole frame if the camera gives no bounds. */
VOXGPU_EXPORT void voxgpu_screen_bounds (const float root_min[3], const float root_max[3], const voxgpu_camera *camera,
                                         int w, int h, int out[4]);

/** Development trace: when non-NULL, every following unlit BEST voxgpu_render launch records per
    traversed 8x4 tile {start ns, end ns, smid<<32 | node visits of the slowest lane} (3 x uint64 per
    tile of the WHOLE frame, device memory supplied by the caller).  NULL switches it off. */
VOXGPU_EXPORT void voxgpu_set_trace (void *d_trace);

/** Number of kernels this library has launched in this process (for bench.py's gpu_launches). */
VOXGPU_EXPORT unsigned long long voxgpu_launch_count (void);

/** CUDA-event timing helpers: elapsed milliseconds between two marks on a stream. */
typedef struct voxgpu_timer voxgpu_timer;
VOXGPU_EXPORT voxgpu_timer* voxgpu_timer_create (void);
VOXGPU_EXPORT void  voxgpu_timer_destroy (voxgpu_timer *t);
VOXGPU_EXPORT int   voxgpu_timer_start (voxgpu_timer *t, voxgpu_stream s);
VOXGPU_EXPORT int   voxgpu_timer_stop (voxgpu_timer *t, voxgpu_stream s);
VOXGPU_EXPORT float voxgpu_timer_elapsed_ms (voxgpu_timer *t);   /* synchronises on the stop mark */

#ifdef __cplusplus
}
#endif
#endif
