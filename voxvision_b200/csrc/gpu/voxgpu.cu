/*
 * voxgpu.cu -- CUDA kernels (sm_100a) and the thin C-ABI of include/voxgpu.h.
 *
 *   K1 k_render : camera ray generation (simple-camera.c:21-31 / doom-camera.c:36-57)
 *                 + first-hit octree search (search.c:8-153) + get_color
 *                 (renderer.c:27-52) + linear framebuffer store, one thread per
 *                 pixel, one warp per 8x4 pixel tile.  Replaces the 4x4-block
 *                 dispatch_apply loop of renderer.c:214-318 and copy_squares
 *                 (copy-helper.c:5-40).
 *   K2 k_rays   : the same search for an array of independent rays.
 *   K3 k_balls  : vox_tree_ball_collidep (search.c:155-183) for an array of balls.
 *
 * Compile with -fmad=false (see traverse.cuh for why).
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <math.h>
#include <atomic>
#include <mutex>

#include "voxgpu.h"
#include "traverse.cuh"

/* ------------------------------------------------------------------ errors */

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

static int fail (const char *fmt, ...)
{
    va_list ap;
    va_start (ap, fmt);
    vsnprintf (g_err, sizeof g_err, fmt, ap);
    va_end (ap);
    return -1;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail ("%s failed: %s (%s:%d)", #call, cudaGetErrorString (e_), __FILE__, __LINE__); } while (0)
#define CU_NULL(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fail ("%s failed: %s (%s:%d)", #call, cudaGetErrorString (e_), __FILE__, __LINE__); return NULL; } } while (0)

extern "C" const char* voxgpu_last_error (void) { return g_err; }
extern "C" unsigned long long voxgpu_launch_count (void) { return g_launches.load (); }

/* ------------------------------------------------------------------ device */

extern "C" int voxgpu_device_count (void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount (&n);
    if (e != cudaSuccess) { fail ("cudaGetDeviceCount: %s", cudaGetErrorString (e)); return 0; }
    return n;
}
extern "C" int voxgpu_set_device (int device) { CU (cudaSetDevice (device)); return 0; }
extern "C" int voxgpu_get_device (void) { int d = -1; if (cudaGetDevice (&d) != cudaSuccess) return -1; return d; }
extern "C" int voxgpu_sm_count (void)
{
    int d = 0, n = 0;
    CU (cudaGetDevice (&d));
    CU (cudaDeviceGetAttribute (&n, cudaDevAttrMultiProcessorCount, d));
    return n;
}
extern "C" int voxgpu_synchronize (voxgpu_stream s)
{
    if (s) CU (cudaStreamSynchronize ((cudaStream_t) s));
    else CU (cudaDeviceSynchronize ());
    return 0;
}

extern "C" void* voxgpu_malloc (size_t bytes)
{
    void *p = NULL;
    CU_NULL (cudaMalloc (&p, bytes ? bytes : 16));
    return p;
}
extern "C" void voxgpu_free (void *p) { if (p) cudaFree (p); }
extern "C" void* voxgpu_host_alloc (size_t bytes)
{
    void *p = NULL;
    CU_NULL (cudaHostAlloc (&p, bytes ? bytes : 16, cudaHostAllocMapped | cudaHostAllocPortable));
    return p;
}
extern "C" void voxgpu_host_free (void *p) { if (p) cudaFreeHost (p); }
extern "C" int voxgpu_host_register (void *p, size_t bytes)
{
    CU (cudaHostRegister (p, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    return 0;
}
extern "C" int voxgpu_host_unregister (void *p) { CU (cudaHostUnregister (p)); return 0; }
extern "C" void* voxgpu_host_device_pointer (void *p)
{
    void *d = NULL;
    CU_NULL (cudaHostGetDevicePointer (&d, p, 0));
    return d;
}
extern "C" int voxgpu_memcpy_h2d (void *d, const void *h, size_t n, voxgpu_stream s)
{ CU (cudaMemcpyAsync (d, h, n, cudaMemcpyHostToDevice, (cudaStream_t) s)); return 0; }
extern "C" int voxgpu_memcpy_d2h (void *h, const void *d, size_t n, voxgpu_stream s)
{ CU (cudaMemcpyAsync (h, d, n, cudaMemcpyDeviceToHost, (cudaStream_t) s)); return 0; }
extern "C" int voxgpu_memset (void *d, int v, size_t n, voxgpu_stream s)
{ CU (cudaMemsetAsync (d, v, n, (cudaStream_t) s)); return 0; }
extern "C" voxgpu_stream voxgpu_stream_create (void)
{
    cudaStream_t s = NULL;
    CU_NULL (cudaStreamCreateWithFlags (&s, cudaStreamNonBlocking));
    return (voxgpu_stream) s;
}
extern "C" void voxgpu_stream_destroy (voxgpu_stream s) { if (s) cudaStreamDestroy ((cudaStream_t) s); }
extern "C" int voxgpu_stream_join (voxgpu_stream waiter, voxgpu_stream other)
{
    /* everything queued on `other` so far must finish before what is queued on `waiter` next */
    cudaEvent_t e;
    CU (cudaEventCreateWithFlags (&e, cudaEventDisableTiming));
    cudaError_t rc = cudaEventRecord (e, (cudaStream_t) other);
    if (rc == cudaSuccess) rc = cudaStreamWaitEvent ((cudaStream_t) waiter, e, 0);
    cudaEventDestroy (e);
    if (rc != cudaSuccess) return fail ("voxgpu_stream_join: %s", cudaGetErrorString (rc));
    return 0;
}

/* ------------------------------------------------ peer memory (multi-GPU) */

/* One process per GPU: rank 0 exports a cudaMalloc'ed buffer, the other ranks map it and
   hand the mapped address to voxgpu_render as the framebuffer -- the render kernel then
   stores its pixels straight into GPU 0's memory over NVLink (the "gather" of the
   multi-GPU path, fused into K1's store). */
extern "C" int voxgpu_ipc_export (const void *dptr, unsigned char handle[VOXGPU_IPC_HANDLE_BYTES])
{
    static_assert (sizeof (cudaIpcMemHandle_t) == VOXGPU_IPC_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    CU (cudaIpcGetMemHandle (&h, (void*) dptr));
    memcpy (handle, &h, sizeof h);
    return 0;
}
extern "C" void* voxgpu_ipc_open (const unsigned char handle[VOXGPU_IPC_HANDLE_BYTES])
{
    cudaIpcMemHandle_t h;
    void *p = NULL;
    memcpy (&h, handle, sizeof h);
    CU_NULL (cudaIpcOpenMemHandle (&p, h, cudaIpcMemLazyEnablePeerAccess));
    return p;
}
extern "C" int voxgpu_ipc_close (void *dptr) { CU (cudaIpcCloseMemHandle (dptr)); return 0; }

/* ------------------------------------------------------------------- timer */

struct voxgpu_timer { cudaEvent_t a, b; };
extern "C" voxgpu_timer* voxgpu_timer_create (void)
{
    voxgpu_timer *t = (voxgpu_timer*) calloc (1, sizeof *t);
    if (cudaEventCreate (&t->a) != cudaSuccess || cudaEventCreate (&t->b) != cudaSuccess)
    { fail ("cudaEventCreate failed"); free (t); return NULL; }
    return t;
}
extern "C" void voxgpu_timer_destroy (voxgpu_timer *t)
{ if (!t) return; cudaEventDestroy (t->a); cudaEventDestroy (t->b); free (t); }
extern "C" int voxgpu_timer_start (voxgpu_timer *t, voxgpu_stream s) { CU (cudaEventRecord (t->a, (cudaStream_t) s)); return 0; }
extern "C" int voxgpu_timer_stop (voxgpu_timer *t, voxgpu_stream s) { CU (cudaEventRecord (t->b, (cudaStream_t) s)); return 0; }
extern "C" float voxgpu_timer_elapsed_ms (voxgpu_timer *t)
{
    float ms = -1.0f;
    if (cudaEventSynchronize (t->b) != cudaSuccess) { fail ("cudaEventSynchronize failed"); return -1.0f; }
    if (cudaEventElapsedTime (&ms, t->a, t->b) != cudaSuccess) { fail ("cudaEventElapsedTime failed"); return -1.0f; }
    return ms;
}

/* -------------------------------------------------------------------- tree */

struct voxgpu_tree
{
    int device;
    uint64_t n_nodes, n_dots;
    uint32_t depth;
    NodeRec *rec;    /* 32 B per node: box + meta + link */
    float4 *ctr;     /* 16 B per node: split centre (inner nodes) */
    float4 *dots;
    float voxel[3];
    float root_min[3], root_max[3];   /* host copy of node 0's box, for screen-space culling */
    /* a cut through the tree for the screen mask (k_mark): every node of one level plus the leaves
       above it, so that every voxel lies in the box of exactly one node of the cut */
    unsigned *cut;
    unsigned n_cut, cut_level;
};
#define VOXGPU_CUT_MAX 131072u

extern "C" voxgpu_tree* voxgpu_tree_upload (const voxgpu_flat *f)
{
    if (!f || f->n_nodes == 0) { fail ("voxgpu_tree_upload: empty tree"); return NULL; }
    if (f->n_nodes >= 0x7fffffffull || f->n_dots >= 0x7fffffffull)
    { fail ("voxgpu_tree_upload: tree too large for 31-bit ids (%llu nodes, %llu dots)",
            (unsigned long long) f->n_nodes, (unsigned long long) f->n_dots); return NULL; }
    if (f->depth > VOXGPU_MAX_DEPTH)
    { fail ("voxgpu_tree_upload: tree depth %u exceeds the device stack (%d levels)", f->depth, VOXGPU_MAX_DEPTH); return NULL; }

    voxgpu_tree *t = (voxgpu_tree*) calloc (1, sizeof *t);
    t->n_nodes = f->n_nodes; t->n_dots = f->n_dots; t->depth = f->depth;
    memcpy (t->root_min, f->bb_min, sizeof t->root_min);
    memcpy (t->root_max, f->bb_max, sizeof t->root_max);
    memcpy (t->voxel, f->voxel, sizeof t->voxel);
    if (cudaGetDevice (&t->device) != cudaSuccess) { fail ("no CUDA device: %s", cudaGetErrorString (cudaGetLastError ())); free (t); return NULL; }

    const size_t rec_bytes = (size_t) f->n_nodes * sizeof (NodeRec);
    const size_t ctr_bytes = (size_t) f->n_nodes * sizeof (float4);
    const size_t dot_bytes = (size_t) (f->n_dots ? f->n_dots : 1) * sizeof (float4);
    size_t stage_bytes = rec_bytes > dot_bytes ? rec_bytes : dot_bytes;
    void *stage_v = NULL;
    cudaStream_t s = NULL;
    cudaError_t e;
    if ((e = cudaMalloc (&t->rec, rec_bytes)) != cudaSuccess ||
        (e = cudaMalloc (&t->ctr, ctr_bytes)) != cudaSuccess ||
        (e = cudaMalloc (&t->dots, dot_bytes)) != cudaSuccess ||
        (e = cudaHostAlloc (&stage_v, stage_bytes, cudaHostAllocDefault)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags (&s, cudaStreamNonBlocking)) != cudaSuccess)
    {
        fail ("voxgpu_tree_upload: %s", cudaGetErrorString (e));
        goto bad;
    }

    /* device records, written straight into the pinned staging buffer */
    {
        NodeRec *stage = (NodeRec*) stage_v;
        for (uint64_t id=0; id<f->n_nodes; id++)
        {
            const uint32_t fl = f->flags[id];
            uint32_t meta, link;
            if (fl & VOXGPU_DENSE_LEAF) { meta = VOXGPU_KIND_DENSE; link = 0; }
            else if (fl & VOXGPU_LEAF)
            {
                if (f->dots_num[id] >= (1u << 30)) { fail ("leaf with %u dots", f->dots_num[id]); goto bad; }
                meta = VOXGPU_KIND_LEAF | (f->dots_num[id] << 2);
                link = (uint32_t) f->dots_off[id];
            }
            else { meta = VOXGPU_KIND_INNER | ((f->child_mask[id] & 0xffu) << 8); link = f->first_child[id]; }
            NodeRec r;
            r.lo.x = f->bb_min[3*id]; r.lo.y = f->bb_min[3*id+1]; r.lo.z = f->bb_min[3*id+2];
            memcpy (&r.lo.w, &meta, 4);
            r.hi.x = f->bb_max[3*id]; r.hi.y = f->bb_max[3*id+1]; r.hi.z = f->bb_max[3*id+2];
            memcpy (&r.hi.w, &link, 4);
            stage[id] = r;
        }
    }
    if ((e = cudaMemcpyAsync (t->rec, stage_v, rec_bytes, cudaMemcpyHostToDevice, s)) != cudaSuccess ||
        (e = cudaStreamSynchronize (s)) != cudaSuccess)
    { fail ("voxgpu_tree_upload: node copy: %s", cudaGetErrorString (e)); goto bad; }
    {
        float4 *stage = (float4*) stage_v;
        for (uint64_t id=0; id<f->n_nodes; id++)
        {
            float4 q; q.x = f->center[3*id]; q.y = f->center[3*id+1]; q.z = f->center[3*id+2]; q.w = 0;
            stage[id] = q;
        }
        if ((e = cudaMemcpyAsync (t->ctr, stage_v, ctr_bytes, cudaMemcpyHostToDevice, s)) != cudaSuccess ||
            (e = cudaStreamSynchronize (s)) != cudaSuccess)
        { fail ("voxgpu_tree_upload: centre copy: %s", cudaGetErrorString (e)); goto bad; }
        for (uint64_t i=0; i<f->n_dots; i++)
        {
            float4 q; q.x = f->dots[3*i]; q.y = f->dots[3*i+1]; q.z = f->dots[3*i+2]; q.w = 0;
            stage[i] = q;
        }
        if (f->n_dots &&
            ((e = cudaMemcpyAsync (t->dots, stage_v, f->n_dots * sizeof (float4), cudaMemcpyHostToDevice, s)) != cudaSuccess ||
             (e = cudaStreamSynchronize (s)) != cudaSuccess))
        { fail ("voxgpu_tree_upload: dots copy: %s", cudaGetErrorString (e)); goto bad; }
    }
    {
        /* levels from the breadth-first links; the deepest level whose cut stays below VOXGPU_CUT_MAX nodes */
        unsigned char *level = (unsigned char*) malloc (f->n_nodes);
        uint64_t at[VOXGPU_MAX_DEPTH + 2] = {0}, leaves[VOXGPU_MAX_DEPTH + 2] = {0};
        level[0] = 0;
        for (uint64_t id=0; id<f->n_nodes; id++)
        {
            const unsigned l = level[id];
            at[l]++;
            if (f->flags[id] & 3u) { leaves[l]++; continue; }
            const int nc = __builtin_popcount (f->child_mask[id] & 0xffu);
            for (int j=0; j<nc; j++) level[f->first_child[id] + j] = (unsigned char) (l + 1);
        }
        unsigned best = 0;
        uint64_t above = 0;
        for (unsigned l=0; l<f->depth && l <= VOXGPU_MAX_DEPTH; l++)
        {
            if (at[l] + above <= VOXGPU_CUT_MAX) best = l;
            above += leaves[l];
        }
        unsigned n = 0;
        unsigned *tmp = (unsigned*) malloc ((size_t) VOXGPU_CUT_MAX * sizeof (unsigned));
        for (uint64_t id=0; id<f->n_nodes && n < VOXGPU_CUT_MAX; id++)
            if (level[id] == best || (level[id] < best && (f->flags[id] & 3u))) tmp[n++] = (unsigned) id;
        free (level);
        t->n_cut = n; t->cut_level = best;
        if ((e = cudaMalloc (&t->cut, (size_t) (n ? n : 1) * sizeof (unsigned))) != cudaSuccess ||
            (e = cudaMemcpy (t->cut, tmp, (size_t) n * sizeof (unsigned), cudaMemcpyHostToDevice)) != cudaSuccess)
        { free (tmp); fail ("voxgpu_tree_upload: cut copy: %s", cudaGetErrorString (e)); goto bad; }
        free (tmp);
    }
    cudaFreeHost (stage_v);
    cudaStreamDestroy (s);
    return t;
bad:
    if (stage_v) cudaFreeHost (stage_v);
    if (s) cudaStreamDestroy (s);
    cudaFree (t->rec); cudaFree (t->ctr); cudaFree (t->dots); cudaFree (t->cut);
    free (t);
    return NULL;
}

extern "C" void voxgpu_tree_free (voxgpu_tree *t)
{
    if (!t) return;
    int cur = -1;
    cudaGetDevice (&cur);
    if (cur != t->device) cudaSetDevice (t->device);
    cudaFree (t->rec); cudaFree (t->ctr); cudaFree (t->dots); cudaFree (t->cut);
    if (cur != t->device && cur >= 0) cudaSetDevice (cur);
    free (t);
}
extern "C" size_t voxgpu_tree_device_bytes (const voxgpu_tree *t)
{ return t ? (size_t) t->n_nodes * 48 + (size_t) t->n_dots * 16 : 0; }
extern "C" int voxgpu_tree_device (const voxgpu_tree *t) { return t ? t->device : -1; }

static TreeRef tree_ref (const voxgpu_tree *t)
{
    TreeRef r;
    r.rec = t->rec; r.ctr = t->ctr; r.dots = t->dots;
    r.vx = t->voxel[0]; r.vy = t->voxel[1]; r.vz = t->voxel[2];
    r.top_rec = r.top_ctr = NULL;
    r.ntop = 0;
    r.ntop_avail = 0;
#if VOXGPU_SMEM_TOP
    r.ntop_avail = t->n_nodes < VOXGPU_SMEM_TOP ? (unsigned) t->n_nodes : VOXGPU_SMEM_TOP;
#endif
    return r;
}

static int check_tree (const voxgpu_tree *t, uint32_t start_node, const char *who)
{
    if (!t) return fail ("%s: NULL device tree", who);
    if (start_node >= t->n_nodes) return fail ("%s: start node %u out of range", who, start_node);
    int cur = -1;
    CU (cudaGetDevice (&cur));
    if (cur != t->device) return fail ("%s: tree lives on device %d, current device is %d", who, t->device, cur);
    return 0;
}

/* ---------------------------------------------------------------- K2: rays */

#define RAYS_BLOCK 128

/* Resident CTAs per SM for K2.  C3 (tools/ab_rays.sh): 7 (66 registers) 1446 Mrays/s, 8 (64 registers,
   40 bytes spilled) 1558, 9 1493, 10 1503, 12 1414: the random rays of C3 wait on HBM, one more CTA of
   latency hiding pays for the spill. */
#ifndef VOXGPU_RAYS_MIN_BLOCKS
#define VOXGPU_RAYS_MIN_BLOCKS 8
#endif
template <bool STATS>
__global__ void __launch_bounds__(RAYS_BLOCK, VOXGPU_RAYS_MIN_BLOCKS)
k_rays (TreeRef tr, unsigned start, const float4 * __restrict__ origins, const float4 * __restrict__ dirs,
        size_t n, int * __restrict__ leaf_ids, float4 * __restrict__ points, int * __restrict__ voxels,
        unsigned long long *stats)
{
    const size_t i = (size_t) blockIdx.x * RAYS_BLOCK + threadIdx.x;
    const bool live = i < n;            /* the production traversal is warp-synchronous: nobody leaves early */
    float4 stack[VOXGPU_RAY_STACK];
    const float4 o = live ? __ldg (origins + i) : make_float4 (0.0f, 0.0f, 0.0f, 0.0f);
    const float4 d = live ? __ldg (dirs + i) : make_float4 (1.0f, 1.0f, 1.0f, 0.0f);
    RayResult r;
    RayCounters c = {0, 0, 0, 0, 0, 0, 0};
    if (STATS) { if (live) ray_tree<true> (tr, start, v3 (o), v3 (d), stack, r, &c); }
    else ray_tree_warp (tr, start, v3 (o), v3 (d), live, stack, r);
    if (!live) return;
    if (leaf_ids) leaf_ids[i] = r.leaf;
    if (voxels) voxels[i] = r.voxel;
    if (points && r.leaf >= 0) points[i] = make_float4 (r.point.x, r.point.y, r.point.z, 0.0f);
    if (STATS)
    {
        atomicAdd (stats + 0, 1ull);
        atomicAdd (stats + 1, r.leaf >= 0 ? 1ull : 0ull);
        atomicAdd (stats + 2, (unsigned long long) c.calls);
        atomicAdd (stats + 3, (unsigned long long) c.inner);
        atomicAdd (stats + 4, (unsigned long long) c.leaf);
        atomicAdd (stats + 5, (unsigned long long) c.dense);
        atomicAdd (stats + 6, (unsigned long long) c.dots);
        atomicAdd (stats + 7, (unsigned long long) c.expansions);
        atomicMax (stats + 8, (unsigned long long) c.max_stack);
    }
}

extern "C" int voxgpu_rays_device (const voxgpu_tree *t, uint32_t start, const float *d_o, const float *d_d,
                                   size_t n, int32_t *d_leaf, float *d_pts, int32_t *d_vox, voxgpu_stream s)
{
    if (check_tree (t, start, "voxgpu_rays_device")) return -1;
    if (n == 0) return 0;
    const size_t blocks = (n + RAYS_BLOCK - 1) / RAYS_BLOCK;
    if (blocks > 0x7fffffffull) return fail ("voxgpu_rays_device: too many rays");
    k_rays<false><<<(unsigned) blocks, RAYS_BLOCK, 0, (cudaStream_t) s>>> (
        tree_ref (t), start, (const float4*) d_o, (const float4*) d_d, n, d_leaf, (float4*) d_pts, d_vox, NULL);
    g_launches++;
    CU (cudaGetLastError ());
    return 0;
}

extern "C" int voxgpu_rays_stats (const voxgpu_tree *t, uint32_t start, const float *d_o, const float *d_d,
                                  size_t n, int32_t *d_leaf, float *d_pts, int32_t *d_vox, voxgpu_stats *out)
{
    if (check_tree (t, start, "voxgpu_rays_stats")) return -1;
    unsigned long long *d_st = NULL, h[9];
    CU (cudaMalloc (&d_st, sizeof h));
    CU (cudaMemset (d_st, 0, sizeof h));
    if (n)
    {
        const size_t blocks = (n + RAYS_BLOCK - 1) / RAYS_BLOCK;
        k_rays<true><<<(unsigned) blocks, RAYS_BLOCK>>> (
            tree_ref (t), start, (const float4*) d_o, (const float4*) d_d, n, d_leaf, (float4*) d_pts, d_vox, d_st);
        g_launches++;
    }
    cudaError_t e = cudaMemcpy (h, d_st, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree (d_st);
    if (e != cudaSuccess) return fail ("voxgpu_rays_stats: %s", cudaGetErrorString (e));
    out->rays = h[0]; out->hits = h[1]; out->calls = h[2]; out->inner = h[3]; out->leaf = h[4];
    out->dense = h[5]; out->dots_tested = h[6]; out->expansions = h[7]; out->max_stack = h[8];
    return 0;
}

/* ------------------------------------------------- queries with HOST buffers
 *
 * What vox_ray_tree_intersection[_batch] / vox_tree_ball_collidep[_batch] call.  The caller's
 * arrays are ordinary pageable memory, so a batch moves through a per-device pipeline of
 * HP_SLOTS chunks: the host threads copy chunk c into page-locked staging (OpenMP, one memcpy per
 * thread), the stream of its slot runs H2D copy -> kernel -> D2H copy, and while the device works
 * on chunk c the host drains chunk c-1 from staging into the caller's arrays (hit points only
 * where there is a hit: "res is written only on a hit", search.h:24-26).  Staging and device
 * scratch live as long as the process: a single query (n = 1) costs two small copies and a
 * launch, no allocation.
 */
#define HP_SLOTS 3
#define HP_CHUNK (1u << 20)        /* queries per chunk */

struct HostSlot
{
    cudaStream_t st;
    cudaEvent_t done;
    char *h_in, *h_out;            /* page-locked: 32 B in, 24 B out per query */
    char *d_in, *d_out;
};

struct HostPipe
{
    int ready;
    std::mutex lock;
    HostSlot slot[HP_SLOTS];
};
static HostPipe g_pipe[64];

static std::mutex g_pipe_init;

static int pipe_get (int device, HostPipe **out)
{
    HostPipe *p = &g_pipe[device];
    std::lock_guard<std::mutex> guard (g_pipe_init);
    if (!p->ready)
    {
        for (int i=0; i<HP_SLOTS; i++)
        {
            HostSlot &s = p->slot[i];
            CU (cudaStreamCreateWithFlags (&s.st, cudaStreamNonBlocking));
            CU (cudaEventCreateWithFlags (&s.done, cudaEventDisableTiming));
            CU (cudaHostAlloc (&s.h_in, (size_t) HP_CHUNK * 32, cudaHostAllocDefault));
            CU (cudaHostAlloc (&s.h_out, (size_t) HP_CHUNK * 24, cudaHostAllocDefault));
            CU (cudaMalloc (&s.d_in, (size_t) HP_CHUNK * 32));
            CU (cudaMalloc (&s.d_out, (size_t) HP_CHUNK * 24));
        }
        p->ready = 1;
    }
    *out = p;
    return 0;
}

/* memcpy spread over the host threads (a single thread moves ~10 GB/s, PCIe 5 x16 carries 55) */
static void par_copy (void *dst, const void *src, size_t bytes)
{
    const size_t piece = 1u << 20;
    const long pieces = (long) ((bytes + piece - 1) / piece);
    if (pieces <= 2) { memcpy (dst, src, bytes); return; }
    #pragma omp parallel for schedule(static)
    for (long i=0; i<pieces; i++)
    {
        const size_t off = (size_t) i * piece;
        memcpy ((char*) dst + off, (const char*) src + off, bytes - off < piece ? bytes - off : piece);
    }
}

extern "C" int voxgpu_rays_host (const voxgpu_tree *t, uint32_t start, const float *o, const float *d, size_t n,
                                 int32_t *leaf, float *pts, int32_t *vox, long *n_hits)
{
    if (check_tree (t, start, "voxgpu_rays_host")) return -1;
    if (n_hits) *n_hits = 0;
    if (n == 0) return 0;
    HostPipe *p;
    if (pipe_get (t->device, &p)) return -1;
    std::lock_guard<std::mutex> guard (p->lock);

    const size_t chunks = (n + HP_CHUNK - 1) / HP_CHUNK;
    long hits = 0;
    for (size_t c = 0; c <= chunks; c++)
    {
        if (c < chunks)
        {
            /* stage A: chunk c goes to the device (its slot was drained two iterations ago) */
            HostSlot &s = p->slot[c % HP_SLOTS];
            const size_t lo = c * HP_CHUNK, m = (n - lo < HP_CHUNK) ? n - lo : HP_CHUNK;
            par_copy (s.h_in, o + 4*lo, m * 16);
            par_copy (s.h_in + (size_t) HP_CHUNK * 16, d + 4*lo, m * 16);
            float4 *d_o = (float4*) s.d_in, *d_d = (float4*) (s.d_in + (size_t) HP_CHUNK * 16);
            float4 *d_p = (float4*) s.d_out;
            int *d_l = (int*) (s.d_out + (size_t) HP_CHUNK * 16), *d_v = (int*) (s.d_out + (size_t) HP_CHUNK * 20);
            CU (cudaMemcpyAsync (d_o, s.h_in, m * 16, cudaMemcpyHostToDevice, s.st));
            CU (cudaMemcpyAsync (d_d, s.h_in + (size_t) HP_CHUNK * 16, m * 16, cudaMemcpyHostToDevice, s.st));
            if (voxgpu_rays_device (t, start, (float*) d_o, (float*) d_d, m, d_l, pts ? (float*) d_p : NULL, vox ? d_v : NULL, s.st)) return -1;
            if (pts) CU (cudaMemcpyAsync (s.h_out, d_p, m * 16, cudaMemcpyDeviceToHost, s.st));
            CU (cudaMemcpyAsync (s.h_out + (size_t) HP_CHUNK * 16, d_l, m * 4, cudaMemcpyDeviceToHost, s.st));
            if (vox) CU (cudaMemcpyAsync (s.h_out + (size_t) HP_CHUNK * 20, d_v, m * 4, cudaMemcpyDeviceToHost, s.st));
            CU (cudaEventRecord (s.done, s.st));
        }
        if (c > 0)
        {
            /* stage B: chunk c-1 comes back while the device works on chunk c */
            HostSlot &s = p->slot[(c - 1) % HP_SLOTS];
            const size_t lo = (c - 1) * HP_CHUNK, m = (n - lo < HP_CHUNK) ? n - lo : HP_CHUNK;
            CU (cudaEventSynchronize (s.done));
            const int32_t *h_l = (const int32_t*) (s.h_out + (size_t) HP_CHUNK * 16);
            const float4 *h_p = (const float4*) s.h_out;
            if (leaf) par_copy (leaf + lo, h_l, m * 4);
            if (vox) par_copy (vox + lo, s.h_out + (size_t) HP_CHUNK * 20, m * 4);
            long h = 0;
            if (pts || n_hits)
            {
                float4 *out_p = (float4*) pts + lo;
                #pragma omp parallel for schedule(static) reduction(+:h) if (m > 65536)
                for (long i=0; i<(long) m; i++)
                    if (h_l[i] >= 0) { h++; if (pts) out_p[i] = h_p[i]; }
            }
            hits += h;
        }
    }
    if (n_hits) *n_hits = hits;
    return 0;
}

/* --------------------------------------------------------------- K3: balls */

#define BALLS_BLOCK 128

__global__ void __launch_bounds__(BALLS_BLOCK)
k_balls (TreeRef tr, const float4 * __restrict__ centers, const float * __restrict__ radii, size_t n,
         unsigned char * __restrict__ out)
{
    const size_t i = (size_t) blockIdx.x * BALLS_BLOCK + threadIdx.x;
    if (i >= n) return;
    unsigned stack[VOXGPU_BALL_STACK];
    const float4 c = __ldg (centers + i);
    out[i] = ball_tree (tr, 0, v3 (c), __ldg (radii + i), stack) ? 1 : 0;
}

extern "C" int voxgpu_balls_device (const voxgpu_tree *t, const float *d_c, const float *d_r, size_t n,
                                    uint8_t *d_out, voxgpu_stream s)
{
    if (check_tree (t, 0, "voxgpu_balls_device")) return -1;
    if (n == 0) return 0;
    const size_t blocks = (n + BALLS_BLOCK - 1) / BALLS_BLOCK;
    if (blocks > 0x7fffffffull) return fail ("voxgpu_balls_device: too many balls");
    k_balls<<<(unsigned) blocks, BALLS_BLOCK, 0, (cudaStream_t) s>>> (tree_ref (t), (const float4*) d_c, d_r, n, d_out);
    g_launches++;
    CU (cudaGetLastError ());
    return 0;
}

extern "C" int voxgpu_balls_host (const voxgpu_tree *t, const float *c, const float *r, size_t n,
                                  uint8_t *out, long *n_hits)
{
    if (check_tree (t, 0, "voxgpu_balls_host")) return -1;
    if (n_hits) *n_hits = 0;
    if (n == 0) return 0;
    HostPipe *p;
    if (pipe_get (t->device, &p)) return -1;
    std::lock_guard<std::mutex> guard (p->lock);

    /* the same pipeline as voxgpu_rays_host: 20 B in, 1 B out per ball */
    const size_t chunks = (n + HP_CHUNK - 1) / HP_CHUNK;
    long hits = 0;
    for (size_t k = 0; k <= chunks; k++)
    {
        if (k < chunks)
        {
            HostSlot &s = p->slot[k % HP_SLOTS];
            const size_t lo = k * HP_CHUNK, m = (n - lo < HP_CHUNK) ? n - lo : HP_CHUNK;
            par_copy (s.h_in, c + 4*lo, m * 16);
            par_copy (s.h_in + (size_t) HP_CHUNK * 16, r + lo, m * 4);
            float4 *d_c = (float4*) s.d_in;
            float *d_r = (float*) (s.d_in + (size_t) HP_CHUNK * 16);
            CU (cudaMemcpyAsync (d_c, s.h_in, m * 16, cudaMemcpyHostToDevice, s.st));
            CU (cudaMemcpyAsync (d_r, s.h_in + (size_t) HP_CHUNK * 16, m * 4, cudaMemcpyHostToDevice, s.st));
            if (voxgpu_balls_device (t, (float*) d_c, d_r, m, (uint8_t*) s.d_out, s.st)) return -1;
            CU (cudaMemcpyAsync (s.h_out, s.d_out, m, cudaMemcpyDeviceToHost, s.st));
            CU (cudaEventRecord (s.done, s.st));
        }
        if (k > 0)
        {
            HostSlot &s = p->slot[(k - 1) % HP_SLOTS];
            const size_t lo = (k - 1) * HP_CHUNK, m = (n - lo < HP_CHUNK) ? n - lo : HP_CHUNK;
            CU (cudaEventSynchronize (s.done));
            memcpy (out + lo, s.h_out, m);
            if (n_hits) for (size_t i=0; i<m; i++) hits += s.h_out[i] != 0;
        }
    }
    if (n_hits) *n_hits = hits;
    return 0;
}

/* -------------------------------------------------------------- K1: render */

#define RENDER_BLOCK 128
#define RENDER_STRIPS 64u

struct DevCamera
{
    int kind;
    float px, py, pz;
    float mul, xsub, ysub;
    float qw, qi, qj, qk;
    float sinphi, cosphi, k;
    const float4 *table;
};

/* Flattened M-tree of lights (voxgpu_lights_flat): 2 x float4 per node {bounding centre, squared
   radius}{skip, leaf, first sphere, count} and per sphere {centre, squared radius}{colour, radius}. */
struct DevLights
{
    const float4 *nodes;
    const float4 *spheres;
    unsigned n_nodes;
    float ar, ag, ab;
};

struct DevFrame
{
    int w, y0, rows, pitch;
    unsigned amask;
    unsigned char rloss, gloss, bloss, rshift, gshift, bshift;
    DevLights lt;
    int il_world, il_rank;          /* this launch draws the 4-row tile rows r with r % il_world == il_rank */
    /* pixels outside [cx0, cx1) x [cy0, cy1) (frame coordinates) cannot see the tree's root box:
       tiles entirely outside are only cleared (screen_bounds) */
    int cx0, cx1, cy0, cy1;
    /* screen mask (k_mark): one byte per 8x4 tile of the whole frame, 0 = no voxel can project into
       the tile; mask[mask_tiles] is the "no bounds this frame" flag.  NULL: not in use */
    const unsigned char *mask;
    int mask_tiles;
    unsigned mask_epoch;            /* a tile is marked iff its byte equals this (buffers are reused without clearing) */
};

/* vect-ops.c:8-16 vox_cross_product */
__device__ __forceinline__ V3 cross3 (const V3 a, const V3 b)
{
    V3 r;
    r.x = __fsub_rn (__fmul_rn (a.y, b.z), __fmul_rn (a.z, b.y));
    r.y = __fadd_rn (__fmul_rn (-a.x, b.z), __fmul_rn (a.z, b.x));
    r.z = __fsub_rn (__fmul_rn (a.x, b.y), __fmul_rn (a.y, b.x));
    return r;
}

/* simple-camera.c:21-31 + vect-ops.c:28-36, doom-camera.c:24-57 */
__device__ __forceinline__ V3 screen2world (const DevCamera &cam, int sx, int sy, int w)
{
    V3 dir;
    if (cam.kind == VOXGPU_CAMERA_TABLE)
    {
        const float4 t = __ldg (cam.table + (size_t) sy * w + sx);
        dir.x = t.x; dir.y = t.y; dir.z = t.z;
        return dir;
    }
    V3 ray;
    ray.x = __fsub_rn (__fmul_rn (cam.mul, (float) sx), cam.xsub);
    ray.y = 1.0f;
    ray.z = __fsub_rn (cam.ysub, __fmul_rn (cam.mul, (float) sy));
    if (cam.kind == VOXGPU_CAMERA_DOOM)
    {
        ray.z = __fadd_rn (ray.z, cam.k);
        dir.x = __fsub_rn (__fmul_rn (ray.x, cam.cosphi), __fmul_rn (ray.y, cam.sinphi));
        dir.y = __fadd_rn (__fmul_rn (ray.y, cam.cosphi), __fmul_rn (ray.x, cam.sinphi));
        dir.z = ray.z;
        return dir;
    }
    V3 q; q.x = cam.qi; q.y = cam.qj; q.z = cam.qk;
    V3 t = cross3 (q, ray);
    t.x = __fmul_rn (t.x, 2.0f); t.y = __fmul_rn (t.y, 2.0f); t.z = __fmul_rn (t.z, 2.0f);
    const V3 t2 = cross3 (q, t);
    dir.x = __fadd_rn (__fadd_rn (ray.x, __fmul_rn (cam.qw, t.x)), t2.x);
    dir.y = __fadd_rn (__fadd_rn (ray.y, __fmul_rn (cam.qw, t.y)), t2.y);
    dir.z = __fadd_rn (__fadd_rn (ray.z, __fmul_rn (cam.qw, t.z)), t2.z);
    return dir;
}

/* geom.c:133-139 vox_sqr_metric: ((dx^2 + dy^2) + dz^2) */
__device__ __forceinline__ float sqr_metric4 (const float4 c, const V3 p)
{
    const float dx = __fsub_rn (c.x, p.x), dy = __fsub_rn (c.y, p.y), dz = __fsub_rn (c.z, p.z);
    return __fadd_rn (__fadd_rn (__fmul_rn (dx, dx), __fmul_rn (dy, dy)), __fmul_rn (dz, dz));
}

/* lights.c:91-102 get_color_callback: a sphere's contribution fades linearly from its centre */
__device__ __forceinline__ void add_light (const DevLights &lt, unsigned sphere, const V3 p, float &lr, float &lg, float &lb)
{
    const float4 g = __ldg (lt.spheres + 2*sphere), c = __ldg (lt.spheres + 2*sphere + 1);
    const float dist = fminf (c.w, __fsqrt_rn (sqr_metric4 (g, p)));
    const float add = __fsub_rn (1.0f, __fdiv_rn (dist, c.w));
    lr = __fadd_rn (lr, __fmul_rn (add, c.x));
    lg = __fadd_rn (lg, __fmul_rn (add, c.y));
    lb = __fadd_rn (lb, __fmul_rn (add, c.z));
}

/* lights.c:104-116 vox_get_light over mtree.c:463-493, on the pre-order node array: same nodes
   entered, same spheres in the same order as the reference's recursion, so the same fp32 sums. */
__device__ __noinline__ void get_light (const DevLights &lt, const V3 p, float &lr, float &lg, float &lb)
{
    lr = lt.ar; lg = lt.ag; lb = lt.ab;
    unsigned i = 0;
    while (i < lt.n_nodes)
    {
        const float4 b = __ldg (lt.nodes + 2*i);
        const float4 l = __ldg (lt.nodes + 2*i + 1);
        if (!(sqr_metric4 (b, p) < b.w)) { i = __float_as_uint (l.x); continue; }
        if (__float_as_uint (l.y))
        {
            const unsigned first = __float_as_uint (l.z), num = __float_as_uint (l.w);
            if (num == 1) add_light (lt, first, p, lr, lg, lb);      /* mtree.c:475-482: no second test */
            else for (unsigned k=0; k<num; k++)
            {
                const float4 g = __ldg (lt.spheres + 2*(first + k));
                if (sqr_metric4 (g, p) < g.w) add_light (lt, first + k, p, lr, lg, lb);
            }
        }
        i++;
    }
}

/* renderer.c:27-52 get_color + SDL_MapRGB; LIT = a light manager is attached (renderer.c:41-49) */
template <bool LIT>
__device__ __forceinline__ unsigned shade (const TreeRef &tr, const DevFrame &fr, const unsigned char *tex, const V3 p)
{
    const int x = abs (__float2int_rz (__fdiv_rn (p.x, tr.vx))) & 31;
    const int y = abs (__float2int_rz (__fdiv_rn (p.y, tr.vy))) & 31;
    const int z = abs (__float2int_rz (__fdiv_rn (p.z, tr.vz))) & 31;
    const unsigned c = __ldg (tex + (x*1024 + y*32 + z));
    if (!LIT)
        return ((c >> fr.rloss) << fr.rshift) | ((c >> fr.gloss) << fr.gshift) | ((c >> fr.bloss) << fr.bshift) | fr.amask;
    float lr, lg, lb;
    get_light (fr.lt, p, lr, lg, lb);
    const float fc = (float) c;
    const unsigned r = __float2uint_rz (__fmul_rn (fminf (lr, 1.0f), fc)) & 0xffu;
    const unsigned g = __float2uint_rz (__fmul_rn (fminf (lg, 1.0f), fc)) & 0xffu;
    const unsigned b = __float2uint_rz (__fmul_rn (fminf (lb, 1.0f), fc)) & 0xffu;
    return ((r >> fr.rloss) << fr.rshift) | ((g >> fr.gloss) << fr.gshift) | ((b >> fr.bloss) << fr.bshift) | fr.amask;
}

/*
 * One warp renders an 8x4 pixel tile (coherent rays, 32-byte row segments on
 * the store); a block of 4 warps covers a 16x8 tile pair.  Every pixel of the
 * band is written, 0 on a miss: that is vox_redraw's memset of the squares
 * (renderer.c:324) folded into the store.
 */
template <bool OUT_HITS>
__global__ void __launch_bounds__(RENDER_BLOCK)
k_render (TreeRef tr, DevCamera cam, DevFrame fr, const unsigned char * __restrict__ tex,
          unsigned * __restrict__ pixels, int * __restrict__ leaf_ids, float4 * __restrict__ points)
{
    const int tiles_x = fr.w >> 3;
    const int ntiles = tiles_x * (fr.rows >> 2);
    const int tile = blockIdx.x * (RENDER_BLOCK/32) + (threadIdx.x >> 5);
    if (tile >= ntiles) return;
    const int lane = threadIdx.x & 31;
    const int sx = (tile % tiles_x) * 8 + (lane & 7);
    const int ry = (tile / tiles_x) * 4 + (lane >> 3);   /* row inside the band */
    const int sy = fr.y0 + ry;

    float4 stack[VOXGPU_RAY_STACK];
    V3 origin; origin.x = cam.px; origin.y = cam.py; origin.z = cam.pz;
    const V3 dir = screen2world (cam, sx, sy, fr.w);
    RayResult r;
    ray_tree<false> (tr, 0, origin, dir, stack, r, NULL);

    unsigned color = 0;
    if (r.leaf >= 0) color = shade<false> (tr, fr, tex, r.point);
    pixels[(size_t) ry * fr.pitch + sx] = color;
    if (OUT_HITS)
    {
        const size_t pos = (size_t) ry * fr.w + sx;
        if (leaf_ids) leaf_ids[pos] = r.leaf;
        if (points && r.leaf >= 0) points[pos] = make_float4 (r.point.x, r.point.y, r.point.z, 0.0f);
    }
}

/*
 * Persistent form of K1 (the one voxgpu_render launches).  70-80% of the 8x4 tiles
 * of a C2 frame show no voxel (half of them do not even see the root box) and finish early; with
 * one CTA per 4 tiles the SMs spent a third of the kernel waiting for CTA
 * launches (profiles/r01_*).  Here the grid is sized to the resident capacity
 * of the GPU and every warp pulls tile indices from a global counter until the
 * frame is done.  The counter resets itself: the last warp to leave zeroes it,
 * so no memset precedes the launch.
 */
__device__ __forceinline__ unsigned long long globaltimer_ns ()
{
    unsigned long long t;
    asm volatile ("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned smid ()
{
    unsigned r;
    asm ("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}

/*
 * Expensive-tiles-first scheduling across consecutive frames.  A few percent of
 * the tiles of a frame (rays grazing along a surface) take 30-100 times longer
 * than the median tile; when such a tile is handed out late the whole GPU waits
 * for it (the per-tile trace showed the last 40% of the kernel to be a tail).
 * Frames of an animation are coherent, so every launch records the cost of each
 * tile (iterations of its warp), and before the next launch on the same schedule
 * a small kernel sorts the tiles into three classes by the largest cost seen in
 * their 5x5-tile neighbourhood (exactly which tile is expensive changes with the
 * slightest camera move; that its neighbourhood is does not).  The render kernel
 * then hands out class 0, 1, 2 in that order.  A pure performance hint: every
 * tile is still rendered exactly once and results do not depend on it.
 */
struct DevSched
{
    const unsigned *list[3];      /* tiles by class, filled by k_sched_build */
    unsigned *count;              /* [3]; zeroed again by the render kernel's last warp */
    unsigned char *cost_next;     /* per tile, written by this launch */
};

#define SCHED_RADIUS 2

__global__ void k_sched_build (const unsigned char * __restrict__ cost_prev, int tiles_x, int tiles_y,
                               unsigned t_heavy, unsigned t_mid,
                               unsigned *list0, unsigned *list1, unsigned *list2, unsigned *count)
{
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    const int ntiles = tiles_x * tiles_y;
    unsigned cls = 3;
    if (tile < ntiles)
    {
        const int tx = tile % tiles_x, ty = tile / tiles_x;
        unsigned m = 0;
        for (int dy=-SCHED_RADIUS; dy<=SCHED_RADIUS; dy++)
            for (int dx=-SCHED_RADIUS; dx<=SCHED_RADIUS; dx++)
            {
                const int x = tx + dx, y = ty + dy;
                if (x >= 0 && x < tiles_x && y >= 0 && y < tiles_y) m = max (m, (unsigned) __ldg (cost_prev + y * tiles_x + x));
            }
        cls = m >= t_heavy ? 0u : (m >= t_mid ? 1u : 2u);
    }
    /* warp-aggregated append to the three lists */
    const int lane = threadIdx.x & 31;
    unsigned *lists[3] = {list0, list1, list2};
    #pragma unroll
    for (unsigned c=0; c<3; c++)
    {
        const unsigned m = __ballot_sync (0xffffffffu, cls == c);
        if (!m) continue;
        unsigned base = 0;
        if (lane == __ffs (m) - 1) base = atomicAdd (count + c, (unsigned) __popc (m));
        base = __shfl_sync (0xffffffffu, base, __ffs (m) - 1);
        if (cls == c) lists[c][base + __popc (m & ((1u << lane) - 1u))] = (unsigned) tile;
    }
}

/* Resident CTAs per SM the compiler must allow for K1.  Measured on C2 (tools/ab_variants.sh):
   7 (72 registers, what ptxas picks on its own) 12.31 Grays/s, 8 (61 registers, no spills) 12.48,
   9 (56 registers, 12 bytes spilled) 12.01. */
#ifndef VOXGPU_RENDER_MIN_BLOCKS
#define VOXGPU_RENDER_MIN_BLOCKS 8
#endif
template <bool OUT_HITS, unsigned STRIPS, bool LIT>
__global__ void __launch_bounds__(RENDER_BLOCK, VOXGPU_RENDER_MIN_BLOCKS)
k_render_persistent (TreeRef tr, DevCamera cam, DevFrame fr, const unsigned char * __restrict__ tex,
                     unsigned * __restrict__ pixels, int * __restrict__ leaf_ids, float4 * __restrict__ points,
                     unsigned *counters, unsigned long long *trace, DevSched sch)
{
    const int tiles_x = fr.w >> 3;
    /* tile rows of the band this launch owns (all of them unless the frame is interleaved over GPUs) */
    const int band_rows = fr.rows >> 2;
    const int own_rows = fr.il_rank < band_rows ? (band_rows - fr.il_rank + fr.il_world - 1) / fr.il_world : 0;
    const unsigned ntiles = (unsigned) (tiles_x * own_rows);
    const int lane = threadIdx.x & 31;
    const unsigned total_warps = gridDim.x * (RENDER_BLOCK/32);
    float4 stack[VOXGPU_RAY_STACK];
    V3 origin; origin.x = cam.px; origin.y = cam.py; origin.z = cam.pz;
#if VOXGPU_SMEM_TOP
    /* the top levels of the tree (breadth-first order: simply the first nodes), once per persistent CTA */
    __shared__ float4 s_top[3 * VOXGPU_SMEM_TOP];
    for (unsigned i = threadIdx.x; i < tr.ntop_avail; i += RENDER_BLOCK)
    {
        float4 a, b;
        load_rec (tr.rec + i, a, b);
        s_top[2*i] = a; s_top[2*i + 1] = b;
        s_top[2*VOXGPU_SMEM_TOP + i] = __ldg (tr.ctr + i);
    }
    __syncthreads ();
    tr.top_rec = s_top; tr.top_ctr = s_top + 2*VOXGPU_SMEM_TOP; tr.ntop = tr.ntop_avail;
#endif

    /* Tiles are handed out strip-interleaved: consecutive tickets land in
       RENDER_STRIPS different horizontal strips of the frame.  In raster order
       the empty top and bottom of a frame (half of a skull frame misses the root
       box, 70-80% of it shows no voxel) are done in two bursts that only store pixels while the middle only
       computes; interleaved, stores and traversal overlap for the whole kernel,
       which matters when the pixels go to mapped host memory over PCIe. */
    const unsigned per_strip = (ntiles + STRIPS - 1) / STRIPS;
    unsigned n0 = 0, n1 = 0;
    if (sch.count) { n0 = sch.count[0]; n1 = n0 + sch.count[1]; }
    const unsigned tickets = sch.count ? ntiles : per_strip * STRIPS;

    for (;;)
    {
        unsigned ticket = 0;
        if (lane == 0) ticket = atomicAdd (&counters[0], 1u);
        ticket = __shfl_sync (0xffffffffu, ticket, 0);
        if (ticket >= tickets) break;
        unsigned tile;
        if (sch.count)
            tile = ticket < n0 ? sch.list[0][ticket] : (ticket < n1 ? sch.list[1][ticket - n0] : sch.list[2][ticket - n1]);
        else
        {
            tile = (ticket % STRIPS) * per_strip + ticket / STRIPS;
            if (tile >= ntiles) continue;
        }

        unsigned long long t0 = 0;
        if (trace) t0 = globaltimer_ns ();
        const int sx = (int) (tile % tiles_x) * 8 + (lane & 7);
        const int ry = ((int) (tile / tiles_x) * fr.il_world + fr.il_rank) * 4 + (lane >> 3);
        const int sy = fr.y0 + ry;
        {
            /* the whole 8x4 tile lies outside the root box's screen bounds: every ray misses
               (search.c:22-27 at the root), the pixels are the zeros of vox_redraw's memset */
            const int tx0 = sx - (lane & 7), ty0 = sy - (lane >> 3);
            bool dark = tx0 >= fr.cx1 || tx0 + 8 <= fr.cx0 || ty0 >= fr.cy1 || ty0 + 4 <= fr.cy0;
            if (!dark && fr.mask)
                dark = __ldg (fr.mask + (size_t) (ty0 >> 2) * tiles_x + (tx0 >> 3)) != fr.mask_epoch &&
                       __ldg (fr.mask + fr.mask_tiles) != fr.mask_epoch;
            if (dark)
            {
                pixels[(size_t) ry * fr.pitch + sx] = 0u;
                if (OUT_HITS && leaf_ids) leaf_ids[(size_t) ry * fr.w + sx] = -1;
                if (sch.cost_next && lane == 0) sch.cost_next[tile] = 0;
                continue;
            }
        }
        const V3 dir = screen2world (cam, sx, sy, fr.w);
        RayResult r;
        const int iterations = ray_tree_warp (tr, 0, origin, dir, true, stack, r);
        if (sch.cost_next && lane == 0) sch.cost_next[tile] = (unsigned char) min (iterations, 255);

        unsigned color = 0;
        if (r.leaf >= 0) color = shade<LIT> (tr, fr, tex, r.point);
        pixels[(size_t) ry * fr.pitch + sx] = color;
        if (OUT_HITS)
        {
            const size_t pos = (size_t) ry * fr.w + sx;
            if (leaf_ids) leaf_ids[pos] = r.leaf;
            if (points && r.leaf >= 0) points[pos] = make_float4 (r.point.x, r.point.y, r.point.z, 0.0f);
        }
        __syncwarp ();
        if (trace && lane == 0)
        {
            /* development trace (voxgpu_set_trace): when and where each tile ran */
            trace[3*(size_t) tile] = t0;
            trace[3*(size_t) tile + 1] = globaltimer_ns ();
            trace[3*(size_t) tile + 2] = ((unsigned long long) smid () << 32) | (unsigned) iterations;
        }
    }
    if (lane == 0)
    {
        const unsigned done = atomicAdd (&counters[1], 1u);
        if (done + 1 == total_warps)
        {
            counters[0] = 0; counters[1] = 0;
            if (sch.count) { sch.count[0] = 0; sch.count[1] = 0; sch.count[2] = 0; }   /* for the next k_sched_build */
            __threadfence ();
        }
    }
}

/*
 * K1 with per-LANE work refill (the form voxgpu_render launches for BEST).  In the
 * tile form above a warp keeps its 32 rays until the slowest has finished: on C2 the
 * average instruction ran with 21 of 32 lanes, most of the missing ones being lanes whose
 * ray was already done.  Here the ticket counter hands out PIXELS (in tile order, so
 * that rays fetched together are neighbours): whenever at least `refill_min` lanes of
 * a warp are idle, they retire their finished pixels, draw new ones with one warp-
 * aggregated atomicAdd (ballot + popc rank), generate their rays and join the search
 * at the root while the other lanes carry on where they were.  The traversal itself
 * is warp_search_step, unchanged; results do not depend on which lane or in which
 * company a ray is searched.
 */
#define STREAM_BATCH 64u
template <bool OUT_HITS, unsigned STRIPS>
__global__ void __launch_bounds__(RENDER_BLOCK, 7)
k_render_stream (TreeRef tr, DevCamera cam, DevFrame fr, const unsigned char * __restrict__ tex,
                 unsigned * __restrict__ pixels, int * __restrict__ leaf_ids, float4 * __restrict__ points,
                 unsigned *counters, int refill_min)
{
    const int tiles_x = fr.w >> 3;
    const unsigned ntiles = (unsigned) (tiles_x * (fr.rows >> 2));
    const int lane = threadIdx.x & 31;
    const unsigned total_warps = gridDim.x * (RENDER_BLOCK/32);
    const unsigned per_strip = (ntiles + STRIPS - 1) / STRIPS;
    const unsigned tickets = per_strip * STRIPS * 32u;     /* pixel tickets; some map to no tile */
    float4 stack[VOXGPU_RAY_STACK];
    V3 origin; origin.x = cam.px; origin.y = cam.py; origin.z = cam.pz;

    LaneSearch s;
    lane_search_begin (s, 0, origin, false);
    RayK k = ray_constants (origin);
    int pix_x = -1, pix_y = 0;          /* the pixel this lane is working on, -1 = none */
    /* warp-uniform: the warp's batch of pixel tickets [next, next + have), and whether the
       global counter has run out.  One atomic per STREAM_BATCH pixels, so that lanes in
       the middle of a long search do not wait for one every time their neighbours refill. */
    unsigned next = 0, have = 0;
    bool dry = false;

    for (;;)
    {
        unsigned idle = __ballot_sync (0xffffffffu, s.cur == VOXGPU_DEAD);
        if ((have || !dry) && __popc (idle) >= refill_min)
        {
            /* retire */
            if (s.cur == VOXGPU_DEAD && pix_x >= 0)
            {
                RayResult r;
                lane_search_result (s, stack, r);
                unsigned color = 0;
                if (r.leaf >= 0) color = shade<false> (tr, fr, tex, r.point);
                pixels[(size_t) pix_y * fr.pitch + pix_x] = color;
                if (OUT_HITS)
                {
                    const size_t pos = (size_t) pix_y * fr.w + pix_x;
                    if (leaf_ids) leaf_ids[pos] = r.leaf;
                    if (points && r.leaf >= 0) points[pos] = make_float4 (r.point.x, r.point.y, r.point.z, 0.0f);
                }
                pix_x = -1;
            }
            if (have == 0)
            {
                if (lane == 0) next = atomicAdd (&counters[0], STREAM_BATCH);
                next = __shfl_sync (0xffffffffu, next, 0);
                if (next >= tickets) dry = true; else have = STREAM_BATCH;     /* tickets is a multiple of the batch */
            }
            const unsigned n = (unsigned) __popc (idle);
            const unsigned take = n < have ? n : have;
            const unsigned rank = (unsigned) __popc (idle & ((1u << lane) - 1u));
            if (s.cur == VOXGPU_DEAD && rank < take)
            {
                const unsigned ticket = next + rank;
                const unsigned tt = ticket >> 5;
                const unsigned tile = (tt % STRIPS) * per_strip + tt / STRIPS;
                if (tile < ntiles)
                {
                    const unsigned in = ticket & 31u;
                    pix_x = (int) (tile % tiles_x) * 8 + (int) (in & 7u);
                    pix_y = (int) (tile / tiles_x) * 4 + (int) (in >> 3);
                    const V3 dir = screen2world (cam, pix_x, fr.y0 + pix_y, fr.w);
                    k = ray_constants (dir);
                    lane_search_begin (s, 0, origin, true);
                }
            }
            next += take; have -= take;
            idle = __ballot_sync (0xffffffffu, s.cur == VOXGPU_DEAD);
        }
        if (idle == 0xffffffffu)
        {
            if (!have && dry) break;
            continue;                   /* only tickets without a tile were drawn, or the batch ran out */
        }
        warp_search_step (tr, k, s, stack, lane);
    }
    if (pix_x >= 0)
    {
        RayResult r;
        lane_search_result (s, stack, r);
        unsigned color = 0;
        if (r.leaf >= 0) color = shade<false> (tr, fr, tex, r.point);
        pixels[(size_t) pix_y * fr.pitch + pix_x] = color;
        if (OUT_HITS)
        {
            const size_t pos = (size_t) pix_y * fr.w + pix_x;
            if (leaf_ids) leaf_ids[pos] = r.leaf;
            if (points && r.leaf >= 0) points[pos] = make_float4 (r.point.x, r.point.y, r.point.z, 0.0f);
        }
    }
    if (lane == 0)
    {
        const unsigned done = atomicAdd (&counters[1], 1u);
        if (done + 1 == total_warps) { counters[0] = 0; counters[1] = 0; __threadfence (); }
    }
}

/*
 * K1b: the reference's FAST / ADAPTIVE / ray-merge quality modes, pixel for pixel
 * (renderer.c:214-318).  In those modes the pixels of a 4x4 block depend on each other --
 * pixel i+1 is first searched inside the leaf pixel i hit, ADAPTIVE decides per block from
 * two corner rays, ray-merge copies the previous pixel's colour -- so a block is sequential.
 * One LANE owns one block and walks its 16 pixels in the reference's Z-order; the 32 lanes of
 * a warp (8x4 blocks = 32x16 pixels) advance together, one pixel index at a time, through
 * the warp-synchronous search with per-lane start nodes.  On the GPU this is the SLOW path:
 * BEST has 16 times the parallelism and is what to ask for when speed matters.
 */
__constant__ int c_rendering_order[16] = {0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15};   /* renderer.c:23-25 */

/* geom.c:133-139 / 141-147: ((0 + a*a) + b*b) + c*c */
__device__ __forceinline__ float sqr_metric3 (const V3 a, const V3 b)
{
    const float x = __fsub_rn (a.x, b.x), y = __fsub_rn (a.y, b.y), z = __fsub_rn (a.z, b.z);
    return __fadd_rn (__fadd_rn (__fmul_rn (x, x), __fmul_rn (y, y)), __fmul_rn (z, z));
}
__device__ __forceinline__ float sqr_norm3 (const V3 a)
{
    return __fadd_rn (__fadd_rn (__fmul_rn (a.x, a.x), __fmul_rn (a.y, a.y)), __fmul_rn (a.z, a.z));
}

/*
 * FAST / ADAPTIVE (+ ray merge), renderer.c:240-316: one lane owns one 4x4 block and walks its 16 pixels in
 * the reference's Z-order, because pixel i+1 is first searched inside the leaf pixel i ended in.  Only that
 * in-leaf search depends on the previous pixel; the search from the ROOT it falls back to (and the two corner
 * rays of ADAPTIVE) is a pure function of the pixel.  PRE = true: those root searches were done for every
 * pixel of the band by the pixel-parallel BEST kernel in a launch before this one (pre_leaf / pre_pts, dense
 * w-pitch), and the walk only looks them up; what is left to do sequentially is one node visit per pixel.
 * PRE = false is the single-launch form (every search inside the walk, 1/16 of the parallelism).
 */
template <bool OUT_HITS, bool LIT, bool PRE>
__global__ void __launch_bounds__(RENDER_BLOCK)
k_render_blocks (TreeRef tr, DevCamera cam, DevFrame fr, unsigned quality, const unsigned char * __restrict__ tex,
                 unsigned * __restrict__ pixels, int * __restrict__ leaf_ids, float4 * __restrict__ points,
                 unsigned *counters, const int * __restrict__ pre_leaf, const float4 * __restrict__ pre_pts)
{
    const int bw = fr.w >> 2, bh = fr.rows >> 2;               /* 4x4 blocks */
    const int tiles_x = (bw + 7) >> 3, tiles_y = (bh + 3) >> 2;
    const unsigned ntiles = (unsigned) (tiles_x * tiles_y);
    const int lane = threadIdx.x & 31;
    const unsigned total_warps = gridDim.x * (RENDER_BLOCK/32);
    const unsigned rnd_mode = quality & 3u, merge_mode = quality & 12u;
    const bool adaptive = rnd_mode == 2u;
    float4 stack[VOXGPU_RAY_STACK];
    unsigned sq[16];
    V3 origin; origin.x = cam.px; origin.y = cam.py; origin.z = cam.pz;

    for (;;)
    {
        unsigned tile = 0;
        if (lane == 0) tile = atomicAdd (&counters[0], 1u);
        tile = __shfl_sync (0xffffffffu, tile, 0);
        if (tile >= ntiles) break;

        const int bx = (int) (tile % tiles_x) * 8 + (lane & 7);
        const int by = (int) (tile / tiles_x) * 4 + (lane >> 3);
        const bool valid = bx < bw && by < bh;
        const int xstart = bx << 2, ystart = by << 2;          /* ystart: row inside the band */
        #pragma unroll
        for (int p=0; p<16; p++) sq[p] = 0;                    /* vox_redraw's memset, renderer.c:324 */
        if (OUT_HITS && valid && leaf_ids)
            for (int p=0; p<16; p++) leaf_ids[(size_t) (ystart + (p >> 2)) * fr.w + xstart + (p & 3)] = -1;

        /* renderer.c:216-239 */
        unsigned blk_mode = adaptive ? 1u : rnd_mode;          /* 0 BEST, 1 FAST */
        bool blk_merge = adaptive && merge_mode == 4u;
        int istart = adaptive ? 1 : 0, iend = 16;
        int leaf = -1;                                         /* node id of the previous hit, -1 = NULL */
        V3 inter = origin;                                     /* inter1 */
        V3 dir_c1 = origin, inter_c1 = origin;
        bool c1 = false;
        int prev_p = 0;

        /* step 0, 1: the two corner rays of ADAPTIVE (renderer.c:240-281);
           step 2+2i: pixel i searched inside the previous leaf (FAST), step 3+2i: from the root */
        for (int step = adaptive ? 0 : 2; step < 34; step++)
        {
            const int i = (step - 2) >> 1;
            const bool phase_leaf = step >= 2 && !(step & 1);
            int p = 0, px = xstart, py = ystart;
            bool merge = false, in_range = false;
            if (step == 1) { px += 3; py += 3; }
            if (step >= 2)
            {
                p = c_rendering_order[i];
                px += p & 3; py += p >> 2;
                merge = blk_merge && (i & 1);
                in_range = valid && i >= istart && i < iend;
            }
            bool want;
            unsigned start = 0;
            if (step == 0) want = valid;
            else if (step == 1) want = valid && c1;
            else if (phase_leaf) { want = in_range && !merge && blk_mode == 1u && leaf >= 0; start = leaf >= 0 ? (unsigned) leaf : 0u; }
            else want = in_range && !merge && (blk_mode == 0u || leaf < 0);

            RayResult r;
            r.leaf = -1;
            V3 dir = origin;
            if (PRE && !phase_leaf)
            {
                /* a search from the root: done already, for this very ray */
                if (want)
                {
                    if (step < 2) dir = screen2world (cam, px, fr.y0 + py, fr.w);     /* the corner criterion needs it */
                    const size_t pos = (size_t) py * fr.w + px;
                    r.leaf = pre_leaf[pos];
                    if (r.leaf >= 0) { const float4 q = pre_pts[pos]; r.point.x = q.x; r.point.y = q.y; r.point.z = q.z; }
                }
            }
            else if (__any_sync (0xffffffffu, want))
            {
                dir = screen2world (cam, valid ? px : 0, fr.y0 + (valid ? py : 0), fr.w);
                ray_tree_warp (tr, start, origin, dir, want, stack, r);
            }

            if (step == 0)
            {
                if (want)
                {
                    leaf = r.leaf; dir_c1 = dir;
                    if (r.leaf >= 0)
                    {
                        c1 = true; iend = 15; inter = r.point; inter_c1 = r.point;
                        sq[0] = shade<LIT> (tr, fr, tex, r.point);
                        if (OUT_HITS)
                        {
                            const size_t pos = (size_t) ystart * fr.w + xstart;
                            if (leaf_ids) leaf_ids[pos] = r.leaf;
                            if (points) points[pos] = make_float4 (r.point.x, r.point.y, r.point.z, 0.0f);
                        }
                    }
                }
            }
            else if (step == 1)
            {
                if (want && r.leaf >= 0)
                {
                    sq[15] = shade<LIT> (tr, fr, tex, r.point);
                    if (OUT_HITS)
                    {
                        const size_t pos = (size_t) (ystart + 3) * fr.w + xstart + 3;
                        if (leaf_ids) leaf_ids[pos] = r.leaf;
                        if (points) points[pos] = make_float4 (r.point.x, r.point.y, r.point.z, 0.0f);
                    }
                    /* renderer.c:262-278 */
                    const float d1 = sqr_metric3 (inter_c1, origin);
                    const float d2 = sqr_norm3 (dir_c1);
                    const float criteria = __fmul_rn (__fdiv_rn (d1, d2), sqr_metric3 (dir_c1, dir));
                    const float dist = sqr_metric3 (inter_c1, r.point);
                    if ((double) __fdiv_rn (dist, criteria) > 1.69) blk_mode = 0u;
                    if (d1 > 22500.0f) blk_merge = (blk_mode == 1u && merge_mode == 8u) ? true : blk_merge;
                    else blk_merge = false;
                }
            }
            else
            {
                if (want)
                {
                    leaf = r.leaf;
                    if (r.leaf >= 0) inter = r.point;
                }
                if (!phase_leaf)
                {
                    /* renderer.c:310-315, once both searches of pixel i are done */
                    if (in_range)
                    {
                        if (leaf >= 0)
                        {
                            sq[p] = merge ? sq[prev_p] : shade<LIT> (tr, fr, tex, inter);
                            if (OUT_HITS)
                            {
                                const size_t pos = (size_t) py * fr.w + px;
                                if (leaf_ids) leaf_ids[pos] = leaf;
                                if (points) points[pos] = make_float4 (inter.x, inter.y, inter.z, 0.0f);
                            }
                        }
                        prev_p = p;
                    }
                }
            }
        }

        if (valid)
        {
            #pragma unroll
            for (int row=0; row<4; row++)
                *(uint4*) (pixels + (size_t) (ystart + row) * fr.pitch + xstart) =
                    make_uint4 (sq[4*row], sq[4*row+1], sq[4*row+2], sq[4*row+3]);
        }
        __syncwarp ();
    }
    if (lane == 0)
    {
        const unsigned done = atomicAdd (&counters[1], 1u);
        if (done + 1 == total_warps) { counters[0] = 0; counters[1] = 0; __threadfence (); }
    }
}

/* Per-device pool of self-resetting tile counters; launches rotate through the
   slots so that renders in flight on different streams never share one. */
#define COUNTER_SLOTS 64
static unsigned *g_counters[64] = {0};
static int g_render_blocks[64][7] = {{0}};
static std::atomic<unsigned> g_counter_seq{0};
static int g_variant = 1;
static int g_refill_min = 8;
static unsigned long long *g_trace = NULL;

extern "C" void voxgpu_set_trace (void *d_trace) { g_trace = (unsigned long long*) d_trace; }

/* ------------------------------------------------------------------ lights */

struct voxgpu_lights
{
    int device;
    float4 *buf;
    DevLights ref;
};

extern "C" voxgpu_lights* voxgpu_lights_upload (const voxgpu_lights_flat *f)
{
    if (!f) { fail ("voxgpu_lights_upload: NULL argument"); return NULL; }
    voxgpu_lights *l = (voxgpu_lights*) calloc (1, sizeof *l);
    const size_t n4 = 2 * ((size_t) f->n_nodes + f->n_spheres);
    float4 *host = (float4*) malloc ((n4 ? n4 : 1) * sizeof (float4));
    for (uint32_t i=0; i<f->n_nodes; i++)
    {
        host[2*i] = make_float4 (f->node_bounds[4*i], f->node_bounds[4*i+1], f->node_bounds[4*i+2], f->node_bounds[4*i+3]);
        memcpy (&host[2*i+1], f->node_link + 4*i, sizeof (float4));
    }
    float4 *sph = host + 2 * (size_t) f->n_nodes;
    for (uint32_t i=0; i<f->n_spheres; i++)
    {
        memcpy (&sph[2*i], f->sphere_geom + 4*i, sizeof (float4));
        memcpy (&sph[2*i+1], f->sphere_light + 4*i, sizeof (float4));
    }
    cudaError_t e;
    if ((e = cudaGetDevice (&l->device)) != cudaSuccess ||
        (e = cudaMalloc (&l->buf, (n4 ? n4 : 1) * sizeof (float4))) != cudaSuccess ||
        (e = cudaMemcpy (l->buf, host, n4 * sizeof (float4), cudaMemcpyHostToDevice)) != cudaSuccess)
    {
        fail ("voxgpu_lights_upload: %s", cudaGetErrorString (e));
        free (host); cudaFree (l->buf); free (l);
        return NULL;
    }
    free (host);
    l->ref.nodes = l->buf;
    l->ref.spheres = l->buf + 2 * (size_t) f->n_nodes;
    l->ref.n_nodes = f->n_nodes;
    l->ref.ar = f->ambient[0]; l->ref.ag = f->ambient[1]; l->ref.ab = f->ambient[2];
    return l;
}

extern "C" void voxgpu_lights_free (voxgpu_lights *l)
{
    if (!l) return;
    cudaFree (l->buf);
    free (l);
}

extern "C" void voxgpu_set_variant (int v) { g_variant = v; }
extern "C" void voxgpu_set_refill_min (int n) { g_refill_min = n < 1 ? 1 : (n > 32 ? 32 : n); }

/* A render launch borrows slot `idx` of its device: a pair of tile counters and a screen-mask buffer.
   Launches rotate through COUNTER_SLOTS slots; should a slot come round again while its previous
   borrower is still queued or running on ANOTHER stream, the new stream waits for it (slot_acquire /
   slot_release: one event per slot), so sharing is always ordered. */
struct SlotGuard { cudaEvent_t ev; cudaStream_t last; int used; };
static SlotGuard g_guard[64][64];

static unsigned slot_acquire (int device, cudaStream_t st);
static void slot_release (int device, unsigned idx, cudaStream_t st);

static unsigned* counter_slot (int device, unsigned idx)
{
    if (!g_counters[device])
    {
        unsigned *p = NULL;
        if (cudaMalloc (&p, COUNTER_SLOTS * 2 * sizeof (unsigned)) != cudaSuccess) return NULL;
        if (cudaMemset (p, 0, COUNTER_SLOTS * 2 * sizeof (unsigned)) != cudaSuccess) return NULL;
        int sms = 0, b0 = 0, b1 = 0;
        cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, device);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor (&b0, k_render_persistent<false, RENDER_STRIPS, false>, RENDER_BLOCK, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor (&b1, k_render_persistent<true, RENDER_STRIPS, false>, RENDER_BLOCK, 0);
        g_render_blocks[device][0] = sms * (b0 > 0 ? b0 : 1);
        g_render_blocks[device][1] = sms * (b1 > 0 ? b1 : 1);
        int b2 = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor (&b2, k_render_blocks<true, false, false>, RENDER_BLOCK, 0);
        g_render_blocks[device][2] = sms * (b2 > 0 ? b2 : 1);
        int b5 = 0, b6 = 0;     /* with a light manager attached */
        cudaOccupancyMaxActiveBlocksPerMultiprocessor (&b5, k_render_persistent<true, RENDER_STRIPS, true>, RENDER_BLOCK, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor (&b6, k_render_blocks<true, true, false>, RENDER_BLOCK, 0);
        g_render_blocks[device][5] = sms * (b5 > 0 ? b5 : 1);
        g_render_blocks[device][6] = sms * (b6 > 0 ? b6 : 1);
        int b3 = 0, b4 = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor (&b3, k_render_stream<false, RENDER_STRIPS>, RENDER_BLOCK, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor (&b4, k_render_stream<true, RENDER_STRIPS>, RENDER_BLOCK, 0);
        g_render_blocks[device][3] = sms * (b3 > 0 ? b3 : 1);
        g_render_blocks[device][4] = sms * (b4 > 0 ? b4 : 1);
        /* stream-ordered scratch (cudaMallocAsync in voxgpu_render): keep freed blocks in the pool */
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool (&pool, device) == cudaSuccess)
        {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute (pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        g_counters[device] = p;
    }
    return g_counters[device] + 2 * idx;
}

static unsigned slot_acquire (int device, cudaStream_t st)
{
    const unsigned idx = g_counter_seq++ % COUNTER_SLOTS;
    SlotGuard &g = g_guard[device][idx];
    if (g.used && g.last != st) cudaStreamWaitEvent (st, g.ev, 0);
    return idx;
}

static void slot_release (int device, unsigned idx, cudaStream_t st)
{
    SlotGuard &g = g_guard[device][idx];
    if (!g.ev && cudaEventCreateWithFlags (&g.ev, cudaEventDisableTiming) != cudaSuccess) return;
    cudaEventRecord (g.ev, st);
    g.last = st; g.used = 1;
}

struct voxgpu_sched
{
    int device;
    int w, y0, rows;              /* geometry the recorded costs refer to */
    unsigned cap;                 /* tiles the buffers can hold */
    unsigned *list[3];
    unsigned *count;
    unsigned char *cost[2];
    int cur;                      /* cost buffer the next launch reads */
    unsigned t_heavy, t_mid;
};

extern "C" voxgpu_sched* voxgpu_sched_create (void)
{
    voxgpu_sched *sc = (voxgpu_sched*) calloc (1, sizeof *sc);
    const char *e1 = getenv ("VOXGPU_HEAVY_ITERS"), *e2 = getenv ("VOXGPU_MID_ITERS");
    sc->t_heavy = e1 ? (unsigned) atoi (e1) : 24u;
    sc->t_mid = e2 ? (unsigned) atoi (e2) : 8u;
    sc->device = -1;
    return sc;
}

static void sched_release (voxgpu_sched *sc)
{
    for (int i=0; i<3; i++) { cudaFree (sc->list[i]); sc->list[i] = NULL; }
    cudaFree (sc->cost[0]); cudaFree (sc->cost[1]); cudaFree (sc->count);
    sc->cost[0] = sc->cost[1] = NULL; sc->count = NULL; sc->cap = 0;
}

extern "C" void voxgpu_sched_destroy (voxgpu_sched *sc)
{
    if (!sc) return;
    sched_release (sc);
    free (sc);
}

/* Sort the tiles for this launch from the costs the previous one recorded
   (all zero after a change of geometry: plain order). */
static int sched_prepare (voxgpu_sched *sc, int device, int w, int y0, int rows, cudaStream_t st, DevSched *out)
{
    const int tiles_x = w >> 3, tiles_y = rows >> 2;
    const unsigned ntiles = (unsigned) (tiles_x * tiles_y);
    const bool restart = sc->device != device || sc->w != w || sc->y0 != y0 || sc->rows != rows || !sc->count;
    if (restart)
    {
        if (sc->cap < ntiles || sc->device != device)
        {
            sched_release (sc);
            for (int i=0; i<3; i++) CU (cudaMalloc (&sc->list[i], (size_t) ntiles * sizeof (unsigned)));
            for (int i=0; i<2; i++) CU (cudaMalloc (&sc->cost[i], ntiles));
            CU (cudaMalloc (&sc->count, 3 * sizeof (unsigned)));
            sc->cap = ntiles;
        }
        CU (cudaMemsetAsync (sc->count, 0, 3 * sizeof (unsigned), st));
        CU (cudaMemsetAsync (sc->cost[0], 0, ntiles, st));
        sc->device = device; sc->w = w; sc->y0 = y0; sc->rows = rows; sc->cur = 0;
    }
    const int p = sc->cur, q = p ^ 1;
    k_sched_build<<<(ntiles + 255) / 256, 256, 0, st>>> (sc->cost[p], tiles_x, tiles_y, sc->t_heavy, sc->t_mid,
                                                         sc->list[0], sc->list[1], sc->list[2], sc->count);
    g_launches++;
    CU (cudaGetLastError ());
    out->list[0] = sc->list[0]; out->list[1] = sc->list[1]; out->list[2] = sc->list[2];
    out->count = sc->count;
    out->cost_next = sc->cost[q];
    sc->cur = q;
    return 0;
}

/*
 * Screen mask.  Every voxel lies in the box of one node of the tree's cut (voxgpu_tree.cut); a ray that
 * hits the voxel passes through that box, so its pixel lies in the convex hull of the box's projected
 * corners (see voxgpu_screen_bounds for the argument) and therefore in their bounding rectangle.  One
 * thread per cut node projects its box (fp32 here: the error is ~1e-3 pixel, the rectangle is widened by 2)
 * and marks the 8x4 tiles the rectangle touches.  A tile nobody marked cannot show a voxel: the render
 * kernel clears it without casting its rays -- on a C2 frame half of the primary rays enter the root box
 * but only a fifth hit a voxel, and the ones in between cost 40-50 % of all node fetches.
 * A box not safely in front of the camera, or covering more than 4096 tiles, raises the "no bounds"
 * flag behind the mask instead (the frame is then rendered without the mask).
 */
__global__ void k_mark (TreeRef tr, const unsigned * __restrict__ cut, unsigned n_cut, DevCamera cam,
                        int w, int h, unsigned char * __restrict__ mask, int mask_tiles, unsigned char epoch,
                        int band_y0, int band_y1)
{
    /* eight lanes per node, one per corner of its box; min/max over the eight by shuffles */
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned i = gid >> 3;
    const int k = threadIdx.x & 7;
    const bool live = i < n_cut;
    float4 q0 = make_float4 (0, 0, 0, 0), q1 = q0;
    if (live) load_rec (tr.rec + cut[i], q0, q1);
    const float px = ((k & 1) ? q1.x : q0.x) - cam.px;
    const float py = ((k & 2) ? q1.y : q0.y) - cam.py;
    const float pz = ((k & 4) ? q1.z : q0.z) - cam.pz;
    float vx, vy, vz;
    if (cam.kind == VOXGPU_CAMERA_SIMPLE)
    {
        /* rotate by the conjugate quaternion */
        const float ux = -cam.qi, uy = -cam.qj, uz = -cam.qk;
        const float tx = 2.0f * (uy*pz - uz*py), ty = 2.0f * (uz*px - ux*pz), tz = 2.0f * (ux*py - uy*px);
        vx = px + cam.qw*tx + (uy*tz - uz*ty);
        vy = py + cam.qw*ty + (uz*tx - ux*tz);
        vz = pz + cam.qw*tz + (ux*ty - uy*tx);
    }
    else
    {
        vx = px*cam.cosphi + py*cam.sinphi;
        vy = py*cam.cosphi - px*cam.sinphi;
        vz = pz;
    }
    const float scale = fabsf (px) + fabsf (py) + fabsf (pz);
    bool bad = !(vy > 1e-3f * scale) || !(vy > 0.0f);
    const float rz = vz / vy - (cam.kind == VOXGPU_CAMERA_DOOM ? cam.k : 0.0f);
    const float sx = (vx / vy + cam.xsub) / cam.mul, sy = (cam.ysub - rz) / cam.mul;
    bad = bad || !(fabsf (sx) < 1.0e7f) || !(fabsf (sy) < 1.0e7f);
    float lox = bad ? 0.0f : sx, hix = lox, loy = bad ? 0.0f : sy, hiy = loy;
    unsigned anybad = bad ? 1u : 0u;
    #pragma unroll
    for (int off = 4; off; off >>= 1)
    {
        lox = fminf (lox, __shfl_xor_sync (0xffffffffu, lox, off)); hix = fmaxf (hix, __shfl_xor_sync (0xffffffffu, hix, off));
        loy = fminf (loy, __shfl_xor_sync (0xffffffffu, loy, off)); hiy = fmaxf (hiy, __shfl_xor_sync (0xffffffffu, hiy, off));
        anybad |= __shfl_xor_sync (0xffffffffu, anybad, off);
    }
    if (!live) return;
    if (!anybad)
    {
        const float x0 = floorf (lox) - 2.0f, x1 = ceilf (hix) + 3.0f, y0 = floorf (loy) - 2.0f, y1 = ceilf (hiy) + 3.0f;
        if (x1 <= 0.0f || y1 <= 0.0f || x0 >= (float) w || y0 >= (float) h) return;       /* off screen */
        const int tiles_x = w >> 3;
        const int tx0 = max ((int) x0, 0) >> 3, tx1 = (min ((int) x1, w) + 7) >> 3;
        /* only the rows [band_y0, band_y1) of the frame are drawn by the launch this mask is for */
        const int ty0 = max ((int) y0, band_y0) >> 2, ty1 = (min ((int) y1, band_y1) + 3) >> 2;
        if (ty1 <= ty0) return;
        const int nx = tx1 - tx0, n = nx * (ty1 - ty0);
        if (n <= 4096)
        {
            for (int j = k; j < n; j += 8) mask[(size_t) (ty0 + j / nx) * tiles_x + tx0 + j % nx] = epoch;
            return;
        }
    }
    if (k == 0) mask[mask_tiles] = epoch;   /* no bounds for this frame */
}

extern "C" int voxgpu_render_sched (const voxgpu_tree *t, const voxgpu_camera *c, const voxgpu_frame *f,
                                    const uint8_t *d_tex, uint32_t *d_pixels, int32_t *d_leaf, float *d_pts,
                                    voxgpu_stream s, voxgpu_sched *sched);

/*
 * Screen-space bounds of the tree's root box for a pinhole camera: a primary ray through pixel
 * (sx, sy) can only hit the box if (sx, sy) lies in the convex hull of the projections of the
 * box's corners, because every point of the box in front of the camera projects into that hull
 * and all points of a ray project onto its pixel.  The bounding rectangle of the eight projected
 * corners, widened by 2 pixels -- three orders of magnitude more than fp32 rounding can move
 * the reference's hit_box decision (geom.c:69-99) -- therefore contains every pixel whose ray the
 * reference lets into the tree.  If a corner is not safely in front of the camera there are no
 * bounds.  Projection = the inverse of screen2world (simple-camera.c:21-31, doom-camera.c:36-57),
 * in double.
 */
static int g_cull = !(getenv ("VOXGPU_NO_CULL") && atoi (getenv ("VOXGPU_NO_CULL")));
static int g_mask = !(getenv ("VOXGPU_NO_MASK") && atoi (getenv ("VOXGPU_NO_MASK")));
/* 0: no culling, 1: root-box bounds + screen mask (default), 2: root-box bounds only */
extern "C" void voxgpu_set_cull (int on) { g_cull = on != 0; g_mask = on == 1; }

extern "C" void voxgpu_screen_bounds (const float root_min[3], const float root_max[3], const voxgpu_camera *c,
                                      int w, int h, int out[4])
{
    out[0] = 0; out[1] = w; out[2] = 0; out[3] = h;
    if ((c->kind != VOXGPU_CAMERA_SIMPLE && c->kind != VOXGPU_CAMERA_DOOM) || !(c->mul > 0.0f)) return;
    double lo[2] = { 1e300, 1e300 }, hi[2] = { -1e300, -1e300 };
    for (int k=0; k<8; k++)
    {
        const double p[3] = { (double) ((k & 1) ? root_max[0] : root_min[0]) - c->position[0],
                              (double) ((k & 2) ? root_max[1] : root_min[1]) - c->position[1],
                              (double) ((k & 4) ? root_max[2] : root_min[2]) - c->position[2] };
        double v[3];
        if (c->kind == VOXGPU_CAMERA_SIMPLE)
        {
            /* rotate by the conjugate quaternion: v = p + w*t' + u x t', t' = 2 (u x p), u = -q.ijk */
            const double qw = c->rotation[0], u[3] = { -(double) c->rotation[1], -(double) c->rotation[2], -(double) c->rotation[3] };
            const double n2 = qw*qw + u[0]*u[0] + u[1]*u[1] + u[2]*u[2];
            if (!(n2 > 0.9999 && n2 < 1.0001)) return;       /* not a rotation: no claim */
            const double tt[3] = { 2*(u[1]*p[2] - u[2]*p[1]), 2*(u[2]*p[0] - u[0]*p[2]), 2*(u[0]*p[1] - u[1]*p[0]) };
            v[0] = p[0] + qw*tt[0] + (u[1]*tt[2] - u[2]*tt[1]);
            v[1] = p[1] + qw*tt[1] + (u[2]*tt[0] - u[0]*tt[2]);
            v[2] = p[2] + qw*tt[2] + (u[0]*tt[1] - u[1]*tt[0]);
        }
        else
        {
            /* doom: dir = (x cos - y sin, y cos + x sin, z) with ray = (x, 1, z + k): undo the yaw */
            const double cs = c->cosphi, sn = c->sinphi;
            const double n2 = cs*cs + sn*sn;
            if (!(n2 > 0.9999 && n2 < 1.0001)) return;
            v[0] = p[0]*cs + p[1]*sn;
            v[1] = p[1]*cs - p[0]*sn;
            v[2] = p[2];
        }
        const double depth = v[1];
        const double scale = fabs (p[0]) + fabs (p[1]) + fabs (p[2]);
        if (!(depth > 1e-3 * scale) || !(depth > 0.0)) return;           /* beside or behind the camera: no bounds */
        const double rx = v[0] / depth, rz = v[2] / depth - (c->kind == VOXGPU_CAMERA_DOOM ? (double) c->k : 0.0);
        const double sx = (rx + c->xsub) / c->mul, sy = (c->ysub - rz) / c->mul;
        if (!(fabs (sx) < 1e8) || !(fabs (sy) < 1e8)) return;
        if (sx < lo[0]) lo[0] = sx;
        if (sx > hi[0]) hi[0] = sx;
        if (sy < lo[1]) lo[1] = sy;
        if (sy > hi[1]) hi[1] = sy;
    }
    const double x0 = floor (lo[0]) - 2, x1 = ceil (hi[0]) + 3, y0 = floor (lo[1]) - 2, y1 = ceil (hi[1]) + 3;
    out[0] = x0 < 0 ? 0 : (x0 > w ? w : (int) x0);
    out[1] = x1 < 0 ? 0 : (x1 > w ? w : (int) x1);
    out[2] = y0 < 0 ? 0 : (y0 > h ? h : (int) y0);
    out[3] = y1 < 0 ? 0 : (y1 > h ? h : (int) y1);
}

extern "C" int voxgpu_render (const voxgpu_tree *t, const voxgpu_camera *c, const voxgpu_frame *f,
                              const uint8_t *d_tex, uint32_t *d_pixels, int32_t *d_leaf, float *d_pts,
                              voxgpu_stream s)
{
    return voxgpu_render_sched (t, c, f, d_tex, d_pixels, d_leaf, d_pts, s, NULL);
}

/* Per-device ring of screen-mask buffers for launches that bring no mask of their own, one per render
   slot (slot_acquire).  A buffer is never cleared between uses: "marked" is the slot's epoch value,
   bumped per use, and the buffer is zeroed only when the epoch wraps or the slot grows. */
#define MASK_SLOTS COUNTER_SLOTS
struct MaskSlot { unsigned char *buf; size_t cap; unsigned epoch; };
static MaskSlot g_mask_slots[64][MASK_SLOTS];

static unsigned char* mask_slot (int device, unsigned idx, size_t bytes, cudaStream_t st, unsigned *epoch)
{
    MaskSlot &m = g_mask_slots[device][idx];
    if (m.cap < bytes)
    {
        if (m.buf) { cudaDeviceSynchronize (); cudaFree (m.buf); m.buf = NULL; m.cap = 0; }
        if (cudaMalloc (&m.buf, bytes) != cudaSuccess) return NULL;
        m.cap = bytes; m.epoch = 255;              /* forces the clear below */
    }
    if (++m.epoch > 255u)
    {
        m.epoch = 1;
        if (cudaMemsetAsync (m.buf, 0, m.cap, st) != cudaSuccess) return NULL;
    }
    *epoch = m.epoch;
    return m.buf;
}

static int mask_usable (const voxgpu_tree *t, const voxgpu_camera *c)
{
    return g_cull && g_mask && t->n_cut > 1 && (c->kind == VOXGPU_CAMERA_SIMPLE || c->kind == VOXGPU_CAMERA_DOOM) &&
           c->mul > 0.0f && g_variant >= 1 && g_variant <= 2;
}

static DevCamera dev_camera (const voxgpu_camera *c)
{
    DevCamera dc;
    dc.kind = c->kind;
    dc.px = c->position[0]; dc.py = c->position[1]; dc.pz = c->position[2];
    dc.mul = c->mul; dc.xsub = c->xsub; dc.ysub = c->ysub;
    dc.qw = c->rotation[0]; dc.qi = c->rotation[1]; dc.qj = c->rotation[2]; dc.qk = c->rotation[3];
    dc.sinphi = c->sinphi; dc.cosphi = c->cosphi; dc.k = c->k;
    dc.table = (const float4*) c->dir_table;
    return dc;
}

/* ext_mask / ext_epoch: a screen mask the caller has already marked for this frame (the presenter marks
   one per frame and shares it between the bands); NULL: make one for this launch */
static int render_impl (const voxgpu_tree *t, const voxgpu_camera *c, const voxgpu_frame *f,
                        const uint8_t *d_tex, uint32_t *d_pixels, int32_t *d_leaf, float *d_pts,
                        voxgpu_stream s, voxgpu_sched *sched, const unsigned char *ext_mask, unsigned ext_epoch);

extern "C" int voxgpu_render_sched (const voxgpu_tree *t, const voxgpu_camera *c, const voxgpu_frame *f,
                                    const uint8_t *d_tex, uint32_t *d_pixels, int32_t *d_leaf, float *d_pts,
                                    voxgpu_stream s, voxgpu_sched *sched)
{
    return render_impl (t, c, f, d_tex, d_pixels, d_leaf, d_pts, s, sched, NULL, 0);
}

static int render_impl (const voxgpu_tree *t, const voxgpu_camera *c, const voxgpu_frame *f,
                        const uint8_t *d_tex, uint32_t *d_pixels, int32_t *d_leaf, float *d_pts,
                        voxgpu_stream s, voxgpu_sched *sched, const unsigned char *ext_mask, unsigned ext_epoch)
{
    if (check_tree (t, 0, "voxgpu_render")) return -1;
    if (!c || !f || !d_tex || !d_pixels) return fail ("voxgpu_render: NULL argument");
    if ((f->w & 0xf) || (f->h & 0x3) || f->w <= 0 || f->h <= 0)
        return fail ("voxgpu_render: bad geometry %dx%d (renderer.c:97-100)", f->w, f->h);
    if (f->y0 < 0 || f->y1 > f->h || f->y0 >= f->y1 || (f->y0 & 3) || (f->y1 & 3))
        return fail ("voxgpu_render: bad row band [%d,%d)", f->y0, f->y1);
    if (c->kind == VOXGPU_CAMERA_TABLE && !c->dir_table) return fail ("voxgpu_render: TABLE camera without table");

    DevCamera dc;
    dc.kind = c->kind;
    dc.px = c->position[0]; dc.py = c->position[1]; dc.pz = c->position[2];
    dc.mul = c->mul; dc.xsub = c->xsub; dc.ysub = c->ysub;
    dc.qw = c->rotation[0]; dc.qi = c->rotation[1]; dc.qj = c->rotation[2]; dc.qk = c->rotation[3];
    dc.sinphi = c->sinphi; dc.cosphi = c->cosphi; dc.k = c->k;
    dc.table = (const float4*) c->dir_table;
    DevFrame df;
    df.w = f->w; df.y0 = f->y0; df.rows = f->y1 - f->y0; df.pitch = f->pitch_pixels;
    df.amask = f->fmt.amask;
    df.rloss = f->fmt.rloss; df.gloss = f->fmt.gloss; df.bloss = f->fmt.bloss;
    df.rshift = f->fmt.rshift; df.gshift = f->fmt.gshift; df.bshift = f->fmt.bshift;
    memset (&df.lt, 0, sizeof df.lt);
    {
        int cull[4] = { 0, f->w, 0, f->h };
        if (g_cull) voxgpu_screen_bounds (t->root_min, t->root_max, c, f->w, f->h, cull);
        df.cx0 = cull[0]; df.cx1 = cull[1]; df.cy0 = cull[2]; df.cy1 = cull[3];
    }
    df.il_world = f->interleave_world > 1 ? f->interleave_world : 1;
    df.il_rank = f->interleave_world > 1 ? f->interleave_rank : 0;
    if (df.il_rank < 0 || df.il_rank >= df.il_world) return fail ("voxgpu_render: interleave rank %d of %d", df.il_rank, df.il_world);
    if (df.il_world > 1 && ((f->quality & 3u) != 0u || g_variant == 0 || g_variant >= 3))
        return fail ("voxgpu_render: interleaved tile rows need quality BEST and the default kernel variant");
    const bool lit = f->lights != NULL;
    if (lit)
    {
        if (f->lights->device != t->device) return fail ("voxgpu_render: lights live on device %d, tree on %d", f->lights->device, t->device);
        df.lt = f->lights->ref;
    }

    const unsigned slot = slot_acquire (t->device, (cudaStream_t) s);
    /* screen mask for this launch: mark from the tree's cut into the slot's buffer */
    unsigned char *mask = NULL;
    df.mask = NULL; df.mask_tiles = 0; df.mask_epoch = 1;
    if (ext_mask)
    {
        df.mask = ext_mask; df.mask_tiles = (f->w >> 3) * (f->h >> 2); df.mask_epoch = ext_epoch;
    }
    else if (mask_usable (t, c) && !(f->interleave_world > 1))
    {
        /* (not for a frame interleaved over N GPUs: every rank would mark the WHOLE mask for 1/N of the tiles --
           measured on C5 at N = 8: 1.69 ms per 8K frame with the mask, 0.68 ms without) */
        const int mtiles = (f->w >> 3) * (f->h >> 2);
        unsigned epoch = 1;
        mask = mask_slot (t->device, slot, (size_t) mtiles + 4, (cudaStream_t) s, &epoch);
        if (!mask) return fail ("voxgpu_render: cannot allocate the screen mask");
        k_mark<<<(t->n_cut * 8 + 127) / 128, 128, 0, (cudaStream_t) s>>> (tree_ref (t), t->cut, t->n_cut, dc, f->w, f->h, mask, mtiles, (unsigned char) epoch, f->y0, f->y1);
        g_launches++;
        df.mask = mask; df.mask_tiles = mtiles; df.mask_epoch = epoch;
    }

    const int ntiles = (df.w >> 3) * (df.rows >> 2);
    const int blocks = (ntiles + RENDER_BLOCK/32 - 1) / (RENDER_BLOCK/32);
    if ((f->quality & 3u) != 0u)
    {
        /* FAST / ADAPTIVE (+ ray merge): the block-sequential kernel */
        if ((f->quality & 3u) > 2u || (f->quality & 12u) > 8u || ((f->quality & 12u) && (f->quality & 3u) != 2u))
            return fail ("voxgpu_render: invalid quality %u (renderer.c:179-195)", f->quality);
        if (((size_t) d_pixels & 15) || (df.pitch & 3))
            return fail ("voxgpu_render: quality modes other than BEST need 16-byte aligned rows");
        unsigned *counters = counter_slot (t->device, slot);
        if (!counters) return fail ("voxgpu_render: cannot allocate tile counters");
        const int warps = (((df.w >> 2) + 7) >> 3) * (((df.rows >> 2) + 3) >> 2);
        int grid = (warps + RENDER_BLOCK/32 - 1) / (RENDER_BLOCK/32);
        const int cap = g_render_blocks[t->device][lit ? 6 : 2];
        if (grid > cap) grid = cap;
        cudaStream_t st = (cudaStream_t) s;
        const bool hits = d_leaf || d_pts;
        static const bool onepass = getenv ("VOXGPU_BLOCKS_ONEPASS") && atoi (getenv ("VOXGPU_BLOCKS_ONEPASS"));
        int *pre_leaf = NULL;
        float4 *pre_pts = NULL;
        if (!onepass)
        {
            /* launch 1: every root search of the band, pixel-parallel (K1 as for BEST), into stream-ordered scratch */
            const size_t npix = (size_t) df.w * df.rows;
            CU (cudaMallocAsync (&pre_leaf, npix * sizeof (int), st));
            CU (cudaMallocAsync (&pre_pts, npix * sizeof (float4), st));
            unsigned *c1 = counters;      /* the two launches run one after the other on this stream; the counters reset themselves */
            int g1 = g_render_blocks[t->device][1];
            if (g1 > blocks) g1 = blocks;
            DevSched none;
            memset (&none, 0, sizeof none);
            DevFrame plain = df;
            memset (&plain.lt, 0, sizeof plain.lt);
            k_render_persistent<true, RENDER_STRIPS, false><<<g1, RENDER_BLOCK, 0, st>>> (tree_ref (t), dc, plain, d_tex, d_pixels, pre_leaf, pre_pts, c1, NULL, none);
            g_launches++;
            /* launch 2: the walk, looking those results up */
            if (lit)
                k_render_blocks<true, true, true><<<grid, RENDER_BLOCK, 0, st>>> (tree_ref (t), dc, df, f->quality, d_tex, d_pixels, d_leaf, (float4*) d_pts, counters, pre_leaf, pre_pts);
            else if (hits)
                k_render_blocks<true, false, true><<<grid, RENDER_BLOCK, 0, st>>> (tree_ref (t), dc, df, f->quality, d_tex, d_pixels, d_leaf, (float4*) d_pts, counters, pre_leaf, pre_pts);
            else
                k_render_blocks<false, false, true><<<grid, RENDER_BLOCK, 0, st>>> (tree_ref (t), dc, df, f->quality, d_tex, d_pixels, NULL, NULL, counters, pre_leaf, pre_pts);
            CU (cudaGetLastError ());
            CU (cudaFreeAsync (pre_leaf, st));
            CU (cudaFreeAsync (pre_pts, st));
        }
        else if (lit)
            k_render_blocks<true, true, false><<<grid, RENDER_BLOCK, 0, st>>> (tree_ref (t), dc, df, f->quality, d_tex, d_pixels, d_leaf, (float4*) d_pts, counters, NULL, NULL);
        else if (hits)
            k_render_blocks<true, false, false><<<grid, RENDER_BLOCK, 0, st>>> (tree_ref (t), dc, df, f->quality, d_tex, d_pixels, d_leaf, (float4*) d_pts, counters, NULL, NULL);
        else
            k_render_blocks<false, false, false><<<grid, RENDER_BLOCK, 0, st>>> (tree_ref (t), dc, df, f->quality, d_tex, d_pixels, NULL, NULL, counters, NULL, NULL);
    }
    else if (g_variant == 0 && !lit)
    {
        if (d_leaf || d_pts)
            k_render<true><<<blocks, RENDER_BLOCK, 0, (cudaStream_t) s>>> (tree_ref (t), dc, df, d_tex, d_pixels, d_leaf, (float4*) d_pts);
        else
            k_render<false><<<blocks, RENDER_BLOCK, 0, (cudaStream_t) s>>> (tree_ref (t), dc, df, d_tex, d_pixels, NULL, NULL);
    }
    else
    {
        unsigned *counters = counter_slot (t->device, slot);
        if (!counters) return fail ("voxgpu_render: cannot allocate tile counters");
        const bool hits = d_leaf || d_pts;
        int grid = g_render_blocks[t->device][hits ? 1 : 0];
        if (grid > blocks) grid = blocks;
        cudaStream_t st = (cudaStream_t) s;
        float4 *pts = (float4*) d_pts;
        DevSched sch;
        memset (&sch, 0, sizeof sch);
        if (sched && g_variant == 1 && sched->t_heavy && df.il_world == 1 &&
            sched_prepare (sched, t->device, df.w, df.y0, df.rows, st, &sch)) return -1;
        if (lit)
        {   /* light manager attached: the default kernel, shading through the flattened M-tree */
            grid = g_render_blocks[t->device][5];
            if (grid > blocks) grid = blocks;
            k_render_persistent<true, RENDER_STRIPS, true><<<grid, RENDER_BLOCK, 0, st>>> (tree_ref (t), dc, df, d_tex, d_pixels, d_leaf, pts, counters, g_trace, sch);
        }
        else if (g_variant >= 3)
        {   /* per-lane refill; g_variant - 2 = refill threshold when >= 3 ... */
            const int rmin = g_refill_min;
            grid = g_render_blocks[t->device][hits ? 4 : 3];
            if (grid > blocks) grid = blocks;
            if (hits) k_render_stream<true, RENDER_STRIPS><<<grid, RENDER_BLOCK, 0, st>>> (tree_ref (t), dc, df, d_tex, d_pixels, d_leaf, pts, counters, rmin);
            else      k_render_stream<false, RENDER_STRIPS><<<grid, RENDER_BLOCK, 0, st>>> (tree_ref (t), dc, df, d_tex, d_pixels, NULL, NULL, counters, rmin);
        }
        else if (g_variant == 2)
        {   /* raster tile order, kept for A/B measurements */
            if (hits) k_render_persistent<true, 1u, false><<<grid, RENDER_BLOCK, 0, st>>> (tree_ref (t), dc, df, d_tex, d_pixels, d_leaf, pts, counters, g_trace, sch);
            else      k_render_persistent<false, 1u, false><<<grid, RENDER_BLOCK, 0, st>>> (tree_ref (t), dc, df, d_tex, d_pixels, NULL, NULL, counters, g_trace, sch);
        }
        else
        {
            if (hits) k_render_persistent<true, RENDER_STRIPS, false><<<grid, RENDER_BLOCK, 0, st>>> (tree_ref (t), dc, df, d_tex, d_pixels, d_leaf, pts, counters, g_trace, sch);
            else      k_render_persistent<false, RENDER_STRIPS, false><<<grid, RENDER_BLOCK, 0, st>>> (tree_ref (t), dc, df, d_tex, d_pixels, NULL, NULL, counters, g_trace, sch);
        }
    }
    g_launches++;
    CU (cudaGetLastError ());
    slot_release (t->device, slot, (cudaStream_t) s);
    return 0;
}

/* --------------------------------------------------- frames to host memory */

/*
 * Presenter: what vox_render + vox_redraw need when the surface lives in host
 * memory (north_star item 4).  Measured on B200 (tools/probes/pcie_probe.cu,
 * tools/kbench.py): a kernel CAN store the frame straight into mapped pinned
 * memory, but those stores back up the SM's memory pipeline behind PCIe and the
 * traversal stalls with them -- the frame takes kernel time PLUS transfer time
 * (0.25 + 0.15 ms) however the tiles are ordered.  So the frame is rendered into
 * a device framebuffer in horizontal bands, one kernel per band on its own
 * stream, and every band is followed on that stream by a DMA copy into the
 * (page-locked) surface.  The streams have descending priorities: all band
 * kernels are in flight at once, the block scheduler drains them in order, the
 * tail of one band is back-filled by the next, and the copy engine moves band i
 * while band i+1 renders.  Only the last band's copy is exposed.
 */
#define PRESENT_MAX_BANDS 16

struct voxgpu_present
{
    int device, w, h, bands;
    uint32_t *fb;                         /* device framebuffer, w*h */
    cudaStream_t stream[PRESENT_MAX_BANDS];
    cudaEvent_t done[PRESENT_MAX_BANDS];
    voxgpu_sched *sched[PRESENT_MAX_BANDS];  /* heavy-first tile order, frame to frame */
    cudaEvent_t start;
    int in_flight;                        /* a frame has been queued and not yet waited for */
    int used_bands;                       /* bands of the frame in flight */
    unsigned char *mask;                  /* screen mask shared by the bands of a frame (k_mark), w/8 * h/4 + 4 bytes */
    unsigned epoch;                       /* value that means "marked" in the current frame: no clearing between frames */
    cudaEvent_t marked;
    int fixed_bands;                      /* band count was asked for explicitly: never adapt */
};

/* Presenters with a frame in flight, process-wide.  A caller that keeps several contexts
   in flight (a swap chain of surfaces) already overlaps one frame's copy-out with the next
   frame's kernels; splitting such frames into bands only adds launches (measured on C2:
   6 contexts x 1 band 12.0 Grays/s, x 2 bands 11.4; one context: 4 bands 6.0, 1 band 5.0). */
static std::atomic<int> g_presents_in_flight (0);

extern "C" void voxgpu_present_destroy (voxgpu_present *p)
{
    if (!p) return;
    for (int i=0; i<p->bands; i++)
    {
        if (p->stream[i]) { cudaStreamSynchronize (p->stream[i]); cudaStreamDestroy (p->stream[i]); }
        if (p->done[i]) cudaEventDestroy (p->done[i]);
        voxgpu_sched_destroy (p->sched[i]);
    }
    if (p->start) cudaEventDestroy (p->start);
    if (p->marked) cudaEventDestroy (p->marked);
    if (p->in_flight) g_presents_in_flight--;
    cudaFree (p->mask);
    cudaFree (p->fb);
    free (p);
}

extern "C" voxgpu_present* voxgpu_present_create (int w, int h, int bands)
{
    if (w <= 0 || h <= 0 || (w & 0xf) || (h & 0x3)) { fail ("voxgpu_present_create: bad geometry %dx%d", w, h); return NULL; }
    int fixed = bands > 0;
    if (bands <= 0)
    {
        const char *env = getenv ("VOXGPU_BANDS");
        bands = env ? atoi (env) : 4;
        fixed = env != NULL;
    }
    if (bands < 1) bands = 1;
    if (bands > PRESENT_MAX_BANDS) bands = PRESENT_MAX_BANDS;
    if (bands > h / 4) bands = h / 4;
    voxgpu_present *p = (voxgpu_present*) calloc (1, sizeof *p);
    p->w = w; p->h = h; p->bands = bands; p->fixed_bands = fixed;
    int least = 0, greatest = 0;
    cudaError_t e;
    if ((e = cudaGetDevice (&p->device)) != cudaSuccess ||
        (e = cudaDeviceGetStreamPriorityRange (&least, &greatest)) != cudaSuccess ||
        (e = cudaMalloc (&p->fb, (size_t) w * h * 4)) != cudaSuccess ||
        (e = cudaMalloc (&p->mask, (size_t) (w >> 3) * (h >> 2) + 4)) != cudaSuccess ||
        (e = cudaMemset (p->mask, 0, (size_t) (w >> 3) * (h >> 2) + 4)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags (&p->marked, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreate (&p->start)) != cudaSuccess) goto bad;
    for (int i=0; i<bands; i++)
    {
        int prio = greatest + i;            /* numerically lower = served first */
        if (prio > least) prio = least;
        if ((e = cudaStreamCreateWithPriority (&p->stream[i], cudaStreamNonBlocking, prio)) != cudaSuccess ||
            (e = cudaEventCreate (&p->done[i])) != cudaSuccess) goto bad;
        /* opt-in: measured on C2 the schedule shortens a lone kernel by 9% but the bands already
           fill each other's tails, so end to end it only adds its sorting kernel */
        if (getenv ("VOXGPU_SCHED") && atoi (getenv ("VOXGPU_SCHED"))) p->sched[i] = voxgpu_sched_create ();
    }
    return p;
bad:
    fail ("voxgpu_present_create: %s", cudaGetErrorString (e));
    voxgpu_present_destroy (p);
    return NULL;
}

extern "C" int voxgpu_present_bands (const voxgpu_present *p) { return p ? p->bands : 0; }

extern "C" int voxgpu_present_render (voxgpu_present *p, const voxgpu_tree *t, const voxgpu_camera *c,
                                      const voxgpu_frame *f, const uint8_t *d_tex,
                                      void *host_pixels, size_t host_pitch_bytes)
{
    if (!p || !f || !host_pixels) return fail ("voxgpu_present_render: NULL argument");
    if (f->w != p->w || f->h != p->h) return fail ("voxgpu_present_render: frame is %dx%d, presenter %dx%d", f->w, f->h, p->w, p->h);
    if (host_pitch_bytes < (size_t) p->w * 4) return fail ("voxgpu_present_render: pitch %zu too small", host_pitch_bytes);
    int cur = -1;
    CU (cudaGetDevice (&cur));
    if (cur != p->device) return fail ("voxgpu_present_render: presenter lives on device %d, current device is %d", p->device, cur);

    const int units = p->h / 4;             /* bands are multiples of 4 rows (renderer.c:97-100) */
    if (p->in_flight && voxgpu_present_wait (p)) return -1;   /* the previous frame owns fb and the events */
    const int others = g_presents_in_flight.load ();
    const int nbands = (!p->fixed_bands && others > 0) ? 1 : p->bands;
    CU (cudaEventRecord (p->start, p->stream[0]));
    /* one screen mask per frame, marked on the first band's stream, shared by all bands */
    const unsigned char *mask = NULL;
    if (t && mask_usable (t, c) && (f->quality & 3u) == 0u)
    {
        const int mtiles = (p->w >> 3) * (p->h >> 2);
        if (++p->epoch > 255u)
        {
            p->epoch = 1;
            CU (cudaMemsetAsync (p->mask, 0, (size_t) mtiles + 4, p->stream[0]));
        }
        k_mark<<<(t->n_cut * 8 + 127) / 128, 128, 0, p->stream[0]>>> (tree_ref (t), t->cut, t->n_cut, dev_camera (c), p->w, p->h,
                                                                    p->mask, mtiles, (unsigned char) p->epoch, 0, p->h);
        g_launches++;
        mask = p->mask;
    }
    CU (cudaEventRecord (p->marked, p->stream[0]));
    for (int i=0; i<nbands; i++)
    {
        /* the later bands wait for the mask (and with it the start mark: the elapsed time covers all of them) */
        if (i) CU (cudaStreamWaitEvent (p->stream[i], p->marked, 0));
        const int y0 = (int) ((long long) units * i / nbands) * 4;
        const int y1 = (int) ((long long) units * (i + 1) / nbands) * 4;
        if (y1 <= y0) { CU (cudaEventRecord (p->done[i], p->stream[i])); continue; }
        voxgpu_frame band = *f;
        band.y0 = y0; band.y1 = y1; band.pitch_pixels = p->w;
        uint32_t *src = p->fb + (size_t) y0 * p->w;
        if (render_impl (t, c, &band, d_tex, src, NULL, NULL, (voxgpu_stream) p->stream[i], p->sched[i], mask, p->epoch)) return -1;
        char *dst = (char*) host_pixels + (size_t) y0 * host_pitch_bytes;
        if (host_pitch_bytes == (size_t) p->w * 4)
            CU (cudaMemcpyAsync (dst, src, (size_t) (y1 - y0) * p->w * 4, cudaMemcpyDeviceToHost, p->stream[i]));
        else
            CU (cudaMemcpy2DAsync (dst, host_pitch_bytes, src, (size_t) p->w * 4, (size_t) p->w * 4, y1 - y0,
                                   cudaMemcpyDeviceToHost, p->stream[i]));
        CU (cudaEventRecord (p->done[i], p->stream[i]));
    }
    p->used_bands = nbands;
    p->in_flight = 1;
    g_presents_in_flight++;
    return 0;
}

extern "C" int voxgpu_present_wait (voxgpu_present *p)
{
    if (!p || !p->in_flight) return 0;
    for (int i=0; i<p->used_bands; i++) CU (cudaEventSynchronize (p->done[i]));
    p->in_flight = 0;
    g_presents_in_flight--;
    return 0;
}

extern "C" float voxgpu_present_last_ms (voxgpu_present *p)
{
    if (!p || !p->used_bands) return -1.0f;
    float worst = -1.0f;
    for (int i=0; i<p->used_bands; i++)
    {
        float ms = 0.0f;
        if (cudaEventSynchronize (p->done[i]) != cudaSuccess ||
            cudaEventElapsedTime (&ms, p->start, p->done[i]) != cudaSuccess) { fail ("voxgpu_present_last_ms: event query failed"); return -1.0f; }
        if (ms > worst) worst = ms;
    }
    return worst;
}

extern "C" const uint32_t* voxgpu_present_device_frame (const voxgpu_present *p) { return p ? p->fb : NULL; }
