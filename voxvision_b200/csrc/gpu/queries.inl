/* queries.inl -- part of the single translation unit voxgpu.cu (included there, in this order):
   K2 k_rays, K3 k_balls and the host-buffer pipeline behind voxgpu_rays_host / voxgpu_balls_host. */
/* ---------------------------------------------------------------- K2: rays */

#define RAYS_BLOCK 128

/* Resident CTAs per SM for K2.  C3 (tools/ab_rays.sh): 7 (66 registers) 1446 Mrays/s, 8 (64 registers,
   40 bytes spilled) 1558, 9 1493, 10 1503, 12 1414: the random rays of C3 wait on HBM, one more CTA of
   latency hiding pays for the spill. */
#ifndef VOXGPU_RAYS_MIN_BLOCKS
#define VOXGPU_RAYS_MIN_BLOCKS 8
#endif
/* PERM: thread p searches ray perm[p] (entry order, sort.inl) and writes its result to the ray's own position */
template <bool STATS, bool PERM>
__global__ void __launch_bounds__(RAYS_BLOCK, VOXGPU_RAYS_MIN_BLOCKS)
k_rays (TreeRef tr, unsigned start, const float4 * __restrict__ origins, const float4 * __restrict__ dirs,
        size_t n, int * __restrict__ leaf_ids, float4 * __restrict__ points, int * __restrict__ voxels,
        unsigned long long *stats, const unsigned * __restrict__ perm)
{
    const size_t p = (size_t) blockIdx.x * RAYS_BLOCK + threadIdx.x;
    const bool live = p < n;            /* the production traversal is warp-synchronous: nobody leaves early */
    const size_t i = (PERM && live) ? (size_t) __ldg (perm + p) : p;
    VOXGPU_DECLARE_STACK (stack);
    const float4 o = live ? __ldg (origins + i) : make_float4 (0.0f, 0.0f, 0.0f, 0.0f);
    const float4 d = live ? __ldg (dirs + i) : make_float4 (1.0f, 1.0f, 1.0f, 0.0f);
    RayResult r;
    RayCounters c = {0, 0, 0, 0, 0, 0, 0};
    if (STATS) { if (live) ray_tree<true> (tr, start, v3 (o), v3 (d), stack.loc, r, &c); }
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

/*
 * K2 with per-LANE refill, the form large batches take.  In k_rays a warp keeps its 32 rays until the
 * slowest has finished: on C3 (every ray hits after ~20 node visits, but anywhere between 3 and 80) the
 * average instruction ran with 13 of 32 lanes, and the kernel is bound by instruction issue (ncu, r02:
 * 66-73 % of the issue slots busy, DRAM at 13 %), so the idle lanes are the loss.  Here the warps are
 * persistent and draw rays from a global counter in batches; whenever at least `refill_min` lanes of a
 * warp are idle, they retire their finished rays (results go to the rays' own positions), take the next
 * rays of the warp's batch, and join the search at the start node while the other lanes carry on where they
 * were.  The traversal itself is warp_search_step, unchanged: the result of a ray does not depend on which
 * lane or in which company it is searched.  Combined with the entry order of sort.inl (PERM) the rays a
 * warp takes one after the other are neighbours, so the refilled lanes walk through nodes the warp has
 * just had in L1.
 */
#define RAYS_STREAM_BATCH 256u
template <bool PERM>
__global__ void __launch_bounds__(RAYS_BLOCK, VOXGPU_RAYS_MIN_BLOCKS)
k_rays_stream (TreeRef tr, unsigned start, const float4 * __restrict__ origins, const float4 * __restrict__ dirs,
               unsigned n, int * __restrict__ leaf_ids, float4 * __restrict__ points, int * __restrict__ voxels,
               const unsigned * __restrict__ perm, unsigned *counters, int refill_min)
{
    const int lane = threadIdx.x & 31;
    const unsigned total_warps = gridDim.x * (RAYS_BLOCK/32);
    VOXGPU_DECLARE_STACK (stack);
    LaneSearch s;
    V3 zero; zero.x = zero.y = zero.z = 0.0f;
    lane_search_begin (s, start, zero, false);
    V3 one; one.x = one.y = one.z = 1.0f;
    RayK k = ray_constants (one);
    unsigned mine = 0xffffffffu;        /* the ray this lane is working on (position in the caller's arrays) */
    /* warp-uniform: the warp's batch of positions [next, next + have) in search order, and whether the
       global counter has run out.  One atomic per RAYS_STREAM_BATCH rays. */
    unsigned next = 0, have = 0;
    bool dry = false;

    for (;;)
    {
        unsigned idle = __ballot_sync (0xffffffffu, s.cur == VOXGPU_DEAD);
        if ((have || !dry) && __popc (idle) >= refill_min)
        {
            if (s.cur == VOXGPU_DEAD && mine != 0xffffffffu)
            {
                RayResult r;
                lane_search_result (s, stack, r);
                if (leaf_ids) leaf_ids[mine] = r.leaf;
                if (voxels) voxels[mine] = r.voxel;
                if (points && r.leaf >= 0) points[mine] = make_float4 (r.point.x, r.point.y, r.point.z, 0.0f);
                mine = 0xffffffffu;
            }
            if (have == 0)
            {
                if (lane == 0) next = atomicAdd (&counters[0], RAYS_STREAM_BATCH);
                next = __shfl_sync (0xffffffffu, next, 0);
                if (next >= n) dry = true; else have = min (RAYS_STREAM_BATCH, n - next);
            }
            const unsigned take = min ((unsigned) __popc (idle), have);
            const unsigned rank = (unsigned) __popc (idle & ((1u << lane) - 1u));
            if (s.cur == VOXGPU_DEAD && rank < take)
            {
                const unsigned p = next + rank;
                mine = PERM ? __ldg (perm + p) : p;
                const float4 o = __ldg (origins + mine), d = __ldg (dirs + mine);
                k = ray_constants (v3 (d));
                lane_search_begin (s, start, v3 (o), true);
            }
            next += take; have -= take;
            idle = __ballot_sync (0xffffffffu, s.cur == VOXGPU_DEAD);
        }
        if (idle == 0xffffffffu)
        {
            if (!have && dry) break;
            continue;                   /* the batch ran out between two refills */
        }
        warp_search_step (tr, k, s, stack, lane);
    }
    if (mine != 0xffffffffu)
    {
        RayResult r;
        lane_search_result (s, stack, r);
        if (leaf_ids) leaf_ids[mine] = r.leaf;
        if (voxels) voxels[mine] = r.voxel;
        if (points && r.leaf >= 0) points[mine] = make_float4 (r.point.x, r.point.y, r.point.z, 0.0f);
    }
    if (lane == 0)
    {
        const unsigned done = atomicAdd (&counters[1], 1u);
        if (done + 1 == total_warps) { counters[0] = 0; counters[1] = 0; __threadfence (); }
    }
}

/* -1: batches of RAYS_SORT_AUTO_MIN rays and more are searched in entry order (default), 0: never, 1: always */
#define RAYS_SORT_AUTO_MIN (1u << 18)
#ifndef VOXGPU_RAYS_REFILL_DEFAULT
#define VOXGPU_RAYS_REFILL_DEFAULT 8
#endif
static int g_rays_sort = getenv ("VOXGPU_RAYS_SORT") ? atoi (getenv ("VOXGPU_RAYS_SORT")) : -1;
extern "C" void voxgpu_set_rays_sort (int mode) { g_rays_sort = mode; }
/* batches searched in entry order: lanes idle before a warp refills (k_rays_stream); 0 = one ray per thread (k_rays) */
static int g_rays_refill = getenv ("VOXGPU_RAYS_REFILL") ? atoi (getenv ("VOXGPU_RAYS_REFILL")) : VOXGPU_RAYS_REFILL_DEFAULT;
extern "C" void voxgpu_set_rays_refill (int lanes) { g_rays_refill = lanes < 0 ? 0 : (lanes > 32 ? 32 : lanes); }

static int rays_stream_launch (const voxgpu_tree *t, uint32_t start, const float *d_o, const float *d_d, size_t n,
                               int32_t *d_leaf, float *d_pts, int32_t *d_vox, const unsigned *perm, cudaStream_t st);

extern "C" int voxgpu_rays_device (const voxgpu_tree *t, uint32_t start, const float *d_o, const float *d_d,
                                   size_t n, int32_t *d_leaf, float *d_pts, int32_t *d_vox, voxgpu_stream s)
{
    if (check_tree (t, start, "voxgpu_rays_device")) return -1;
    if (n == 0) return 0;
    const size_t blocks = (n + RAYS_BLOCK - 1) / RAYS_BLOCK;
    if (blocks > 0x7fffffffull) return fail ("voxgpu_rays_device: too many rays");
    cudaStream_t st = (cudaStream_t) s;
    const bool sorted = n < 0x7fffffffull && (g_rays_sort > 0 || (g_rays_sort < 0 && n >= RAYS_SORT_AUTO_MIN));
    if (sorted)
    {
        /* entry order (sort.inl): the scratch belongs to this call until the search is queued behind the sort */
        SortScratch &sc = g_sort[t->device];
        std::lock_guard<std::mutex> guard (sc.lock);
        const unsigned *perm = NULL;
        if (sort_rays (t, start, (const float4*) d_o, (const float4*) d_d, n, st, &perm)) return -1;
        int rc = 0;
        if (g_rays_refill > 0) rc = rays_stream_launch (t, start, d_o, d_d, n, d_leaf, d_pts, d_vox, perm, st);
        else
        {
            k_rays<false, true><<<(unsigned) blocks, RAYS_BLOCK, 0, st>>> (
                tree_ref (t), start, (const float4*) d_o, (const float4*) d_d, n, d_leaf, (float4*) d_pts, d_vox, NULL, perm);
            g_launches++;
            const cudaError_t e = cudaGetLastError ();
            if (e != cudaSuccess) rc = fail ("voxgpu_rays_device: launch failed: %s", cudaGetErrorString (e));
        }
        sort_release (t, st);
        return rc;
    }
    k_rays<false, false><<<(unsigned) blocks, RAYS_BLOCK, 0, st>>> (
        tree_ref (t), start, (const float4*) d_o, (const float4*) d_d, n, d_leaf, (float4*) d_pts, d_vox, NULL, NULL);
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
        k_rays<true, false><<<(unsigned) blocks, RAYS_BLOCK>>> (
            tree_ref (t), start, (const float4*) d_o, (const float4*) d_d, n, d_leaf, (float4*) d_pts, d_vox, d_st, NULL);
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
