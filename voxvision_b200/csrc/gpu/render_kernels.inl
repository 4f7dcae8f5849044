/* render_kernels.inl -- part of the single translation unit voxgpu.cu (included there, in this order):
   K1: device camera / frame / lights, shading, k_render, k_sched_build, k_render_persistent, k_render_stream, k_render_blocks. */
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
