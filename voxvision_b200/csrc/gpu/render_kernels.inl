/* render_kernels.inl -- part of the single translation unit voxgpu.cu (included there, in this order):
   K1: device camera / frame / lights / mask, shading, k_render_persistent, k_render_blocks. */
/* -------------------------------------------------------------- K1: render */

#define RENDER_BLOCK 128
#define RENDER_STRIPS 64u
/* tile cost feedback: node visits from which a tile counts as expensive (0 = feedback off); see DESIGN.md for the measurement */
#ifndef VOXGPU_TILE_FEEDBACK_DEFAULT
#define VOXGPU_TILE_FEEDBACK_DEFAULT 48
#endif

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
 * Screen mask of one frame (k_mark, render_launch.inl): per 8x4 tile of the WHOLE frame one word, the epoch of
 * the last frame in which the projection of a cut box touched the tile; a tile whose word is not the current
 * epoch is dark: no voxel can project into it.  Buffers are reused without clearing.  Besides the words,
 * per band of the frame (a band = the rows one render launch draws), the compact list of its marked tiles in
 * the order they were first marked.
 */
struct DevMask
{
    const unsigned *mark;       /* NULL: no mask for this launch */
    const unsigned *list;       /* this band's marked tiles (frame tile indices) */
    const unsigned *heavy;      /* those of them the last frame found expensive (tile cost feedback): handed out first */
    unsigned *ctl;              /* this band's control words: [0] listed tiles, [2] heavy tiles (both zeroed again by the
                                   render kernel), [1] == epoch: a box gave no bounds, draw everything */
    unsigned char *cost;        /* per tile of the frame: node visits of its slowest lane, written for the next frame; or NULL */
    unsigned epoch;
};

/*
 * K1: one warp renders an 8x4 pixel tile (coherent rays, 32-byte row segments on the store).
 * Every pixel of the band is written, 0 on a miss: that is vox_redraw's memset of the squares
 * (renderer.c:324) folded into the store.
 *
 * PERSISTENT: 70-80% of the 8x4 tiles of a C2 frame show no voxel (half of them do not even see the
 * root box) and finish at once, 1% take 70-170 us.  The grid is sized to the resident capacity of the
 * GPU and every warp pulls tickets from a global counter until the frame is done.  The counter resets
 * itself: the last warp to leave zeroes it, so no memset precedes the launch.
 *
 * TICKET ORDER.  With a screen mask the tickets run through the band's marked tiles, one per ticket -- (H) first
 * those in whose 3x3 neighbourhood the previous frame of this tree and geometry met an expensive tile (tile
 * cost feedback: every traversed tile records the node visits of its slowest lane), (M) then the others --
 * and then (D) through the band's tile rows in groups of 32 tiles per ticket, one tile per lane, where the
 * marked tiles are skipped (done already) and the dark ones are cleared.  The r01 profile showed a lone
 * launch with its SMs busy only 75-80 % of the elapsed time: the longest tiles of a frame (60-80 node
 * visits of dependent loads and arithmetic, ~0.1 ms) were handed out wherever the raster order put them,
 * and the kernel ended when the last of them did.  Now everything that traverses starts first and the
 * dark tiles -- 75 % of a skull frame, stores only -- fill the tail, 2 000 tickets instead of 48 000
 * on the one counter.  Without a mask (foreign cameras, culling off, a box behind the camera) all tiles
 * are drawn, one per ticket, in strip-interleaved order.  Results do not depend on the order.
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

/* Resident CTAs per SM the compiler must allow for K1.  Measured on C2 (tools/ab_variants.sh):
   7 (72 registers, what ptxas picks on its own) 12.31 Grays/s, 8 (61 registers, no spills) 12.48,
   9 (56 registers, 12 bytes spilled) 12.01. */
#ifndef VOXGPU_RENDER_MIN_BLOCKS
#define VOXGPU_RENDER_MIN_BLOCKS 8
#endif
template <bool OUT_HITS, unsigned STRIPS, bool LIT>
__global__ void __launch_bounds__(RENDER_BLOCK, VOXGPU_RENDER_MIN_BLOCKS)
k_render_persistent (TreeRef tr, unsigned start, DevCamera cam, DevFrame fr, DevMask mk, const unsigned char * __restrict__ tex,
                     unsigned * __restrict__ pixels, int * __restrict__ leaf_ids, float4 * __restrict__ points,
                     unsigned *counters, unsigned long long *trace)
{
    const int tiles_x = fr.w >> 3;
    /* tile rows of the band this launch owns (all of them unless the frame is interleaved over GPUs) */
    const int band_rows = fr.rows >> 2;
    const int own_rows = fr.il_rank < band_rows ? (band_rows - fr.il_rank + fr.il_world - 1) / fr.il_world : 0;
    const unsigned ntiles = (unsigned) (tiles_x * own_rows);
    const int lane = threadIdx.x & 31;
    const unsigned total_warps = gridDim.x * (RENDER_BLOCK/32);
    VOXGPU_DECLARE_STACK (stack);
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

    /* No mask: tiles are handed out strip-interleaved, consecutive tickets land in STRIPS different
       horizontal strips of the band, so that stores and traversal overlap for the whole kernel instead of
       coming in bursts (the empty top and bottom of a frame only store). */
    const unsigned per_strip = (ntiles + STRIPS - 1) / STRIPS;
    unsigned n_listed = 0, n_heavy = 0;
    bool masked = false;
    if (mk.mark)
    {
        masked = mk.ctl[1] != mk.epoch;
        if (masked) { n_heavy = mk.ctl[2]; n_listed = n_heavy + mk.ctl[0]; }
    }
    const unsigned groups_x = ((unsigned) tiles_x + 31u) >> 5;          /* phase D: 32 tiles of a row per ticket */
    const unsigned tickets = masked ? n_listed + groups_x * (unsigned) own_rows : per_strip * STRIPS;
    const int band_ty0 = fr.y0 >> 2;

    for (;;)
    {
        unsigned ticket = 0;
        if (lane == 0) ticket = atomicAdd (&counters[0], 1u);
        ticket = __shfl_sync (0xffffffffu, ticket, 0);
        if (ticket >= tickets) break;

        int tx, ty;                     /* the tile, in tiles of the whole frame */
        if (ticket < n_listed)
        {
            const unsigned gt = ticket < n_heavy ? mk.heavy[ticket] : mk.list[ticket - n_heavy];
            tx = (int) (gt % (unsigned) tiles_x); ty = (int) (gt / (unsigned) tiles_x);
        }
        else if (masked)
        {
            /* phase D: lane l looks at tile 32*g + l of one tile row and clears it unless it is listed */
            const unsigned t = ticket - n_listed;
            const int gx = (int) ((t % groups_x) << 5) + lane;
            const int gy = band_ty0 + (int) (t / groups_x) * fr.il_world + fr.il_rank;
            if (gx < tiles_x && mk.mark[(size_t) gy * tiles_x + gx] != mk.epoch)
            {
                const int ry0 = (gy << 2) - fr.y0;
                #pragma unroll
                for (int r = 0; r < 4; r++)
                {
                    unsigned *row = pixels + (size_t) (ry0 + r) * fr.pitch + (gx << 3);
                    if ((((size_t) row) & 15) == 0)
                    {
                        ((uint4*) row)[0] = make_uint4 (0u, 0u, 0u, 0u);
                        ((uint4*) row)[1] = make_uint4 (0u, 0u, 0u, 0u);
                    }
                    else for (int c = 0; c < 8; c++) row[c] = 0u;
                    if (OUT_HITS && leaf_ids)
                    {
                        int *lrow = leaf_ids + (size_t) (ry0 + r) * fr.w + (gx << 3);
                        for (int c = 0; c < 8; c++) lrow[c] = -1;
                    }
                }
            }
            continue;
        }
        else
        {
            const unsigned tile = (ticket % STRIPS) * per_strip + ticket / STRIPS;
            if (tile >= ntiles) continue;
            tx = (int) (tile % (unsigned) tiles_x);
            ty = band_ty0 + (int) (tile / (unsigned) tiles_x) * fr.il_world + fr.il_rank;
        }
        const int sx = (tx << 3) + (lane & 7);
        const int sy = (ty << 2) + (lane >> 3);
        const int ry = sy - fr.y0;                      /* row inside the band */
        if (!masked)
        {
            /* the whole 8x4 tile lies outside the root box's screen bounds: every ray misses
               (search.c:22-27 at the root), the pixels are the zeros of vox_redraw's memset */
            const int tx0 = tx << 3, ty0 = ty << 2;
            if (tx0 >= fr.cx1 || tx0 + 8 <= fr.cx0 || ty0 >= fr.cy1 || ty0 + 4 <= fr.cy0)
            {
                pixels[(size_t) ry * fr.pitch + sx] = 0u;
                if (OUT_HITS && leaf_ids) leaf_ids[(size_t) ry * fr.w + sx] = -1;
                continue;
            }
        }

        unsigned long long t0 = 0;
        if (trace) t0 = globaltimer_ns ();
        const V3 dir = screen2world (cam, sx, sy, fr.w);
        RayResult r;
        const int iterations = ray_tree_warp (tr, start, origin, dir, true, stack, r);

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
        if (mk.cost && lane == 0) mk.cost[(size_t) ty * tiles_x + tx] = (unsigned char) min (iterations, 255);
        if (trace && lane == 0)
        {
            /* development trace (voxgpu_set_trace): when and where each tile ran, per tile of the FRAME */
            const size_t gt = (size_t) ty * tiles_x + tx;
            trace[3*gt] = t0;
            trace[3*gt + 1] = globaltimer_ns ();
            trace[3*gt + 2] = ((unsigned long long) smid () << 32) | (unsigned) iterations;
        }
    }
    if (lane == 0)
    {
        const unsigned done = atomicAdd (&counters[1], 1u);
        if (done + 1 == total_warps)
        {
            counters[0] = 0; counters[1] = 0;
            if (mk.mark) { mk.ctl[0] = 0; mk.ctl[2] = 0; }      /* for the next k_list into this buffer */
            __threadfence ();
        }
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
k_render_blocks (TreeRef tr, unsigned root, DevCamera cam, DevFrame fr, unsigned quality, const unsigned char * __restrict__ tex,
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
    VOXGPU_DECLARE_STACK (stack);
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
            unsigned start = root;
            if (step == 0) want = valid;
            else if (step == 1) want = valid && c1;
            else if (phase_leaf) { want = in_range && !merge && blk_mode == 1u && leaf >= 0; start = leaf >= 0 ? (unsigned) leaf : root; }
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
