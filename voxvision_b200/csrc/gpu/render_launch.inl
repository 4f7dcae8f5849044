/* render_launch.inl -- part of the single translation unit voxgpu.cu (included there, in this order):
   host side of K1: per-device render state (slots), lights upload, k_mark + screen bounds, voxgpu_render. */

static unsigned long long *g_trace = NULL;
static int g_variant = 1;

extern "C" void voxgpu_set_trace (void *d_trace) { g_trace = (unsigned long long*) d_trace; }
extern "C" void voxgpu_set_variant (int v) { g_variant = v == 2 ? 2 : 1; }

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
    if (!l || !host) { fail ("voxgpu_lights_upload: out of memory"); free (l); free (host); return NULL; }
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

/* ------------------------------------------------------- screen-mask buffers */

#define MASK_MAX_BANDS 16
#define MASK_CTL_WORDS (4 * MASK_MAX_BANDS)

/* One frame's screen mask (see DevMask): mark[mtiles] | list[mtiles] | heavy[mtiles] | ctl[MASK_CTL_WORDS].
   Zeroed when allocated and when the epoch wraps; otherwise never cleared. */
struct MaskBuf
{
    unsigned *buf;
    size_t tiles;               /* capacity */
    unsigned epoch;
};

static void maskbuf_free (MaskBuf &m) { if (m.buf) cudaFree (m.buf); m.buf = NULL; m.tiles = 0; m.epoch = 0; }

/* make room for `mtiles` tiles and start a new epoch; everything is queued on `st` */
static int maskbuf_begin (MaskBuf &m, size_t mtiles, cudaStream_t st)
{
    if (m.tiles < mtiles)
    {
        if (m.buf) { CU (cudaDeviceSynchronize ()); cudaFree (m.buf); m.buf = NULL; m.tiles = 0; }
        CU (cudaMalloc (&m.buf, (3 * mtiles + MASK_CTL_WORDS) * sizeof (unsigned)));
        m.tiles = mtiles; m.epoch = 0xffffffffu;    /* forces the clear below */
    }
    if (++m.epoch == 0u)
    {
        m.epoch = 1;
        CU (cudaMemsetAsync (m.buf, 0, (3 * m.tiles + MASK_CTL_WORDS) * sizeof (unsigned), st));
    }
    return 0;
}

/* after a failure between k_mark and the render kernel that would have consumed (and reset) the lists */
static void maskbuf_reset_lists (MaskBuf &m, cudaStream_t st)
{
    if (m.buf) cudaMemsetAsync (m.buf + 3 * m.tiles, 0, MASK_CTL_WORDS * sizeof (unsigned), st);
}

static DevMask maskbuf_band (const MaskBuf &m, int band, size_t first_tile, unsigned char *cost)
{
    DevMask k;
    k.mark = m.buf; k.list = m.buf + m.tiles + first_tile; k.heavy = m.buf + 2 * m.tiles + first_tile;
    k.ctl = m.buf + 3 * m.tiles + 4 * band;
    k.epoch = m.epoch;
    k.cost = cost;
    return k;
}

/* --------------------------------------------------- per-device render state */

/* A render launch borrows a SLOT of its device: a pair of self-resetting tile counters and a screen-mask
   buffer; the quality modes other than BEST also borrow one of a few SCRATCH buffers for the root searches
   (20 B per pixel of the band: too big to keep one per slot).  Launches rotate through
   RENDER_SLOTS slots; should a slot come round again while its previous borrower is still queued or
   running on ANOTHER stream, the new stream waits for it (one event per slot), so sharing is always
   ordered.  (16 slots: a slot's mask buffer is 8 B per 8x4 tile of the frame -- 8 MB for an 8K frame -- and
   is allocated when a slot first sees a frame that size.)  vox_render may be called from any thread (Manual.md:171-175): everything here is guarded by
   the device's mutex, which is held from slot acquisition until the launch is queued and the slot's
   event recorded. */
#define RENDER_SLOTS 16
struct RenderSlot
{
    cudaEvent_t ev;
    cudaStream_t last;
    int used;
    MaskBuf mask;
};

#define RENDER_SCRATCHES 4
struct RenderScratch
{
    cudaEvent_t ev;
    cudaStream_t last;
    int used;
    void *buf;                  /* 20 B per pixel of a band: hit points, then leaf ids */
    size_t bytes;
};

struct RenderDevice
{
    std::mutex lock;
    int ready;
    unsigned seq;
    unsigned *counters;         /* 2 per slot */
    int sms;
    int blocks[6];              /* resident grid of: K1, K1 with hits, K1 lit, K1b, K1b lit, streaming K2 */
    RenderSlot slot[RENDER_SLOTS];
    unsigned scratch_seq;
    RenderScratch scratch[RENDER_SCRATCHES];
};
static RenderDevice g_rdev[64];

enum { GRID_K1 = 0, GRID_K1_HITS, GRID_K1_LIT, GRID_K1B, GRID_K1B_LIT, GRID_RAYS_STREAM };

/* call with the device's lock held */
static int render_device_init (RenderDevice &rd, int device)
{
    if (rd.ready) return 0;
    unsigned *p = NULL;
    CU (cudaMalloc (&p, RENDER_SLOTS * 2 * sizeof (unsigned)));
    cudaError_t e = cudaMemset (p, 0, RENDER_SLOTS * 2 * sizeof (unsigned));
    if (e != cudaSuccess) { cudaFree (p); return fail ("cudaMemset of the tile counters failed: %s", cudaGetErrorString (e)); }
    int sms = 0, b = 0;
    cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, device);
#define RESIDENT(slot_, kernel_) do { b = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor (&b, kernel_, RENDER_BLOCK, 0); \
                                      rd.blocks[slot_] = sms * (b > 0 ? b : 1); } while (0)
    RESIDENT (GRID_K1, (k_render_persistent<false, RENDER_STRIPS, false>));
    RESIDENT (GRID_K1_HITS, (k_render_persistent<true, RENDER_STRIPS, false>));
    RESIDENT (GRID_K1_LIT, (k_render_persistent<true, RENDER_STRIPS, true>));
    RESIDENT (GRID_K1B, (k_render_blocks<true, false, true>));
    RESIDENT (GRID_K1B_LIT, (k_render_blocks<true, true, true>));
#undef RESIDENT
    b = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor (&b, k_rays_stream<true>, RAYS_BLOCK, 0);
    rd.blocks[GRID_RAYS_STREAM] = sms * (b > 0 ? b : 1);
    rd.sms = sms;
    rd.counters = p;
    rd.ready = 1;
    return 0;
}

/* K2 in its streaming form borrows a slot's pair of self-resetting counters like a render launch does */
static int rays_stream_launch (const voxgpu_tree *t, uint32_t start, const float *d_o, const float *d_d, size_t n,
                               int32_t *d_leaf, float *d_pts, int32_t *d_vox, const unsigned *perm, cudaStream_t st)
{
    RenderDevice &rd = g_rdev[t->device];
    std::lock_guard<std::mutex> guard (rd.lock);
    if (render_device_init (rd, t->device)) return -1;
    const unsigned idx = rd.seq++ % RENDER_SLOTS;
    RenderSlot &sl = rd.slot[idx];
    if (!sl.ev) CU (cudaEventCreateWithFlags (&sl.ev, cudaEventDisableTiming));
    if (sl.used && sl.last != st) CU (cudaStreamWaitEvent (st, sl.ev, 0));
    const size_t want = (n + RAYS_BLOCK - 1) / RAYS_BLOCK;
    int grid = rd.blocks[GRID_RAYS_STREAM];
    if ((size_t) grid > want) grid = (int) want;
    if (perm)
        k_rays_stream<true><<<grid, RAYS_BLOCK, 0, st>>> (tree_ref (t), start, (const float4*) d_o, (const float4*) d_d, (unsigned) n,
                                                         d_leaf, (float4*) d_pts, d_vox, perm, rd.counters + 2 * idx, g_rays_refill);
    else
        k_rays_stream<false><<<grid, RAYS_BLOCK, 0, st>>> (tree_ref (t), start, (const float4*) d_o, (const float4*) d_d, (unsigned) n,
                                                          d_leaf, (float4*) d_pts, d_vox, NULL, rd.counters + 2 * idx, g_rays_refill);
    g_launches++;
    const cudaError_t e = cudaGetLastError ();
    if (cudaEventRecord (sl.ev, st) == cudaSuccess) { sl.last = st; sl.used = 1; }
    if (e != cudaSuccess) return fail ("voxgpu_rays_device: launch failed: %s", cudaGetErrorString (e));
    return 0;
}

/* ------------------------------------------------------------ k_mark + bounds */

/*
 * Screen mask.  Every voxel lies in the box of one node of the tree's cut (voxgpu_tree.cut); a ray that
 * hits the voxel passes through that box, so its pixel lies in the convex hull of the box's projected
 * corners (see voxgpu_screen_bounds for the argument) and therefore in their bounding rectangle.  Eight
 * lanes per cut node project its box (fp32 here: the error is ~1e-3 pixel, the rectangle is widened by 2)
 * and mark the 8x4 tiles the rectangle touches with the frame's epoch -- only tiles of rows the render
 * launches of this mask will draw ([y0, y1), and of those the tile rows r with r % il_world == il_rank).
 * A tile nobody marked cannot show a voxel: the render kernel clears it without casting its rays --
 * on a C2 frame half of the primary rays enter the root box but only a fifth hit a voxel, and the ones in
 * between cost 40-50 % of all node fetches.  A box not safely in front of the camera, or covering more
 * than 4096 tiles, raises the "no bounds" flag of every band instead (the frame is then rendered without
 * the mask).  k_list then compacts the marked tiles of every band into that band's list (see
 * k_render_persistent for what the list is for).
 * (Tried and dropped: building the lists inside k_mark with atomics -- the first box to reach a tile
 * appending it -- made k_mark the longest kernel of the frame: 17 000 appends on ONE counter, where the
 * stores below wait for nobody; and a second list of "heavy" tiles, many boxes along the line of sight,
 * handed out first: on C2 the number of boxes over a tile does not predict its cost, correlation
 * 0.02-0.14, profiles/r02_tile_trace.txt.)
 */
struct MarkBands
{
    int y0, y1;                 /* frame rows the launches of this mask draw */
    int nbands;                 /* [y0, y1) is cut into nbands bands of whole 4-row units: band i starts at unit (units*i)/nbands */
    int il_world, il_rank;      /* only tile rows r (counted from y0) with r % il_world == il_rank */
};

__global__ void k_mark (TreeRef tr, const unsigned * __restrict__ cut, unsigned n_cut, DevCamera cam,
                        int w, int h, unsigned *buf, unsigned mtiles, unsigned epoch, MarkBands mb)
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
    unsigned *mark = buf, *ctl = buf + 3 * (size_t) mtiles;
    if (!anybad)
    {
        const float x0 = floorf (lox) - 2.0f, x1 = ceilf (hix) + 3.0f, y0 = floorf (loy) - 2.0f, y1 = ceilf (hiy) + 3.0f;
        if (x1 <= 0.0f || y1 <= 0.0f || x0 >= (float) w || y0 >= (float) h) return;       /* off screen */
        const int tiles_x = w >> 3;
        const int tx0 = max ((int) x0, 0) >> 3, tx1 = (min ((int) x1, w) + 7) >> 3;
        const int ty0 = max ((int) y0, mb.y0) >> 2, ty1 = (min ((int) y1, mb.y1) + 3) >> 2;
        if (ty1 <= ty0) return;
        const int nx = tx1 - tx0;
        if (nx * (ty1 - ty0) <= 4096)
        {
            /* only the tile rows this launch draws: with the frame split over N GPUs every N-th row, counted from y0 */
            const int stride = mb.il_world > 1 ? mb.il_world : 1;
            const int first = ty0 + (stride > 1 ? ((mb.il_rank - (ty0 - (mb.y0 >> 2))) % stride + stride) % stride : 0);
            const int n = first < ty1 ? nx * ((ty1 - first + stride - 1) / stride) : 0;
            for (int j = k; j < n; j += 8)
                mark[(size_t) (first + (j / nx) * stride) * tiles_x + tx0 + j % nx] = epoch;    /* a plain store: nobody waits for it */
            return;
        }
    }
    if (k == 0) for (int b = 0; b < mb.nbands; b++) ctl[4*b + 1] = epoch;   /* no bounds for this frame */
}

/* The marked tiles of every band, compacted: one warp looks at 32 tiles of one tile row (one coalesced load),
   and appends the marked ones to the lists of the row's band with ONE atomicAdd per list (ballot + popc ranks).
   Tile cost feedback: a marked tile goes to the band's HEAVY list if the previous frame met a tile of at least
   `heavy_min` node visits in its 3x3 neighbourhood (which tile of a silhouette is the expensive one changes with
   the slightest camera move; that its neighbourhood is does not). */
__global__ void k_list (unsigned *buf, unsigned mtiles, unsigned epoch, int tiles_x, int tiles_y, MarkBands mb,
                        const unsigned char * __restrict__ cost, unsigned heavy_min)
{
    const unsigned *mark = buf;
    unsigned *list = buf + mtiles, *heavy = buf + 2 * (size_t) mtiles, *ctl = buf + 3 * (size_t) mtiles;
    const int lane = threadIdx.x & 31;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned groups_x = ((unsigned) tiles_x + 31u) >> 5;
    const int units = (mb.y1 - mb.y0) >> 2, u0 = mb.y0 >> 2;
    const int il_world = mb.il_world > 1 ? mb.il_world : 1, il_rank = mb.il_world > 1 ? mb.il_rank : 0;
    const int r = (int) (warp / groups_x) * il_world + il_rank;           /* tile row inside [y0, y1) */
    if (r >= units) return;
    const int ty = u0 + r, tx = (int) ((warp % groups_x) << 5) + lane;
    const bool marked = tx < tiles_x && mark[(size_t) ty * tiles_x + tx] == epoch;
    bool hv = false;
    if (marked && cost)
    {
        unsigned m = 0;
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++)
            {
                const int x = tx + dx, y = ty + dy;
                if (x >= 0 && x < tiles_x && y >= 0 && y < tiles_y) m = max (m, (unsigned) cost[(size_t) y * tiles_x + x]);
            }
        hv = m >= heavy_min;
    }
    const unsigned mh = __ballot_sync (0xffffffffu, marked && hv), mn = __ballot_sync (0xffffffffu, marked && !hv);
    if (!(mh | mn)) return;
    /* band b = the i with (units*i)/nbands <= r < (units*(i+1))/nbands; its lists start at its first tile */
    const int b = mb.nbands > 1 ? (int) (((long long) (r + 1) * mb.nbands - 1) / units) : 0;
    const size_t first = (size_t) (u0 + (int) ((long long) units * b / mb.nbands)) * tiles_x;
    unsigned base_n = 0, base_h = 0;
    if (lane == 0)
    {
        if (mn) base_n = atomicAdd (ctl + 4*b, (unsigned) __popc (mn));
        if (mh) base_h = atomicAdd (ctl + 4*b + 2, (unsigned) __popc (mh));
    }
    base_n = __shfl_sync (0xffffffffu, base_n, 0);
    base_h = __shfl_sync (0xffffffffu, base_h, 0);
    const unsigned below = (1u << lane) - 1u;
    if (marked && !hv) list[first + base_n + __popc (mn & below)] = (unsigned) (ty * tiles_x + tx);
    if (marked && hv) heavy[first + base_h + __popc (mh & below)] = (unsigned) (ty * tiles_x + tx);
}

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

static int mask_usable (const voxgpu_tree *t, const voxgpu_camera *c)
{
    return g_cull && g_mask && t->n_cut > 1 && (c->kind == VOXGPU_CAMERA_SIMPLE || c->kind == VOXGPU_CAMERA_DOOM) &&
           c->mul > 0.0f;
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

/* Tile cost feedback: the cost map of (tree, frame geometry), made on first use.  Call with the device's render
   lock held.  NULL: feedback is off, or no memory for it (it is only a hint). */
static int g_feedback = getenv ("VOXGPU_TILE_FEEDBACK") ? atoi (getenv ("VOXGPU_TILE_FEEDBACK")) : VOXGPU_TILE_FEEDBACK_DEFAULT;
extern "C" void voxgpu_set_tile_feedback (int heavy_min) { g_feedback = heavy_min < 0 ? 0 : (heavy_min > 255 ? 255 : heavy_min); }

static unsigned char* cost_map (const voxgpu_tree *ct, int w, int h, cudaStream_t st)
{
    if (!g_feedback) return NULL;
    voxgpu_tree *t = (voxgpu_tree*) ct;
    const int tx = w >> 3, ty = h >> 2;
    if (t->cost && t->cost_tiles_x == tx && t->cost_tiles_y == ty) return t->cost;
    if (t->cost) { if (cudaDeviceSynchronize () != cudaSuccess) return NULL; cudaFree (t->cost); t->cost = NULL; }
    if (cudaMalloc (&t->cost, (size_t) tx * ty) != cudaSuccess) { cudaGetLastError (); t->cost = NULL; return NULL; }
    if (cudaMemsetAsync (t->cost, 0, (size_t) tx * ty, st) != cudaSuccess) { cudaFree (t->cost); t->cost = NULL; return NULL; }
    t->cost_tiles_x = tx; t->cost_tiles_y = ty;
    return t->cost;
}

/* queue k_mark + k_list for the rows `mb` of a w x h frame into `m` (maskbuf_begin has been called) */
static int mark_launch (const voxgpu_tree *t, const voxgpu_camera *c, int w, int h, const MaskBuf &m, const MarkBands &mb, cudaStream_t st,
                        const unsigned char *cost)
{
    k_mark<<<(t->n_cut * 8 + 127) / 128, 128, 0, st>>> (tree_ref (t), t->cut, t->n_cut, dev_camera (c), w, h, m.buf,
                                                       (unsigned) m.tiles, m.epoch, mb);
    const int il = mb.il_world > 1 ? mb.il_world : 1;
    const unsigned rows = (unsigned) ((((mb.y1 - mb.y0) >> 2) + il - 1) / il);
    const unsigned warps = rows * ((unsigned) ((w >> 3) + 31) >> 5);
    k_list<<<(warps + 3) / 4, 128, 0, st>>> (m.buf, (unsigned) m.tiles, m.epoch, w >> 3, h >> 2, mb, cost, (unsigned) g_feedback);
    g_launches += 2;
    CU (cudaGetLastError ());
    return 0;
}

/* ------------------------------------------------------------- voxgpu_render */

/* ext_mask: a screen mask the caller has already marked for this frame and this band (the presenter marks one
   per frame and shares it between the bands); NULL: make one for this launch */
static int render_impl (const voxgpu_tree *t, const voxgpu_camera *c, const voxgpu_frame *f,
                        const uint8_t *d_tex, uint32_t *d_pixels, int32_t *d_leaf, float *d_pts,
                        voxgpu_stream s, const DevMask *ext_mask);

extern "C" int voxgpu_render (const voxgpu_tree *t, const voxgpu_camera *c, const voxgpu_frame *f,
                              const uint8_t *d_tex, uint32_t *d_pixels, int32_t *d_leaf, float *d_pts,
                              voxgpu_stream s)
{
    return render_impl (t, c, f, d_tex, d_pixels, d_leaf, d_pts, s, NULL);
}

static int render_impl (const voxgpu_tree *t, const voxgpu_camera *c, const voxgpu_frame *f,
                        const uint8_t *d_tex, uint32_t *d_pixels, int32_t *d_leaf, float *d_pts,
                        voxgpu_stream s, const DevMask *ext_mask)
{
    /* ---- every argument is checked before anything is borrowed or queued ---- */
    if (!c || !f || !d_tex || !d_pixels) return fail ("voxgpu_render: NULL argument");
    if (check_tree (t, f->start_node, "voxgpu_render")) return -1;
    if ((f->w & 0xf) || (f->h & 0x3) || f->w <= 0 || f->h <= 0)
        return fail ("voxgpu_render: bad geometry %dx%d (renderer.c:97-100)", f->w, f->h);
    if (f->y0 < 0 || f->y1 > f->h || f->y0 >= f->y1 || (f->y0 & 3) || (f->y1 & 3))
        return fail ("voxgpu_render: bad row band [%d,%d)", f->y0, f->y1);
    if (c->kind == VOXGPU_CAMERA_TABLE && !c->dir_table) return fail ("voxgpu_render: TABLE camera without table");
    const unsigned mode = f->quality & 3u, merge = f->quality & 12u;
    if (mode > 2u || merge > 8u || (merge && mode != 2u))
        return fail ("voxgpu_render: invalid quality %u (renderer.c:179-195)", f->quality);
    const int il_world = f->interleave_world > 1 ? f->interleave_world : 1;
    const int il_rank = f->interleave_world > 1 ? f->interleave_rank : 0;
    if (il_rank < 0 || il_rank >= il_world) return fail ("voxgpu_render: interleave rank %d of %d", il_rank, il_world);
    if (il_world > 1 && mode != 0u) return fail ("voxgpu_render: interleaved tile rows need quality BEST");
    if (mode != 0u && (((size_t) d_pixels & 15) || (f->pitch_pixels & 3)))
        return fail ("voxgpu_render: quality modes other than BEST need 16-byte aligned rows");
    const bool lit = f->lights != NULL;
    if (lit && f->lights->device != t->device)
        return fail ("voxgpu_render: lights live on device %d, tree on %d", f->lights->device, t->device);

    const DevCamera dc = dev_camera (c);
    DevFrame df;
    df.w = f->w; df.y0 = f->y0; df.rows = f->y1 - f->y0; df.pitch = f->pitch_pixels;
    df.amask = f->fmt.amask;
    df.rloss = f->fmt.rloss; df.gloss = f->fmt.gloss; df.bloss = f->fmt.bloss;
    df.rshift = f->fmt.rshift; df.gshift = f->fmt.gshift; df.bshift = f->fmt.bshift;
    memset (&df.lt, 0, sizeof df.lt);
    if (lit) df.lt = f->lights->ref;
    {
        int cull[4] = { 0, f->w, 0, f->h };
        if (g_cull) voxgpu_screen_bounds (t->root_min, t->root_max, c, f->w, f->h, cull);
        df.cx0 = cull[0]; df.cx1 = cull[1]; df.cy0 = cull[2]; df.cy1 = cull[3];
    }
    df.il_world = il_world; df.il_rank = il_rank;

    cudaStream_t st = (cudaStream_t) s;
    RenderDevice &rd = g_rdev[t->device];
    std::lock_guard<std::mutex> guard (rd.lock);
    if (render_device_init (rd, t->device)) return -1;

    /* borrow a slot; its previous borrower on another stream goes first */
    const unsigned idx = rd.seq++ % RENDER_SLOTS;
    RenderSlot &sl = rd.slot[idx];
    if (!sl.ev) CU (cudaEventCreateWithFlags (&sl.ev, cudaEventDisableTiming));
    if (sl.used && sl.last != st) CU (cudaStreamWaitEvent (st, sl.ev, 0));
    unsigned *counters = rd.counters + 2 * idx;
    /* is the previous render launch of this device still at work?  (its slot's event, recorded behind it) */
    bool busy = false;
    if (idx != (idx + RENDER_SLOTS - 1) % RENDER_SLOTS)
    {
        const RenderSlot &prev = rd.slot[(idx + RENDER_SLOTS - 1) % RENDER_SLOTS];
        if (prev.used && prev.ev && cudaEventQuery (prev.ev) == cudaErrorNotReady) busy = true;
    }
    int rc = 0;
    bool marked_here = false;
    RenderScratch *scratch = NULL;
    /* whatever happens below, the slot's event is recorded behind what has been queued */
#define LEAVE(code_) do { rc = (code_); goto leave; } while (0)

    {
        /* screen mask for this launch: mark from the tree's cut into the slot's buffer */
        DevMask mk;
        memset (&mk, 0, sizeof mk);
        if (ext_mask) mk = *ext_mask;
        else if (mask_usable (t, c))
        {
            const size_t mtiles = (size_t) (f->w >> 3) * (f->h >> 2);
            if (maskbuf_begin (sl.mask, mtiles, st)) LEAVE (-1);
            MarkBands mb;
            mb.y0 = f->y0; mb.y1 = f->y1; mb.nbands = 1; mb.il_world = il_world; mb.il_rank = il_rank;
            unsigned char *cost = mode == 0u && !f->start_node ? cost_map (t, f->w, f->h, st) : NULL;
            if (mark_launch (t, c, f->w, f->h, sl.mask, mb, st, cost)) LEAVE (-1);
            marked_here = true;
            mk = maskbuf_band (sl.mask, 0, (size_t) (f->y0 >> 2) * (f->w >> 3), cost);
        }

        const int ntiles = (df.w >> 3) * (df.rows >> 2);
        const int blocks = (ntiles + RENDER_BLOCK/32 - 1) / (RENDER_BLOCK/32);
        const bool hits = d_leaf || d_pts;
        float4 *pts = (float4*) d_pts;
        const TreeRef tr = tree_ref (t);
        const unsigned start = f->start_node;
        if (mode != 0u)
        {
            /* FAST / ADAPTIVE (+ ray merge).  Launch 1: every root search of the band, pixel-parallel (K1 as for
               BEST, unlit), into the slot's scratch; launch 2: the block walk, looking those results up. */
            const size_t npix = (size_t) df.w * df.rows, need = npix * (sizeof (int) + sizeof (float4));
            RenderScratch &sc = rd.scratch[rd.scratch_seq++ % RENDER_SCRATCHES];
            if (!sc.ev && cudaEventCreateWithFlags (&sc.ev, cudaEventDisableTiming) != cudaSuccess) { fail ("voxgpu_render: cudaEventCreate failed"); LEAVE (-1); }
            if (sc.bytes < need)
            {
                if (sc.buf) { cudaError_t e = cudaDeviceSynchronize (); cudaFree (sc.buf); sc.buf = NULL; sc.bytes = 0;
                              if (e != cudaSuccess) { fail ("voxgpu_render: %s", cudaGetErrorString (e)); LEAVE (-1); } }
                cudaError_t e = cudaMalloc (&sc.buf, need);
                if (e != cudaSuccess) { fail ("voxgpu_render: scratch of %zu bytes: %s", need, cudaGetErrorString (e)); LEAVE (-1); }
                sc.bytes = need;
            }
            if (sc.used && sc.last != st && cudaStreamWaitEvent (st, sc.ev, 0) != cudaSuccess) { fail ("voxgpu_render: cudaStreamWaitEvent failed"); LEAVE (-1); }
            scratch = &sc;
            float4 *pre_pts = (float4*) sc.buf;
            int *pre_leaf = (int*) ((char*) sc.buf + npix * sizeof (float4));
            int g1 = rd.blocks[GRID_K1_HITS];
            if (g1 > blocks) g1 = blocks;
            DevFrame plain = df;
            memset (&plain.lt, 0, sizeof plain.lt);
            k_render_persistent<true, RENDER_STRIPS, false><<<g1, RENDER_BLOCK, 0, st>>> (tr, start, dc, plain, mk, d_tex, d_pixels, pre_leaf, pre_pts, counters, NULL);
            g_launches++;
            marked_here = false;            /* the lists are consumed */
            const int warps = (((df.w >> 2) + 7) >> 3) * (((df.rows >> 2) + 3) >> 2);
            int grid = (warps + RENDER_BLOCK/32 - 1) / (RENDER_BLOCK/32);
            const int cap = rd.blocks[lit ? GRID_K1B_LIT : GRID_K1B];
            if (grid > cap) grid = cap;
            /* the two launches run one after the other on this stream; the counters reset themselves */
            if (lit)
                k_render_blocks<true, true, true><<<grid, RENDER_BLOCK, 0, st>>> (tr, start, dc, df, f->quality, d_tex, d_pixels, d_leaf, pts, counters, pre_leaf, pre_pts);
            else if (hits)
                k_render_blocks<true, false, true><<<grid, RENDER_BLOCK, 0, st>>> (tr, start, dc, df, f->quality, d_tex, d_pixels, d_leaf, pts, counters, pre_leaf, pre_pts);
            else
                k_render_blocks<false, false, true><<<grid, RENDER_BLOCK, 0, st>>> (tr, start, dc, df, f->quality, d_tex, d_pixels, NULL, NULL, counters, pre_leaf, pre_pts);
        }
        else
        {
            int grid = rd.blocks[lit ? GRID_K1_LIT : (hits ? GRID_K1_HITS : GRID_K1)];
            {
                /* A lone launch takes every resident slot of the GPU (8 CTAs per SM).  One queued while another is
                   still at work takes 5 per SM: kernels in flight then share the SMs side by side instead of each
                   holding all slots until its last ticket is done -- 4 frames in flight 0.1400 -> 0.1378 ms per frame,
                   a ring of six contexts 13.7 -> 14.1 Grays/s on the same box; a lone launch loses with anything below
                   7, and so do the 5 ms launches of C5's 8K frames even in flight (4.96 -> 5.01 ms), so launches of
                   more than 4 Mi pixels keep the full grid (profiles/r02_ab_variants.txt).
                   $VOXGPU_K1_BUSY_CTAS: that 5; 0 = always the full grid. */
                static const int ctas = getenv ("VOXGPU_K1_BUSY_CTAS") ? atoi (getenv ("VOXGPU_K1_BUSY_CTAS")) : 5;
                if (busy && ctas > 0 && (long long) df.w * df.rows <= (4ll << 20) && grid > rd.sms * ctas) grid = rd.sms * ctas;
            }
            if (grid > blocks) grid = blocks;
            if (lit)      /* light manager attached: shading through the flattened M-tree */
                k_render_persistent<true, RENDER_STRIPS, true><<<grid, RENDER_BLOCK, 0, st>>> (tr, start, dc, df, mk, d_tex, d_pixels, d_leaf, pts, counters, g_trace);
            else if (g_variant == 2)
            {   /* raster tile order, kept for A/B measurements */
                if (hits) k_render_persistent<true, 1u, false><<<grid, RENDER_BLOCK, 0, st>>> (tr, start, dc, df, mk, d_tex, d_pixels, d_leaf, pts, counters, g_trace);
                else      k_render_persistent<false, 1u, false><<<grid, RENDER_BLOCK, 0, st>>> (tr, start, dc, df, mk, d_tex, d_pixels, NULL, NULL, counters, g_trace);
            }
            else
            {
                if (hits) k_render_persistent<true, RENDER_STRIPS, false><<<grid, RENDER_BLOCK, 0, st>>> (tr, start, dc, df, mk, d_tex, d_pixels, d_leaf, pts, counters, g_trace);
                else      k_render_persistent<false, RENDER_STRIPS, false><<<grid, RENDER_BLOCK, 0, st>>> (tr, start, dc, df, mk, d_tex, d_pixels, NULL, NULL, counters, g_trace);
            }
            marked_here = false;
        }
        g_launches++;
        cudaError_t e = cudaGetLastError ();
        if (e != cudaSuccess) { fail ("voxgpu_render: launch failed: %s", cudaGetErrorString (e)); LEAVE (-1); }
    }
leave:
#undef LEAVE
    if (marked_here) maskbuf_reset_lists (sl.mask, st);     /* k_mark ran, its consumer did not */
    if (scratch && cudaEventRecord (scratch->ev, st) == cudaSuccess) { scratch->last = st; scratch->used = 1; }
    if (cudaEventRecord (sl.ev, st) == cudaSuccess) { sl.last = st; sl.used = 1; }
    return rc;
}
