/* render_launch.inl -- part of the single translation unit voxgpu.cu (included there, in this order):
   host side of K1: render slots, lights upload, tile schedule, k_mark + screen bounds, voxgpu_render. */
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
