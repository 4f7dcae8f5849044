/* present.inl -- part of the single translation unit voxgpu.cu (included there, in this order):
   frames to host memory: voxgpu_present_* (one device) and voxgpu_multi_* (all devices of the process). */
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
 *
 * A presenter can also draw ONE SHARE of a frame that several GPUs split
 * (frame->interleave_world / interleave_rank: the 4-row tile rows r with
 * r % world == rank): one launch, then one strided DMA copy of exactly those rows
 * into the surface over this GPU's own PCIe link.  voxgpu_multi_* below drives
 * one such presenter per device.
 */
#define PRESENT_MAX_BANDS MASK_MAX_BANDS

struct voxgpu_present
{
    int device, w, h, bands;
    uint32_t *fb;                         /* device framebuffer, w*h */
    cudaStream_t stream[PRESENT_MAX_BANDS];
    cudaEvent_t done[PRESENT_MAX_BANDS];
    cudaEvent_t start;
    int in_flight;                        /* a frame has been queued and not yet waited for */
    int used_bands;                       /* bands of the frame in flight */
    MaskBuf mask;                         /* screen mask shared by the bands of a frame (k_mark) */
    cudaEvent_t marked;
    int fixed_bands;                      /* band count was asked for explicitly: never adapt */
    int *span;                            /* 4 x h/4 ints: the columns the tree can show per group of 16 rows (lo, hi) and the
                                             columns the frame in flight copies per 4-row unit (lo, hi); see visible_spans */
};

/* Presenters with a frame in flight, process-wide.  A caller that keeps several contexts
   in flight (a swap chain of surfaces) already overlaps one frame's copy-out with the next
   frame's kernels; splitting such frames into bands only adds launches (measured on C2:
   6 contexts x 1 band 12.0 Grays/s, x 2 bands 11.4; one context: 4 bands 6.0, 1 band 5.0). */
static std::atomic<int> g_presents_in_flight (0);
/* bytes of frames queued for the copy engines by all presenters of the process (bench.py's e2e.d2h_bytes_per_step) */
static std::atomic<unsigned long long> g_present_bytes (0);
extern "C" unsigned long long voxgpu_present_bytes_copied (void) { return g_present_bytes.load (); }

extern "C" void voxgpu_present_destroy (voxgpu_present *p)
{
    if (!p) return;
    int cur = -1;
    cudaGetDevice (&cur);
    if (cur != p->device) cudaSetDevice (p->device);
    for (int i=0; i<p->bands; i++)
    {
        if (p->stream[i]) { cudaStreamSynchronize (p->stream[i]); cudaStreamDestroy (p->stream[i]); }
        if (p->done[i]) cudaEventDestroy (p->done[i]);
    }
    if (p->start) cudaEventDestroy (p->start);
    if (p->marked) cudaEventDestroy (p->marked);
    if (p->in_flight) g_presents_in_flight--;
    maskbuf_free (p->mask);
    cudaFree (p->fb);
    if (cur != p->device && cur >= 0) cudaSetDevice (cur);
    free (p->span);
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
    if (!p) { fail ("voxgpu_present_create: out of memory"); return NULL; }
    p->w = w; p->h = h; p->bands = bands; p->fixed_bands = fixed;
    p->span = (int*) calloc ((size_t) (h / 4) * 4, sizeof (int));
    if (!p->span) { free (p); fail ("voxgpu_present_create: out of memory"); return NULL; }
    int least = 0, greatest = 0;
    cudaError_t e;
    if ((e = cudaGetDevice (&p->device)) != cudaSuccess ||
        (e = cudaDeviceGetStreamPriorityRange (&least, &greatest)) != cudaSuccess ||
        (e = cudaMalloc (&p->fb, (size_t) w * h * 4)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags (&p->marked, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreate (&p->start)) != cudaSuccess) goto bad;
    for (int i=0; i<bands; i++)
    {
        int prio = greatest + i;            /* numerically lower = served first */
        if (prio > least) prio = least;
        if ((e = cudaStreamCreateWithPriority (&p->stream[i], cudaStreamNonBlocking, prio)) != cudaSuccess ||
            (e = cudaEventCreate (&p->done[i])) != cudaSuccess) goto bad;
    }
    return p;
bad:
    fail ("voxgpu_present_create: %s", cudaGetErrorString (e));
    voxgpu_present_destroy (p);
    return NULL;
}

extern "C" int voxgpu_present_bands (const voxgpu_present *p) { return p ? p->bands : 0; }

/* The visible rectangle of a frame (see voxgpu_present_render): the root box's screen bounds on tile rows and on
   16-pixel columns -- whole 64-byte lines of the surface (whose rows are w*4 bytes, w a multiple of 16) belong either
   to the copy engine or to the threads that zero the margin, never to both.  false: the camera cannot see the tree. */
static bool visible_rect (const voxgpu_tree *t, const voxgpu_camera *c, int w, int h, int *x0, int *x1, int *y0, int *y1)
{
    int cull[4] = { 0, w, 0, h };
    voxgpu_screen_bounds (t->root_min, t->root_max, c, w, h, cull);
    *x0 = cull[0] & ~15; *x1 = (cull[1] + 15) & ~15; *y0 = cull[2] & ~3; *y1 = (cull[3] + 3) & ~3;
    if (*x1 > w) *x1 = w;
    if (*y1 > h) *y1 = h;
    return cull[1] > cull[0] && cull[3] > cull[2];
}

/* THE VISIBLE SPANS.  The rectangle above bounds the whole tree with one box; a voxel also lies in the box of one
   node of any cut through the tree, and a ray that hits it crosses that box, so its pixel lies inside the bounding
   rectangle of that box's projected corners (the argument of voxgpu_screen_bounds, per node).  The union of those
   rectangles, taken per group of `group` rows of the screen, bounds the columns each group can show: lo[u] <= x <
   hi[u] (hi <= lo: nothing).  The cut is the tree's coarse host cut (tree.inl, at most VOXGPU_HOST_CUT_MAX boxes, six
   arrays), projected here in fp32 -- with every corner at a depth of at least 1 % of its distance the error is
   below 0.1 pixel, and the rectangles are widened by 2 like the root's.  false: some box is beside or behind the
   camera (or the camera is not one of ours): no claim, the caller keeps the rectangle.
   On the C2 orbit the rectangle holds 55.5 % of the frame, the spans 41 %; cut into 8 strips of equal height
   (one copy each) 45.7 %. */
/* (fminf / fmaxf are calls into libm without -ffinite-math-only; these are one instruction) */
static inline float span_min (float a, float b) { return a < b ? a : b; }
static inline float span_max (float a, float b) { return a > b ? a : b; }
static bool visible_spans_of (const float *hbox, unsigned n, const voxgpu_camera *c, int w, int h, int group, int *lo, int *hi)
{
    if (!hbox || group < 1 || n < 1 || n > VOXGPU_HOST_CUT_MAX || !(c->mul > 0.0f)) return false;
    /* v = R p: the camera's rotation as a matrix, from the images of the unit vectors (the arithmetic of voxgpu_screen_bounds) */
    float R[3][3];
    for (int a=0; a<3; a++)
    {
        double p[3] = { 0, 0, 0 }, v[3];
        p[a] = 1;
        if (c->kind == VOXGPU_CAMERA_SIMPLE)
        {
            const double qw = c->rotation[0], u[3] = { -(double) c->rotation[1], -(double) c->rotation[2], -(double) c->rotation[3] };
            const double n2 = qw*qw + u[0]*u[0] + u[1]*u[1] + u[2]*u[2];
            if (!(n2 > 0.9999 && n2 < 1.0001)) return false;
            const double tt[3] = { 2*(u[1]*p[2] - u[2]*p[1]), 2*(u[2]*p[0] - u[0]*p[2]), 2*(u[0]*p[1] - u[1]*p[0]) };
            v[0] = p[0] + qw*tt[0] + (u[1]*tt[2] - u[2]*tt[1]);
            v[1] = p[1] + qw*tt[1] + (u[2]*tt[0] - u[0]*tt[2]);
            v[2] = p[2] + qw*tt[2] + (u[0]*tt[1] - u[1]*tt[0]);
        }
        else if (c->kind == VOXGPU_CAMERA_DOOM)
        {
            const double cs = c->cosphi, sn = c->sinphi, n2 = cs*cs + sn*sn;
            if (!(n2 > 0.9999 && n2 < 1.0001)) return false;
            v[0] = p[0]*cs + p[1]*sn; v[1] = p[1]*cs - p[0]*sn; v[2] = p[2];
        }
        else return false;
        for (int r=0; r<3; r++) R[r][a] = (float) v[r];
    }
    const float cx = c->position[0], cy = c->position[1], cz = c->position[2];
    const float kz = c->kind == VOXGPU_CAMERA_DOOM ? c->k : 0.0f, xsub = c->xsub, ysub = c->ysub, inv_mul = 1.0f / c->mul;
    const float *bx0 = hbox, *by0 = bx0 + n, *bz0 = by0 + n, *bx1 = bz0 + n, *by1 = bx1 + n, *bz1 = by1 + n;
    float rx0[VOXGPU_HOST_CUT_MAX], rx1[VOXGPU_HOST_CUT_MAX], ry0[VOXGPU_HOST_CUT_MAX], ry1[VOXGPU_HOST_CUT_MAX];
    int bad = 0;
    const float r00 = R[0][0], r01 = R[0][1], r02 = R[0][2], r10 = R[1][0], r11 = R[1][1], r12 = R[1][2],
                r20 = R[2][0], r21 = R[2][1], r22 = R[2][2];
    const float ycon = ysub + kz;
    #pragma omp simd reduction(|:bad)
    for (unsigned i=0; i<n; i++)
    {
        const float dx0 = bx0[i] - cx, dx1 = bx1[i] - cx, dy0 = by0[i] - cy, dy1 = by1[i] - cy, dz0 = bz0[i] - cz, dz1 = bz1[i] - cz;
        const float scale = span_max (fabsf (dx0), fabsf (dx1)) + span_max (fabsf (dy0), fabsf (dy1)) + span_max (fabsf (dz0), fabsf (dz1));
        /* the three rows of R p, one term per coordinate and end of the box */
        const float xa0 = r00*dx0, xa1 = r00*dx1, ya0 = r01*dy0, ya1 = r01*dy1, za0 = r02*dz0, za1 = r02*dz1;
        const float xb0 = r10*dx0, xb1 = r10*dx1, yb0 = r11*dy0, yb1 = r11*dy1, zb0 = r12*dz0, zb1 = r12*dz1;
        const float xc0 = r20*dx0, xc1 = r20*dx1, yc0 = r21*dy0, yc1 = r21*dy1, zc0 = r22*dz0, zc1 = r22*dz1;
        float lox = 3.0e38f, hix = -3.0e38f, loy = 3.0e38f, hiy = -3.0e38f, near = 3.0e38f;
#define SPAN_CORNER(X, Y, Z) \
        { \
            const float vy = (xb##X + yb##Y) + zb##Z, inv = 1.0f / vy; \
            const float sx = (((xa##X + ya##Y) + za##Z) * inv + xsub) * inv_mul, sy = (ycon - ((xc##X + yc##Y) + zc##Z) * inv) * inv_mul; \
            near = span_min (near, vy); \
            lox = span_min (lox, sx); hix = span_max (hix, sx); loy = span_min (loy, sy); hiy = span_max (hiy, sy); \
        }
        SPAN_CORNER (0, 0, 0) SPAN_CORNER (1, 0, 0) SPAN_CORNER (0, 1, 0) SPAN_CORNER (1, 1, 0)
        SPAN_CORNER (0, 0, 1) SPAN_CORNER (1, 0, 1) SPAN_CORNER (0, 1, 1) SPAN_CORNER (1, 1, 1)
#undef SPAN_CORNER
        bad |= !(near > 1.0e-2f * scale) || !(near > 0.0f);
        /* clamped far outside any screen: the conversions to int below stay defined */
        rx0[i] = span_min (span_max (lox, -1.0e6f), 1.0e6f); rx1[i] = span_min (span_max (hix, -1.0e6f), 1.0e6f);
        ry0[i] = span_min (span_max (loy, -1.0e6f), 1.0e6f); ry1[i] = span_min (span_max (hiy, -1.0e6f), 1.0e6f);
    }
    if (bad) return false;
    const int units = (h + group - 1) / group;
    for (int u=0; u<units; u++) { lo[u] = w; hi[u] = 0; }
    for (unsigned i=0; i<n; i++)
    {
        /* floor - 2 and ceil + 3 without libm: (int) truncates, the corrections make it a floor / a ceiling */
        int x0 = (int) rx0[i], x1 = (int) rx1[i], y0 = (int) ry0[i], y1 = (int) ry1[i];
        x0 -= 2 + ((float) x0 > rx0[i]); y0 -= 2 + ((float) y0 > ry0[i]);
        x1 += 3 + ((float) x1 < rx1[i]); y1 += 3 + ((float) y1 < ry1[i]);
        if (x1 <= 0 || y1 <= 0 || x0 >= w || y0 >= h) continue;
        const int a = x0 < 0 ? 0 : x0, b = x1 > w ? w : x1;
        const int u0 = (y0 < 0 ? 0 : y0) / group, u1 = ((y1 > h ? h : y1) + group - 1) / group;
        for (int u=u0; u<u1; u++) { lo[u] = a < lo[u] ? a : lo[u]; hi[u] = b > hi[u] ? b : hi[u]; }
    }
    return true;
}

static bool visible_spans (const voxgpu_tree *t, const voxgpu_camera *c, int w, int h, int group, int *lo, int *hi)
{ return t->n_hbox > 1 && visible_spans_of (t->hbox, t->n_hbox, c, w, h, group, lo, hi); }
/* the arithmetic alone, for the tests (host only: no device is touched) */
extern "C" int voxgpu_visible_spans (const float *boxes, unsigned n, const voxgpu_camera *c, int w, int h, int *lo, int *hi)
{ return boxes && c && lo && hi && w > 0 && h >= 4 && visible_spans_of (boxes, n, c, w, h, 4, lo, hi) ? 1 : 0; }

/* Zero every pixel of a w x h surface that the frame's copies do not write: row y keeps the columns
   [lo[y/4], hi[y/4]) (hi <= lo: nothing kept; lo == NULL: nothing kept anywhere).  A memset of several MB is
   bound by one core's store bandwidth, so a few threads share the rows (the num_threads clause, not
   $OMP_NUM_THREADS, which launchers like torchrun set to 1). */
static void host_clear_margin (void *pixels, size_t pitch, int w, int h, const int *lo, const int *hi, bool all_devices = false)
{
    /* up to 4 threads, but leave the box to its other GPUs' processes: half of the cores, shared by the visible
       devices -- unless this process drives all of them itself (voxgpu_multi_render): then up to 8 */
    static const int forced = getenv ("VOXGPU_CLEAR_THREADS") ? atoi (getenv ("VOXGPU_CLEAR_THREADS")) : 0;
    static const int per_device = [] {
        int gpus = 1;
        if (cudaGetDeviceCount (&gpus) != cudaSuccess || gpus < 1) { cudaGetLastError (); gpus = 1; }
        const int n = omp_get_num_procs () / (2 * gpus);
        return n < 1 ? 1 : (n > 4 ? 4 : n);
    } ();
    static const int whole_box = [] { const int n = omp_get_num_procs () / 2; return n < 1 ? 1 : (n > 8 ? 8 : n); } ();
    const int threads = forced > 0 ? forced : (all_devices ? whole_box : per_device);
    (void) threads;             /* (used by the pragma below only, which nvcc's front end does not read) */
    char *base = (char*) pixels;
    const size_t row = (size_t) w * 4;
    /* (chunks of 16 rows dealt round: whole rows above and below the spans, parts of rows beside them) */
    #pragma omp parallel for schedule(static, 16) num_threads(threads > 0 ? threads : 1) if (threads > 1)
    for (int y = 0; y < h; y++)
    {
        char *r = base + (size_t) y * pitch;
        const int x0 = lo ? lo[y >> 2] : 0, x1 = lo ? hi[y >> 2] : 0;
        if (x1 <= x0) memset (r, 0, row);
        else
        {
            if (x0 > 0) memset (r, 0, (size_t) x0 * 4);
            if (x1 < w) memset (r + (size_t) x1 * 4, 0, (size_t) (w - x1) * 4);
        }
    }
}

/* Knobs of the visible spans (voxgpu_set_present_spans; the defaults may come from the environment). */
static struct SpanCfg
{
    std::atomic<int> on, strips;
    std::atomic<long long> cost;
    SpanCfg ()
    {
        on = !(getenv ("VOXGPU_SPANS") && !atoi (getenv ("VOXGPU_SPANS")));
        strips = getenv ("VOXGPU_COPY_STRIPS") ? atoi (getenv ("VOXGPU_COPY_STRIPS")) : 16;
        cost = getenv ("VOXGPU_COPY_COST") ? atoll (getenv ("VOXGPU_COPY_COST")) : 98304;
    }
} g_span_cfg;
extern "C" void voxgpu_set_present_spans (int on, int max_strips, long long copy_cost)
{
    if (on >= 0) g_span_cfg.on = on != 0;
    if (max_strips > 0) g_span_cfg.strips = max_strips;
    if (copy_cost >= 0) g_span_cfg.cost = copy_cost;
}

/* The copies of rows [4*b0, 4*b1) cut into `sub` strips of equal height: every strip between the extreme columns
   of its groups (whole 64-byte lines, inside [vx0, vx1)), neighbours with the same columns joined, empty ones
   dropped.  -> number of runs, their bytes. */
#define PRESENT_MAX_STRIPS 32
struct StripRun { int u0, u1, x0, x1; };
static int strip_runs (int b0, int b1, int sub, const int *glo, const int *ghi, int G, int vx0, int vx1, StripRun *runs, long long *bytes)
{
    int n = 0;
    *bytes = 0;
    for (int j=0; j<sub; j++)
    {
        const int s0 = b0 + (int) ((long long) (b1 - b0) * j / sub), s1 = b0 + (int) ((long long) (b1 - b0) * (j + 1) / sub);
        int x0 = vx1, x1 = vx0;
        for (int q=s0*4/G; q<(s1*4+G-1)/G; q++) { if (glo[q] < x0) x0 = glo[q]; if (ghi[q] > x1) x1 = ghi[q]; }
        x0 &= ~15; x1 = (x1 + 15) & ~15;
        if (x0 < vx0) x0 = vx0;
        if (x1 > vx1) x1 = vx1;
        if (x1 <= x0 || s1 <= s0) continue;
        *bytes += (long long) (x1 - x0) * 16 * (s1 - s0);
        if (n && runs[n-1].u1 == s0 && runs[n-1].x0 == x0 && runs[n-1].x1 == x1) runs[n-1].u1 = s1;
        else { runs[n].u0 = s0; runs[n].u1 = s1; runs[n].x0 = x0; runs[n].x1 = x1; n++; }
    }
    return n;
}

/* The copies of one band, strips of equal height: 1, 2, 4, ... (at most max_strips) of them, whichever costs least
   when one more copy is worth copy_cost bytes.  -> number of runs in best[]. */
static int band_copies_equal (int b0, int b1, int max_strips, long long copy_cost, const int *glo, const int *ghi, int G,
                              int vx0, int vx1, StripRun *best)
{
    StripRun runs[PRESENT_MAX_STRIPS];
    int nbest = 0;
    long long best_cost = 0;
    for (int sub=1; sub<=max_strips && sub<=b1-b0 && sub<=PRESENT_MAX_STRIPS; sub*=2)
    {
        long long bytes = 0;
        const int nr = strip_runs (b0, b1, sub, glo, ghi, G, vx0, vx1, runs, &bytes);
        const long long cost = bytes + (long long) nr * copy_cost;
        if (sub == 1 || cost < best_cost) { best_cost = cost; nbest = nr; memcpy (best, runs, sizeof (StripRun) * nr); }
    }
    return nbest;
}

/* The copies of one band (units [b0, b1)): the cut into strips along group boundaries that moves the fewest bytes
   when every copy is worth copy_cost bytes more -- a shortest path over at most 64 cut points (f[j] = the cheapest
   way to cover the units before cut j), ~2 000 steps.  On the C2 orbit 44.4 % of the frames' bytes against 46.6 %
   for strips of equal height with the same ~3 copies per frame.  More than max_strips strips (a copy cost near
   zero): the equal strips instead.  -> number of runs in best[]. */
#define PRESENT_MAX_CUTS 64
static int band_copies (int b0, int b1, int max_strips, long long copy_cost, const int *glo, const int *ghi, int G,
                        int vx0, int vx1, StripRun *best)
{
    static const int dp = !(getenv ("VOXGPU_STRIP_DP") && !atoi (getenv ("VOXGPU_STRIP_DP")));     /* 0: equal strips (A/B) */
    if (!dp) return band_copies_equal (b0, b1, max_strips, copy_cost, glo, ghi, G, vx0, vx1, best);
    if (max_strips <= 1 || b1 - b0 < 2) return band_copies_equal (b0, b1, 1, copy_cost, glo, ghi, G, vx0, vx1, best);
    const int gu = G / 4;                                   /* units per group */
    const int q0 = b0 / gu, q1 = (b1 + gu - 1) / gu;        /* groups the band touches */
    const int step = (q1 - q0 + PRESENT_MAX_CUTS - 1) / PRESENT_MAX_CUTS;
    int cut[PRESENT_MAX_CUTS + 2], lo[PRESENT_MAX_CUTS + 1], hi[PRESENT_MAX_CUTS + 1], m = 0;
    for (int q=q0; q<q1; q+=step)
    {
        /* segment m: the groups [q, q + step) inside the band */
        const int qe = q + step < q1 ? q + step : q1;
        cut[m] = q * gu > b0 ? q * gu : b0;
        int a = vx1, b = vx0;
        for (int r=q; r<qe; r++) { if (glo[r] < a) a = glo[r]; if (ghi[r] > b) b = ghi[r]; }
        a &= ~15; b = (b + 15) & ~15;
        lo[m] = a > vx0 ? a : vx0; hi[m] = b < vx1 ? b : vx1;
        if (hi[m] <= lo[m]) { lo[m] = vx1; hi[m] = vx0; }   /* nothing: neutral for the min / max below */
        m++;
    }
    cut[m] = b1;
    long long f[PRESENT_MAX_CUTS + 1];
    int from[PRESENT_MAX_CUTS + 1];
    f[0] = 0; from[0] = 0;
    for (int j=1; j<=m; j++)
    {
        long long bestc = -1;
        int a = vx1, b = vx0, bi = j - 1;
        for (int i=j-1; i>=0; i--)
        {
            if (lo[i] < a) a = lo[i];
            if (hi[i] > b) b = hi[i];
            const long long bytes = b > a ? (long long) (b - a) * 16 * (cut[j] - cut[i]) : 0;
            const long long c = f[i] + bytes + (bytes ? copy_cost : 0);
            if (bestc < 0 || c < bestc) { bestc = c; bi = i; }
        }
        f[j] = bestc; from[j] = bi;
    }
    /* walk back, then emit top to bottom with equal neighbours joined */
    int ends[PRESENT_MAX_CUTS + 1], ne = 0;
    for (int j=m; j>0; j=from[j]) ends[ne++] = j;
    int n = 0, i = 0;
    for (int e=ne-1; e>=0; e--)
    {
        const int j = ends[e];
        int a = vx1, b = vx0;
        for (int r=i; r<j; r++) { if (lo[r] < a) a = lo[r]; if (hi[r] > b) b = hi[r]; }
        if (b > a)
        {
            if (n && best[n-1].u1 == cut[i] && best[n-1].x0 == a && best[n-1].x1 == b) best[n-1].u1 = cut[j];
            else
            {
                if (n >= max_strips || n >= PRESENT_MAX_STRIPS) return band_copies_equal (b0, b1, max_strips, copy_cost, glo, ghi, G, vx0, vx1, best);
                best[n].u0 = cut[i]; best[n].u1 = cut[j]; best[n].x0 = a; best[n].x1 = b; n++;
            }
        }
        i = j;
    }
    return n;
}

/* What voxgpu_present_render would copy for a frame of a tree whose coarse cut is `boxes` (as for
   voxgpu_visible_spans), cut into nbands bands -- host arithmetic only, for the tests: per 4-row unit u the kept
   columns [keep_lo[u], keep_hi[u]) (everything else of the surface is zeroed on the host) -> bytes copied, or -1
   if the camera cannot see the boxes at all (the whole surface is zeroed). */
extern "C" long long voxgpu_present_plan (const float *boxes, unsigned n, const voxgpu_camera *c, int w, int h, int nbands,
                                          int *keep_lo, int *keep_hi)
{
    if (!boxes || !n || !c || !keep_lo || !keep_hi || w <= 0 || h < 4 || (w & 15) || (h & 3) || nbands < 1) return -2;
    float mn[3] = { boxes[0], boxes[n], boxes[2*(size_t)n] }, mx[3] = { boxes[3*(size_t)n], boxes[4*(size_t)n], boxes[5*(size_t)n] };
    for (unsigned i=1; i<n; i++)
        for (int a=0; a<3; a++)
        {
            if (boxes[(size_t) a * n + i] < mn[a]) mn[a] = boxes[(size_t) a * n + i];
            if (boxes[(size_t) (3 + a) * n + i] > mx[a]) mx[a] = boxes[(size_t) (3 + a) * n + i];
        }
    int cull[4] = { 0, w, 0, h };
    voxgpu_screen_bounds (mn, mx, c, w, h, cull);
    const int units = h / 4, G = 16;
    for (int u=0; u<units; u++) keep_lo[u] = keep_hi[u] = 0;
    if (!(cull[1] > cull[0] && cull[3] > cull[2])) return -1;
    int vx0 = cull[0] & ~15, vx1 = (cull[1] + 15) & ~15, vy0 = cull[2] & ~3, vy1 = (cull[3] + 3) & ~3;
    if (vx1 > w) vx1 = w;
    if (vy1 > h) vy1 = h;
    const int v_units = (vy1 - vy0) / 4, v_u0 = vy0 / 4;
    if (nbands > v_units) nbands = v_units > 0 ? v_units : 1;
    if (nbands > PRESENT_MAX_BANDS) nbands = PRESENT_MAX_BANDS;
    int *glo = (int*) malloc (sizeof (int) * 2 * (size_t) units), *ghi = glo + units;
    const bool spans = g_span_cfg.on.load () && n > 1 && visible_spans_of (boxes, n, c, w, h, G, glo, ghi);
    if (!spans) for (int q=0; q<(h+G-1)/G; q++) { glo[q] = vx0; ghi[q] = vx1; }
    const int frame_strips = g_span_cfg.strips.load ();
    const int band_strips = spans && frame_strips > nbands ? frame_strips / nbands : 1;
    long long bytes = 0;
    for (int i=0; i<nbands; i++)
    {
        const int b0 = v_u0 + (int) ((long long) v_units * i / nbands), b1 = v_u0 + (int) ((long long) v_units * (i + 1) / nbands);
        if (b1 <= b0) continue;
        StripRun best[PRESENT_MAX_STRIPS];
        const int nb = band_copies (b0, b1, band_strips, g_span_cfg.cost.load (), glo, ghi, G, vx0, vx1, best);
        for (int r=0; r<nb; r++)
        {
            bytes += (long long) (best[r].x1 - best[r].x0) * 16 * (best[r].u1 - best[r].u0);
            for (int u=best[r].u0; u<best[r].u1; u++) { keep_lo[u] = best[r].x0; keep_hi[u] = best[r].x1; }
        }
    }
    free (glo);
    return bytes;
}

extern "C" int voxgpu_present_render (voxgpu_present *p, const voxgpu_tree *t, const voxgpu_camera *c,
                                      const voxgpu_frame *f, const uint8_t *d_tex,
                                      void *host_pixels, size_t host_pitch_bytes)
{
    if (!p || !f || !c || !host_pixels) return fail ("voxgpu_present_render: NULL argument");
    if (f->w != p->w || f->h != p->h) return fail ("voxgpu_present_render: frame is %dx%d, presenter %dx%d", f->w, f->h, p->w, p->h);
    if (host_pitch_bytes < (size_t) p->w * 4) return fail ("voxgpu_present_render: pitch %zu too small", host_pitch_bytes);
    int cur = -1;
    CU (cudaGetDevice (&cur));
    if (cur != p->device) return fail ("voxgpu_present_render: presenter lives on device %d, current device is %d", p->device, cur);
    const int il_world = f->interleave_world > 1 ? f->interleave_world : 1;
    const int il_rank = il_world > 1 ? f->interleave_rank : 0;
    if (il_rank < 0 || il_rank >= il_world) return fail ("voxgpu_present_render: interleave rank %d of %d", il_rank, il_world);
    if (il_world > 1 && host_pitch_bytes != (size_t) p->w * 4)
        return fail ("voxgpu_present_render: a share of an interleaved frame needs pitch == 4*w");

    if (p->in_flight && voxgpu_present_wait (p)) return -1;   /* the previous frame owns fb and the events */
    const int others = g_presents_in_flight.load ();
    int nbands = il_world > 1 ? 1 : ((!p->fixed_bands && others > 0) ? 1 : p->bands);

    /* THE VISIBLE RECTANGLE.  Pixels outside the root box's screen bounds (voxgpu_screen_bounds: conservative,
       host arithmetic) are zero whatever the tree holds.  They are neither drawn nor sent over PCIe: the device
       renders and copies the tile-aligned rectangle [vx0,vx1) x [vy0,vy1) only, and the calling thread (with a few
       helpers) zeroes the margin of the surface while the GPU works -- the reference clears its squares on the
       host every frame too (renderer.c:324).  On a C2 frame the rectangle holds half of the pixels, and the PCIe
       link is what bounds vox_render + vox_redraw.  ($VOXGPU_HOST_CLEAR=0: the whole frame comes from the device.) */
    int vx0 = 0, vx1 = p->w, vy0 = 0, vy1 = p->h;
    static const int host_clear = !(getenv ("VOXGPU_HOST_CLEAR") && !atoi (getenv ("VOXGPU_HOST_CLEAR")));
    bool nothing = false;
    if (host_clear && t && g_cull)
    {
        nothing = !visible_rect (t, c, p->w, p->h, &vx0, &vx1, &vy0, &vy1);
        /* (one GPU's share of a split frame: its tile rows of the rectangle; the caller of all shares,
           voxgpu_multi_render, zeroes the margin once) */
    }
    const int v_units = nothing ? 0 : (vy1 - vy0) / 4, v_u0 = vy0 / 4;
    if (nbands > v_units) nbands = v_units > 0 ? v_units : 1;
    CU (cudaEventRecord (p->start, p->stream[0]));
    if (nothing)
    {
        /* the camera cannot see the tree at all: the frame is the margin */
        if (il_world == 1) host_clear_margin (host_pixels, host_pitch_bytes, p->w, p->h, NULL, NULL);
        p->used_bands = 0;
        return 0;
    }
    /* Inside the rectangle, the columns every group of 16 rows can show (visible_spans; $VOXGPU_SPANS=0: the
       rectangle's), and, filled in as the copies are queued, the columns every 4-row unit's copy writes: the rest
       is the margin. */
    const int all_units = p->h / 4;
    int *ulo = p->span, *uhi = ulo + all_units, *clo = uhi + all_units, *chi = clo + all_units;
    const int use_spans = g_span_cfg.on.load (), frame_strips = g_span_cfg.strips.load ();
    const long long copy_cost = g_span_cfg.cost.load ();
    const bool clearing = host_clear && t && g_cull && il_world == 1;
    const int G = 16;                   /* rows per group of the spans: a strip is many groups high */
    /* one screen mask per frame, marked on the first band's stream, shared by all bands */
    bool masked = false;
    unsigned char *cost = NULL;
    if (t && mask_usable (t, c))
    {
        const size_t mtiles = (size_t) (p->w >> 3) * (p->h >> 2);
        if (maskbuf_begin (p->mask, mtiles, p->stream[0])) return -1;
        MarkBands mb;
        mb.y0 = vy0; mb.y1 = vy1; mb.nbands = nbands; mb.il_world = il_world; mb.il_rank = il_rank;
        {
            /* (the cost map belongs to the tree: made under the device's render lock) */
            std::lock_guard<std::mutex> guard (g_rdev[t->device].lock);
            cost = (f->quality & 3u) == 0u && !f->start_node ? cost_map (t, p->w, p->h, p->stream[0]) : NULL;
        }
        if (mark_launch (t, c, p->w, p->h, p->mask, mb, p->stream[0], cost)) return -1;
        masked = true;
    }
    CU (cudaEventRecord (p->marked, p->stream[0]));
    int rc = 0, queued = 0, launched = 0;
    /* Band k of the rectangle is queued on the stream of its RANK.  When the whole frame comes from the device
       (no host clearing) the bands with the fewest rows inside the root box's screen bounds go first (highest
       priority): they are cleared rather than traversed, done at once, and cross PCIe while the others render, so
       that what is left exposed at the end is the copy of one band and not the kernel of a nearly empty one
       behind it ($VOXGPU_BAND_ORDER=0: top to bottom).  Inside the visible rectangle all rows count the same. */
    int order[PRESENT_MAX_BANDS];
    for (int i=0; i<nbands; i++) order[i] = i;
    static const int by_weight = !(getenv ("VOXGPU_BAND_ORDER") && !atoi (getenv ("VOXGPU_BAND_ORDER")));
    if (by_weight && nbands > 1 && t && g_cull)
    {
        int cull[4] = { 0, p->w, 0, p->h }, weight[PRESENT_MAX_BANDS];
        voxgpu_screen_bounds (t->root_min, t->root_max, c, p->w, p->h, cull);
        for (int i=0; i<nbands; i++)
        {
            const int y0 = (v_u0 + (int) ((long long) v_units * i / nbands)) * 4, y1 = (v_u0 + (int) ((long long) v_units * (i + 1) / nbands)) * 4;
            const int lo = y0 > cull[2] ? y0 : cull[2], hi = y1 < cull[3] ? y1 : cull[3];
            weight[i] = hi > lo ? hi - lo : 0;
        }
        for (int i=1; i<nbands; i++)            /* insertion sort, stable: equal weights stay top to bottom */
            for (int j=i; j>0 && weight[order[j]] < weight[order[j-1]]; j--) { const int x = order[j]; order[j] = order[j-1]; order[j-1] = x; }
    }
    for (int k=0; k<nbands && !rc; k++)
    {
        const int i = order[k];                 /* band i of the frame, on the stream / events of rank k */
        /* the later bands wait for the mask (and with it the start mark: the elapsed time covers all of them) */
        if (k && cudaStreamWaitEvent (p->stream[k], p->marked, 0) != cudaSuccess) { rc = fail ("voxgpu_present_render: cudaStreamWaitEvent failed"); break; }
        const int y0 = (v_u0 + (int) ((long long) v_units * i / nbands)) * 4;
        const int y1 = (v_u0 + (int) ((long long) v_units * (i + 1) / nbands)) * 4;
        launched++;
        if (y1 <= y0) continue;
        voxgpu_frame band = *f;
        band.y0 = y0; band.y1 = y1; band.pitch_pixels = p->w;
        DevMask mk;
        if (masked) mk = maskbuf_band (p->mask, i, (size_t) (y0 >> 2) * (p->w >> 3), cost);
        if (render_impl (t, c, &band, d_tex, p->fb + (size_t) y0 * p->w, NULL, NULL, (voxgpu_stream) p->stream[k], masked ? &mk : NULL))
        { rc = -1; launched--; break; }
    }
    /* The kernels are on their way; the host arithmetic of the spans runs while they do, then the copies queue up
       behind them, band by band. */
    const bool spans = clearing && use_spans && launched > 0 && visible_spans (t, c, p->w, p->h, G, ulo, uhi);
    if (!spans) for (int q=0; q<(p->h+G-1)/G; q++) { ulo[q] = vx0; uhi[q] = vx1; }
    if (clearing) for (int u=0; u<all_units; u++) clo[u] = chi[u] = 0;
    const int band_strips = spans && frame_strips > nbands ? frame_strips / nbands : 1;
    for (int k=0; k<launched; k++)
    {
        const int i = order[k];
        const int y0 = (v_u0 + (int) ((long long) v_units * i / nbands)) * 4;
        const int y1 = (v_u0 + (int) ((long long) v_units * (i + 1) / nbands)) * 4;
        if (y1 <= y0) { cudaEventRecord (p->done[k], p->stream[k]); queued++; continue; }
        const uint32_t *src = p->fb + (size_t) y0 * p->w;
        char *dst = (char*) host_pixels + (size_t) y0 * host_pitch_bytes;
        cudaError_t e;
        if (il_world > 1)
        {
            /* this GPU's share: tile rows rank, rank + world, ... of the band.  Four rows are contiguous on both sides */
            const int b_units = (y1 - y0) / 4;
            const int own = il_rank < b_units ? (b_units - il_rank + il_world - 1) / il_world : 0;
            const size_t row4 = (size_t) p->w * 16;
            if (!own) e = cudaSuccess;
            else if (vx0 == 0 && vx1 == p->w)
                e = cudaMemcpy2DAsync (dst + (size_t) il_rank * row4, row4 * il_world, (const char*) src + (size_t) il_rank * row4,
                                       row4 * il_world, row4, own, cudaMemcpyDeviceToHost, p->stream[k]);
            else
            {
                /* the columns of the rectangle only: a 3-D copy, one slice (4 rows of a 4*N-row slab) per tile row */
                cudaMemcpy3DParms q;
                memset (&q, 0, sizeof q);
                q.srcPtr = make_cudaPitchedPtr ((void*) ((const char*) src + (size_t) il_rank * row4), (size_t) p->w * 4, (size_t) p->w * 4, (size_t) 4 * il_world);
                q.dstPtr = make_cudaPitchedPtr ((void*) (dst + (size_t) il_rank * row4), (size_t) p->w * 4, (size_t) p->w * 4, (size_t) 4 * il_world);
                q.srcPos = make_cudaPos ((size_t) vx0 * 4, 0, 0);
                q.dstPos = q.srcPos;
                q.extent = make_cudaExtent ((size_t) (vx1 - vx0) * 4, 4, (size_t) own);
                q.kind = cudaMemcpyDeviceToHost;
                e = cudaMemcpy3DAsync (&q, p->stream[k]);
            }
        }
        else
        {
            /* the band's rows in strips, each copied between the extreme columns of its groups: as many strips
               (1, 2, 4, ... of equal height) as pay for themselves -- a copy more costs what ~100 KB on the wire cost */
            e = cudaSuccess;
            const int b0 = y0 >> 2, b1 = y1 >> 2;
            StripRun best[PRESENT_MAX_STRIPS];
            const int nbest = band_copies (b0, b1, band_strips, copy_cost, ulo, uhi, G, vx0, vx1, best);
            for (int r=0; r<nbest && e == cudaSuccess; r++)
            {
                const int ry0 = best[r].u0 * 4, rows = (best[r].u1 - best[r].u0) * 4, rx0 = best[r].x0, rx1 = best[r].x1;
                char *rd = (char*) host_pixels + (size_t) ry0 * host_pitch_bytes + (size_t) rx0 * 4;
                const uint32_t *rs = p->fb + (size_t) ry0 * p->w + rx0;
                if (host_pitch_bytes == (size_t) p->w * 4 && rx0 == 0 && rx1 == p->w)
                    e = cudaMemcpyAsync (rd, rs, (size_t) rows * p->w * 4, cudaMemcpyDeviceToHost, p->stream[k]);
                else
                    e = cudaMemcpy2DAsync (rd, host_pitch_bytes, rs, (size_t) p->w * 4, (size_t) (rx1 - rx0) * 4, rows,
                                           cudaMemcpyDeviceToHost, p->stream[k]);
                if (e == cudaSuccess) g_present_bytes += (unsigned long long) (rx1 - rx0) * 4 * rows;
                if (clearing) for (int u=best[r].u0; u<best[r].u1; u++) { clo[u] = rx0; chi[u] = rx1; }
            }
        }
        if (e != cudaSuccess) { rc = fail ("voxgpu_present_render: copy to the surface failed: %s", cudaGetErrorString (e)); }
        else if (il_world > 1)
            g_present_bytes += (unsigned long long) (vx1 - vx0) * 16 * (il_rank < (y1 - y0) / 4 ? ((y1 - y0) / 4 - il_rank + il_world - 1) / il_world : 0);
        cudaEventRecord (p->done[k], p->stream[k]);
        queued++;
    }
    if (rc && masked && queued < nbands) maskbuf_reset_lists (p->mask, p->stream[0]);   /* lists nobody will consume */
    /* the margin, while the device works on the rectangle */
    if (clearing && (spans || vx0 > 0 || vx1 < p->w || vy0 > 0 || vy1 < p->h))
        host_clear_margin (host_pixels, host_pitch_bytes, p->w, p->h, clo, chi);
    p->used_bands = queued;
    if (queued)
    {
        p->in_flight = 1;
        g_presents_in_flight++;
    }
    return rc;
}

extern "C" int voxgpu_present_wait (voxgpu_present *p)
{
    if (!p || !p->in_flight) return 0;
    int rc = 0;
    for (int i=0; i<p->used_bands; i++)
    {
        cudaError_t e = cudaEventSynchronize (p->done[i]);
        if (e != cudaSuccess && !rc) rc = fail ("voxgpu_present_wait: %s", cudaGetErrorString (e));
    }
    p->in_flight = 0;
    g_presents_in_flight--;
    return rc;
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

/* ------------------------------------------- one frame on all GPUs of the process
 *
 * The reference fans a frame out over all cores inside vox_render (dispatch_apply over the 4x4 blocks,
 * renderer.c:214-226).  voxgpu_multi does the same over the CUDA devices of ONE process: device k draws
 * the tile rows r with r % N == k of the frame (statistically identical shares, one launch each) into its
 * own framebuffer and DMA-copies exactly those rows into the page-locked surface over ITS OWN PCIe link --
 * no funnel through GPU 0, no peer traffic at all; the "gather" is the host surface itself.  Each device
 * has a worker thread (bound to the device once, so no cudaSetDevice on the frame path) that queues the
 * device's share when voxgpu_multi_render posts a frame; the N submissions run in parallel and
 * voxgpu_multi_render returns when all are queued.  voxgpu_multi_wait = every share has landed.
 */
struct MultiJob
{
    const voxgpu_tree *tree;
    const uint8_t *d_tex;
    const voxgpu_lights *lights;
    voxgpu_camera cam;
    voxgpu_frame frame;
    void *host_pixels;
    size_t pitch;
};

struct MultiWorker
{
    std::thread th;
    int device, rank;
    voxgpu_present *p;
    std::mutex m;
    std::condition_variable cv;
    std::atomic<unsigned> posted, finished;     /* sequence numbers of frames posted to / queued by this worker */
    std::atomic<int> quit;
    MultiJob job;
    int rc;
    char err[512];
};

struct voxgpu_multi
{
    int n, w, h;
    MultiWorker *wk;
    int in_flight;
    int *span;                      /* 2 x h/4 ints: the columns of every 4-row unit the devices' copies write */
};

static void multi_worker_main (voxgpu_multi *mu, MultiWorker *w)
{
    w->rc = 0; w->err[0] = 0;
    if (cudaSetDevice (w->device) != cudaSuccess) { w->rc = -1; snprintf (w->err, sizeof w->err, "cudaSetDevice (%d) failed", w->device); }
    else
    {
        w->p = voxgpu_present_create (mu->w, mu->h, 1);
        if (!w->p) { w->rc = -1; snprintf (w->err, sizeof w->err, "%s", g_err); }
    }
    w->finished.store (1, std::memory_order_release);            /* "frame 1" = the setup */
    unsigned seen = 1;
    for (;;)
    {
        /* a short spin (the next frame usually follows within microseconds), then sleep */
        int spins = 0;
        while (w->posted.load (std::memory_order_acquire) == seen && !w->quit.load (std::memory_order_acquire))
        {
            if (++spins < 4000) { __builtin_ia32_pause (); continue; }
            std::unique_lock<std::mutex> lk (w->m);
            w->cv.wait (lk, [&] { return w->posted.load (std::memory_order_acquire) != seen || w->quit.load (std::memory_order_acquire); });
        }
        if (w->quit.load (std::memory_order_acquire)) break;
        seen = w->posted.load (std::memory_order_acquire);
        MultiJob &j = w->job;
        voxgpu_frame fr = j.frame;
        fr.lights = j.lights;
        fr.interleave_world = mu->n; fr.interleave_rank = w->rank;
        w->rc = w->p ? voxgpu_present_render (w->p, j.tree, &j.cam, &fr, j.d_tex, j.host_pixels, j.pitch) : -1;
        if (w->rc) snprintf (w->err, sizeof w->err, "%s", w->p ? g_err : "worker has no presenter");
        w->finished.store (seen, std::memory_order_release);
    }
    if (w->p) { voxgpu_present_destroy (w->p); w->p = NULL; }
}

static void multi_post (MultiWorker *w, unsigned seq)
{
    {
        std::lock_guard<std::mutex> lk (w->m);
        w->posted.store (seq, std::memory_order_release);
    }
    w->cv.notify_one ();
}

static void multi_await (MultiWorker *w, unsigned seq)
{
    int spins = 0;
    while (w->finished.load (std::memory_order_acquire) != seq)
    {
        if (++spins < 20000) __builtin_ia32_pause ();
        else std::this_thread::yield ();
    }
}

extern "C" void voxgpu_multi_destroy (voxgpu_multi *mu)
{
    if (!mu) return;
    for (int i=0; i<mu->n; i++)
    {
        MultiWorker *w = &mu->wk[i];
        if (!w->th.joinable ()) continue;
        {
            std::lock_guard<std::mutex> lk (w->m);
            w->quit.store (1, std::memory_order_release);
        }
        w->cv.notify_one ();
        w->th.join ();
    }
    delete[] mu->wk;
    free (mu->span);
    free (mu);
}

extern "C" voxgpu_multi* voxgpu_multi_create (int w, int h, int ndev, const int *devices)
{
    if (w <= 0 || h <= 0 || (w & 0xf) || (h & 0x3)) { fail ("voxgpu_multi_create: bad geometry %dx%d", w, h); return NULL; }
    if (ndev < 1 || ndev > 64 || !devices) { fail ("voxgpu_multi_create: %d devices", ndev); return NULL; }
    voxgpu_multi *mu = (voxgpu_multi*) calloc (1, sizeof *mu);
    if (!mu) { fail ("voxgpu_multi_create: out of memory"); return NULL; }
    mu->n = ndev; mu->w = w; mu->h = h;
    mu->span = (int*) calloc ((size_t) (h / 4) * 2, sizeof (int));
    if (!mu->span) { free (mu); fail ("voxgpu_multi_create: out of memory"); return NULL; }
    mu->wk = new MultiWorker[ndev];
    for (int i=0; i<ndev; i++)
    {
        MultiWorker *k = &mu->wk[i];
        k->device = devices[i]; k->rank = i; k->p = NULL;
        k->posted.store (1); k->finished.store (0); k->quit.store (0);
        k->th = std::thread (multi_worker_main, mu, k);
    }
    int bad = -1;
    for (int i=0; i<ndev; i++)
    {
        multi_await (&mu->wk[i], 1);
        if (mu->wk[i].rc && bad < 0) bad = i;
    }
    if (bad >= 0)
    {
        fail ("voxgpu_multi_create: device %d: %s", mu->wk[bad].device, mu->wk[bad].err);
        voxgpu_multi_destroy (mu);
        return NULL;
    }
    return mu;
}

extern "C" int voxgpu_multi_devices (const voxgpu_multi *mu) { return mu ? mu->n : 0; }
extern "C" int voxgpu_multi_device (const voxgpu_multi *mu, int i) { return (mu && i >= 0 && i < mu->n) ? mu->wk[i].device : -1; }

extern "C" int voxgpu_multi_render (voxgpu_multi *mu, const voxgpu_tree *const *trees, const voxgpu_camera *c,
                                    const voxgpu_frame *f, const uint8_t *const *d_tex, const voxgpu_lights *const *lights,
                                    void *host_pixels, size_t host_pitch_bytes)
{
    if (!mu || !trees || !c || !f || !d_tex || !host_pixels) return fail ("voxgpu_multi_render: NULL argument");
    if ((f->quality & 3u) != 0u) return fail ("voxgpu_multi_render: a frame split over devices needs quality BEST");
    if (mu->in_flight && voxgpu_multi_wait (mu)) return -1;
    const unsigned seq = mu->wk[0].posted.load () + 1;
    for (int i=0; i<mu->n; i++)
    {
        MultiJob &j = mu->wk[i].job;
        j.tree = trees[i]; j.d_tex = d_tex[i]; j.lights = lights ? lights[i] : NULL;
        j.cam = *c; j.frame = *f; j.host_pixels = host_pixels; j.pitch = host_pitch_bytes;
        multi_post (&mu->wk[i], seq);
    }
    /* what no device draws (outside the root box's screen bounds: see voxgpu_present_render), while the workers queue */
    {
        static const int host_clear = !(getenv ("VOXGPU_HOST_CLEAR") && !atoi (getenv ("VOXGPU_HOST_CLEAR")));
        if (host_clear && g_cull && trees[0])
        {
            int vx0, vx1, vy0, vy1;
            const bool nothing = !visible_rect (trees[0], c, mu->w, mu->h, &vx0, &vx1, &vy0, &vy1);
            if (nothing) host_clear_margin (host_pixels, host_pitch_bytes, mu->w, mu->h, NULL, NULL, true);
            else if (vx0 > 0 || vx1 < mu->w || vy0 > 0 || vy1 < mu->h)
            {
                const int units = mu->h / 4;
                int *lo = mu->span, *hi = lo + units;
                for (int u=0; u<units; u++) { const bool in = u >= vy0 / 4 && u < vy1 / 4; lo[u] = in ? vx0 : 0; hi[u] = in ? vx1 : 0; }
                host_clear_margin (host_pixels, host_pitch_bytes, mu->w, mu->h, lo, hi, true);
            }
        }
    }
    int rc = 0;
    for (int i=0; i<mu->n; i++)
    {
        multi_await (&mu->wk[i], seq);
        if (mu->wk[i].rc && !rc) rc = fail ("voxgpu_multi_render: device %d: %s", mu->wk[i].device, mu->wk[i].err);
    }
    mu->in_flight = 1;
    return rc;
}

extern "C" int voxgpu_multi_wait (voxgpu_multi *mu)
{
    if (!mu || !mu->in_flight) return 0;
    int rc = 0;
    /* cudaEventSynchronize does not care which device is current */
    for (int i=0; i<mu->n; i++) if (mu->wk[i].p && voxgpu_present_wait (mu->wk[i].p)) rc = -1;
    mu->in_flight = 0;
    return rc;
}

extern "C" float voxgpu_multi_last_ms (voxgpu_multi *mu)
{
    if (!mu) return -1.0f;
    float worst = -1.0f;
    for (int i=0; i<mu->n; i++)
    {
        const float ms = mu->wk[i].p ? voxgpu_present_last_ms (mu->wk[i].p) : -1.0f;
        if (ms > worst) worst = ms;
    }
    return worst;
}
