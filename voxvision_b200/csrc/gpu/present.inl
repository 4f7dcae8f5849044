/* present.inl -- part of the single translation unit voxgpu.cu (included there, in this order):
   frames to host memory: voxgpu_present_*. */
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
