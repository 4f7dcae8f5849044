/* sort.inl -- part of the single translation unit voxgpu.cu (included there, in this order):
   entry-point order for ray batches: k_ray_keys + a two-pass LSD radix sort of the ray indices. */
/* ------------------------------------------------- ray batches in entry order
 *
 * K2 gives a warp 32 consecutive rays of the caller's array.  When those are unrelated (C3: random origins
 * and directions) every lane fetches different node records from the second or third level on: 32 sectors
 * per load instruction, nothing shared in L1, and lanes that finish early idle until the slowest ray of the
 * warp is done -- r01 measured 0.18 of the HBM roofline there against 0.40 for the coherent rays of a frame.
 * The search of a ray is a pure function of the ray (bit-exactness is per ray), so the ORDER in which a batch
 * is searched is free: the batch is searched in the order of the Morton code of the point where each ray
 * enters the box of the start node (6 bits per axis; rays that miss the box sort last and end at once), and
 * results are scattered back to the caller's positions.  Neighbouring lanes then walk the same subtree.
 *
 * The sort is a plain LSD radix sort of 18-bit keys in two 9-bit passes, written for this one job: every warp
 * owns a contiguous run of SORT_WARP_KEYS keys; pass = per-warp digit histogram -> exclusive scan over
 * (digit, warp) -> stable scatter, the rank of a key among the equal digits of its 32-key step coming from
 * __match_any_sync.  Payload: the ray index.  16 Mi keys: 4 x 64 MB of scratch, ~1 ms.
 */
#define SORT_BITS 9
#define SORT_DIGITS (1 << SORT_BITS)
#define SORT_WARP_KEYS 4096u
#define SORT_BLOCK 256
#define SORT_KEY_BITS 18
#define SORT_KEY_MISS ((1u << SORT_KEY_BITS) - 1u)

__device__ __forceinline__ unsigned spread6 (unsigned v)
{
    /* 6 bits -> every third bit */
    v &= 0x3fu;
    v = (v | (v << 8)) & 0x300fu;
    v = (v | (v << 4)) & 0x30c3u;
    v = (v | (v << 2)) & 0x9249u;
    return v;
}

__global__ void __launch_bounds__(256)
k_ray_keys (TreeRef tr, unsigned start, const float4 * __restrict__ origins, const float4 * __restrict__ dirs, size_t n,
            unsigned * __restrict__ keys)
{
    const size_t i = (size_t) blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float4 q0, q1;
    load_rec (tr.rec + start, q0, q1);
    const V3 mn = v3 (q0), mx = v3 (q1);
    const V3 o = v3 (__ldg (origins + i));
    const RayK k = ray_constants (v3 (__ldg (dirs + i)));
    V3 e;
    unsigned key = SORT_KEY_MISS;
    if (ray_box_fast (mn, mx, o, k.d, k.r, k.dok, e))
    {
        /* a hint, not a result: plain arithmetic */
        const float sx = 64.0f / fmaxf (mx.x - mn.x, 1e-30f), sy = 64.0f / fmaxf (mx.y - mn.y, 1e-30f), sz = 64.0f / fmaxf (mx.z - mn.z, 1e-30f);
        const unsigned cx = (unsigned) min (max ((int) ((e.x - mn.x) * sx), 0), 63);
        const unsigned cy = (unsigned) min (max ((int) ((e.y - mn.y) * sy), 0), 63);
        const unsigned cz = (unsigned) min (max ((int) ((e.z - mn.z) * sz), 0), 63);
        key = spread6 (cx) | (spread6 (cy) << 1) | (spread6 (cz) << 2);
        if (key == SORT_KEY_MISS) key--;
    }
    keys[i] = key;
}

/* per-warp digit histogram: hist[digit * nwarps + warp] */
__global__ void __launch_bounds__(SORT_BLOCK)
k_sort_hist (const unsigned * __restrict__ keys, size_t n, unsigned shift, unsigned nwarps, unsigned * __restrict__ hist)
{
    __shared__ unsigned cnt[SORT_BLOCK/32][SORT_DIGITS];
    const unsigned wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned warp = blockIdx.x * (SORT_BLOCK/32) + wib;
    for (unsigned d = lane; d < SORT_DIGITS; d += 32) cnt[wib][d] = 0;
    __syncwarp ();
    if (warp < nwarps)
    {
        const size_t lo = (size_t) warp * SORT_WARP_KEYS;
        const size_t hi = lo + SORT_WARP_KEYS < n ? lo + SORT_WARP_KEYS : n;
        for (size_t i = lo + lane; i < hi; i += 32) atomicAdd (&cnt[wib][(keys[i] >> shift) & (SORT_DIGITS - 1)], 1u);
        __syncwarp ();
        for (unsigned d = lane; d < SORT_DIGITS; d += 32) hist[(size_t) d * nwarps + warp] = cnt[wib][d];
    }
}

/* exclusive scan of hist in (digit, warp) order, in place.  One CTA per digit scans its row and leaves the row's
   total in totals[digit]; k_sort_scatter adds the scanned totals of the lower digits itself. */
__global__ void __launch_bounds__(256)
k_sort_scan_rows (unsigned * __restrict__ hist, unsigned nwarps, unsigned * __restrict__ totals)
{
    __shared__ unsigned part[256];
    unsigned *row = hist + (size_t) blockIdx.x * nwarps;
    const unsigned per = (nwarps + 255) / 256;
    const unsigned lo = threadIdx.x * per, hi = min (lo + per, nwarps);
    unsigned s = 0;
    for (unsigned i = lo; i < hi; i++) s += row[i];
    part[threadIdx.x] = s;
    __syncthreads ();
    /* Hillis-Steele over 256 partial sums */
    for (unsigned off = 1; off < 256; off <<= 1)
    {
        const unsigned v = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
        __syncthreads ();
        part[threadIdx.x] += v;
        __syncthreads ();
    }
    unsigned run = part[threadIdx.x] - s;       /* exclusive prefix of this thread's stretch */
    for (unsigned i = lo; i < hi; i++) { const unsigned v = row[i]; row[i] = run; run += v; }
    if (threadIdx.x == 255) totals[blockIdx.x] = part[255];
}

/* stable scatter of (key, value) by the digit at `shift`; vals_in == NULL: the values are the positions 0..n-1;
   keys_out == NULL: the keys are not needed any more */
__global__ void __launch_bounds__(SORT_BLOCK)
k_sort_scatter (const unsigned * __restrict__ keys, const unsigned * __restrict__ vals_in, size_t n, unsigned shift, unsigned nwarps,
                const unsigned * __restrict__ hist, const unsigned * __restrict__ totals,
                unsigned * __restrict__ keys_out, unsigned * __restrict__ vals_out)
{
    __shared__ unsigned off[SORT_BLOCK/32][SORT_DIGITS];
    __shared__ unsigned base[SORT_DIGITS];
    const unsigned wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned warp = blockIdx.x * (SORT_BLOCK/32) + wib;
    /* exclusive scan of the digit totals (512 values, every CTA for itself) */
    {
        const unsigned a = totals[2*threadIdx.x], b = totals[2*threadIdx.x + 1];
        const unsigned pairsum = a + b;
        unsigned incl = pairsum;
        /* inclusive scan of pair sums over 256 threads: warp shuffles, then across the 8 warps */
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync (0xffffffffu, incl, o); if ((int) lane >= o) incl += v; }
        __shared__ unsigned wsum[SORT_BLOCK/32];
        if (lane == 31) wsum[wib] = incl;
        __syncthreads ();
        unsigned before = 0;
        for (unsigned w = 0; w < wib; w++) before += wsum[w];
        const unsigned excl = before + incl - pairsum;                      /* sum of all digits below 2*tid */
        base[2*threadIdx.x] = excl; base[2*threadIdx.x + 1] = excl + a;
        __syncthreads ();
    }
    if (warp >= nwarps) return;
    for (unsigned d = lane; d < SORT_DIGITS; d += 32) off[wib][d] = base[d] + hist[(size_t) d * nwarps + warp];
    __syncwarp ();
    const size_t lo = (size_t) warp * SORT_WARP_KEYS;
    const size_t hi = lo + SORT_WARP_KEYS < n ? lo + SORT_WARP_KEYS : n;
    for (size_t i0 = lo; i0 < hi; i0 += 32)
    {
        const size_t i = i0 + lane;
        const bool live = i < hi;
        const unsigned key = live ? keys[i] : 0u;
        const unsigned d = live ? (key >> shift) & (SORT_DIGITS - 1) : SORT_DIGITS;      /* dead lanes match only each other */
        const unsigned peers = __match_any_sync (0xffffffffu, d);
        const unsigned rank = __popc (peers & ((1u << lane) - 1u));
        unsigned pos = 0;
        if (live) pos = off[wib][d] + rank;
        __syncwarp ();
        if (live && rank == 0) off[wib][d] += __popc (peers);
        __syncwarp ();
        if (live)
        {
            if (keys_out) keys_out[pos] = key;
            vals_out[pos] = vals_in ? vals_in[i] : (unsigned) i;
        }
    }
}

/* Per-device scratch of the sort, reused from call to call; calls on different streams are ordered by an event. */
struct SortScratch
{
    std::mutex lock;
    unsigned *buf;              /* keys | keys1 | perm1 | perm2 | hist | totals */
    size_t cap;                 /* rays it has room for */
    cudaEvent_t ev;
    cudaStream_t last;
    int used;
};
static SortScratch g_sort[64];

/* Queue on `st`: *perm = device array with the n ray indices in entry order (valid until the next call on this
   device; the caller's kernels must be queued on `st` before releasing the scratch with sort_release). */
static int sort_rays (const voxgpu_tree *t, unsigned start, const float4 *d_o, const float4 *d_d, size_t n,
                      cudaStream_t st, const unsigned **perm)
{
    SortScratch &sc = g_sort[t->device];
    const unsigned nwarps = (unsigned) ((n + SORT_WARP_KEYS - 1) / SORT_WARP_KEYS);
    const size_t hist_words = (size_t) SORT_DIGITS * nwarps;
    if (sc.cap < n)
    {
        if (sc.buf) { CU (cudaDeviceSynchronize ()); cudaFree (sc.buf); sc.buf = NULL; sc.cap = 0; }
        const size_t cap = n + (n >> 3);
        const size_t cap_warps = (cap + SORT_WARP_KEYS - 1) / SORT_WARP_KEYS;
        CU (cudaMalloc (&sc.buf, (4 * cap + SORT_DIGITS * cap_warps + SORT_DIGITS) * sizeof (unsigned)));
        sc.cap = cap;
    }
    if (!sc.ev) CU (cudaEventCreateWithFlags (&sc.ev, cudaEventDisableTiming));
    if (sc.used && sc.last != st) CU (cudaStreamWaitEvent (st, sc.ev, 0));
    unsigned *keys = sc.buf, *keys1 = keys + sc.cap, *perm1 = keys1 + sc.cap, *perm2 = perm1 + sc.cap;
    unsigned *hist = perm2 + sc.cap, *totals = hist + hist_words;
    const unsigned wblocks = (nwarps + SORT_BLOCK/32 - 1) / (SORT_BLOCK/32);
    k_ray_keys<<<(unsigned) ((n + 255) / 256), 256, 0, st>>> (tree_ref (t), start, d_o, d_d, n, keys);
    k_sort_hist<<<wblocks, SORT_BLOCK, 0, st>>> (keys, n, 0, nwarps, hist);
    k_sort_scan_rows<<<SORT_DIGITS, 256, 0, st>>> (hist, nwarps, totals);
    k_sort_scatter<<<wblocks, SORT_BLOCK, 0, st>>> (keys, NULL, n, 0, nwarps, hist, totals, keys1, perm1);
    k_sort_hist<<<wblocks, SORT_BLOCK, 0, st>>> (keys1, n, SORT_BITS, nwarps, hist);
    k_sort_scan_rows<<<SORT_DIGITS, 256, 0, st>>> (hist, nwarps, totals);
    k_sort_scatter<<<wblocks, SORT_BLOCK, 0, st>>> (keys1, perm1, n, SORT_BITS, nwarps, hist, totals, NULL, perm2);
    g_launches += 7;
    CU (cudaGetLastError ());
    *perm = perm2;
    return 0;
}

static void sort_release (const voxgpu_tree *t, cudaStream_t st)
{
    SortScratch &sc = g_sort[t->device];
    if (sc.ev && cudaEventRecord (sc.ev, st) == cudaSuccess) { sc.last = st; sc.used = 1; }
}
