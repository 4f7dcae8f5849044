/* volume.inl -- part of the single translation unit voxgpu.cu (included there, in this order):
   raw volume -> voxel list on the device (threshold + stream compaction), voxgpu_volume_extract. */
/* ------------------------------------------------------ raw volume -> voxels
 *
 * The reference's loader (src/voxtrees/datareader.c:51-76) walks the samples x outer / z inner, assembles
 * each from `samplesize` little-endian bytes (value |= byte << 8l), asks a Block whether to keep it and
 * appends {i*vox_voxel[0], j*vox_voxel[1], k*vox_voxel[2]}; vox_make_tree then takes the list.  The order of
 * the list matters downstream (find_center sums it in array order, tree.c:43-45), so the device version is a
 * STABLE compaction: every thread owns VOL_PER_THREAD consecutive samples, every CTA VOL_TILE of them,
 *   pass 1  count the kept samples of every CTA's tile,
 *   scan    exclusive prefix over the tiles (one CTA; 2^30 samples are 262 144 tiles),
 *   pass 2  recount per thread, block-scan, write the voxels of the tile at its offset in sample order.
 * The predicate is the range test of the Lua binding (modules/voxtrees.c:306-342): min <= sample < max.
 * Purely HBM-bound byte work: 2 reads of the volume + 16 B per kept voxel.
 */
#define VOL_BLOCK 256
#define VOL_PER_THREAD 16
#define VOL_TILE (VOL_BLOCK * VOL_PER_THREAD)

__device__ __forceinline__ unsigned vol_sample (const unsigned char * __restrict__ v, size_t s, unsigned samplesize)
{
    unsigned value = 0;
    for (unsigned l = 0; l < samplesize; l++) value |= (unsigned) v[s * samplesize + l] << (8 * l);     /* datareader.c:65 */
    return value;
}

/* kept samples among this thread's VOL_PER_THREAD, as a bit mask */
__device__ __forceinline__ unsigned vol_keep_mask (const unsigned char * __restrict__ v, size_t first, size_t n, unsigned samplesize,
                                                   unsigned lo, unsigned long long hi)
{
    unsigned m = 0;
    if (samplesize == 1 && first + VOL_PER_THREAD <= n)
    {
        /* the common case (8-bit volumes): the thread's 16 samples are one aligned 16-byte load */
        const uint4 q = *(const uint4*) (v + first);
        const unsigned w[4] = { q.x, q.y, q.z, q.w };
        #pragma unroll
        for (int j = 0; j < VOL_PER_THREAD; j++)
        {
            const unsigned s = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
            if (s >= lo && (unsigned long long) s < hi) m |= 1u << j;
        }
        return m;
    }
    for (int j = 0; j < VOL_PER_THREAD; j++)
    {
        const size_t s = first + j;
        if (s >= n) break;
        const unsigned val = vol_sample (v, s, samplesize);
        if (val >= lo && (unsigned long long) val < hi) m |= 1u << j;
    }
    return m;
}

__global__ void __launch_bounds__(VOL_BLOCK)
k_vol_count (const unsigned char * __restrict__ v, size_t n, unsigned samplesize, unsigned lo, unsigned long long hi,
             unsigned * __restrict__ tile_count)
{
    __shared__ unsigned wsum[VOL_BLOCK/32];
    const size_t first = ((size_t) blockIdx.x * VOL_BLOCK + threadIdx.x) * VOL_PER_THREAD;
    unsigned c = first < n ? __popc (vol_keep_mask (v, first, n, samplesize, lo, hi)) : 0u;
    #pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync (0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads ();
    if (threadIdx.x == 0)
    {
        unsigned t = 0;
        for (int w = 0; w < VOL_BLOCK/32; w++) t += wsum[w];
        tile_count[blockIdx.x] = t;
    }
}

/* exclusive prefix of tile_count[0..ntiles) into tile_off (64-bit), total into *total; ONE CTA of 1024 threads */
__global__ void __launch_bounds__(1024)
k_vol_scan (const unsigned * __restrict__ tile_count, size_t ntiles, unsigned long long * __restrict__ tile_off,
            unsigned long long * __restrict__ total)
{
    __shared__ unsigned long long part[1024];
    const size_t per = (ntiles + 1023) / 1024;
    const size_t lo = (size_t) threadIdx.x * per, hi = lo + per < ntiles ? lo + per : ntiles;
    unsigned long long s = 0;
    for (size_t i = lo; i < hi; i++) s += tile_count[i];
    part[threadIdx.x] = s;
    __syncthreads ();
    for (unsigned off = 1; off < 1024; off <<= 1)
    {
        const unsigned long long x = threadIdx.x >= off ? part[threadIdx.x - off] : 0ull;
        __syncthreads ();
        part[threadIdx.x] += x;
        __syncthreads ();
    }
    unsigned long long run = part[threadIdx.x] - s;
    for (size_t i = lo; i < hi; i++) { tile_off[i] = run; run += tile_count[i]; }
    if (threadIdx.x == 1023) *total = part[1023];
}

__global__ void __launch_bounds__(VOL_BLOCK)
k_vol_emit (const unsigned char * __restrict__ v, size_t n, unsigned samplesize, unsigned lo, unsigned long long hi,
            unsigned dim1, unsigned dim2, float vx, float vy, float vz,
            const unsigned long long * __restrict__ tile_off, float4 * __restrict__ out)
{
    __shared__ unsigned wsum[VOL_BLOCK/32];
    const size_t first = ((size_t) blockIdx.x * VOL_BLOCK + threadIdx.x) * VOL_PER_THREAD;
    const unsigned keep = first < n ? vol_keep_mask (v, first, n, samplesize, lo, hi) : 0u;
    const unsigned c = __popc (keep);
    /* exclusive prefix of c over the CTA, in thread order = sample order */
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned incl = c;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned x = __shfl_up_sync (0xffffffffu, incl, o); if ((int) lane >= o) incl += x; }
    if (lane == 31) wsum[wib] = incl;
    __syncthreads ();
    unsigned before = 0;
    for (unsigned w = 0; w < wib; w++) before += wsum[w];
    unsigned long long pos = tile_off[blockIdx.x] + before + incl - c;
    unsigned m = keep;
    while (m)
    {
        const int j = __ffs (m) - 1;
        m &= m - 1;
        const size_t s = first + j;
        const unsigned k = (unsigned) (s % dim2);
        const size_t ij = s / dim2;
        const unsigned jj = (unsigned) (ij % dim1), i = (unsigned) (ij / dim1);
        /* datareader.c:69-71: unsigned index converted to float, one rounded product */
        out[pos++] = make_float4 (__fmul_rn ((float) i, vx), __fmul_rn ((float) jj, vy), __fmul_rn ((float) k, vz), 0.0f);
    }
}

extern "C" int voxgpu_volume_extract (const void *volume, const unsigned dim[3], unsigned samplesize,
                                      unsigned lo, unsigned long long hi, const float voxel[3],
                                      float **dots, size_t *count)
{
    if (!volume || !dim || !voxel || !dots || !count) return fail ("voxgpu_volume_extract: NULL argument");
    if (samplesize < 1 || samplesize > 4) return fail ("voxgpu_volume_extract: unsupported sample size %u", samplesize);
    *dots = NULL; *count = 0;
    const size_t n = (size_t) dim[0] * dim[1] * dim[2];
    if (n == 0) return 0;
    const size_t ntiles = (n + VOL_TILE - 1) / VOL_TILE;
    if (ntiles > 0x7fffffffull) return fail ("voxgpu_volume_extract: volume too large");
    unsigned char *d_vol = NULL;
    unsigned *d_cnt = NULL;
    unsigned long long *d_off = NULL, total = 0;
    float4 *d_out = NULL;
    float *h_out = NULL;
    cudaStream_t st = NULL;
    cudaError_t e;
#define VOL_TRY(call) do { if ((e = (call)) != cudaSuccess) { fail ("voxgpu_volume_extract: %s: %s", #call, cudaGetErrorString (e)); goto bad; } } while (0)
    VOL_TRY (cudaStreamCreateWithFlags (&st, cudaStreamNonBlocking));
    VOL_TRY (cudaMalloc (&d_vol, n * samplesize + 16));
    VOL_TRY (cudaMalloc (&d_cnt, ntiles * sizeof (unsigned)));
    VOL_TRY (cudaMalloc (&d_off, (ntiles + 1) * sizeof (unsigned long long)));
    VOL_TRY (cudaMemcpyAsync (d_vol, volume, n * samplesize, cudaMemcpyHostToDevice, st));
    k_vol_count<<<(unsigned) ntiles, VOL_BLOCK, 0, st>>> (d_vol, n, samplesize, lo, hi, d_cnt);
    k_vol_scan<<<1, 1024, 0, st>>> (d_cnt, ntiles, d_off, d_off + ntiles);
    g_launches += 2;
    VOL_TRY (cudaGetLastError ());
    VOL_TRY (cudaMemcpyAsync (&total, d_off + ntiles, sizeof total, cudaMemcpyDeviceToHost, st));
    VOL_TRY (cudaStreamSynchronize (st));
    if (total)
    {
        VOL_TRY (cudaMalloc (&d_out, (size_t) total * sizeof (float4)));
        VOL_TRY (cudaHostAlloc (&h_out, (size_t) total * sizeof (float4), cudaHostAllocDefault));
        k_vol_emit<<<(unsigned) ntiles, VOL_BLOCK, 0, st>>> (d_vol, n, samplesize, lo, hi, dim[1], dim[2], voxel[0], voxel[1], voxel[2],
                                                             d_off, d_out);
        g_launches++;
        VOL_TRY (cudaGetLastError ());
        VOL_TRY (cudaMemcpyAsync (h_out, d_out, (size_t) total * sizeof (float4), cudaMemcpyDeviceToHost, st));
        VOL_TRY (cudaStreamSynchronize (st));
    }
#undef VOL_TRY
    cudaFree (d_vol); cudaFree (d_cnt); cudaFree (d_off); cudaFree (d_out);
    cudaStreamDestroy (st);
    *dots = h_out; *count = (size_t) total;
    return 0;
bad:
    cudaFree (d_vol); cudaFree (d_cnt); cudaFree (d_off); cudaFree (d_out);
    if (h_out) cudaFreeHost (h_out);
    if (st) cudaStreamDestroy (st);
    return -1;
}
