// Probe: how fast can a kernel store a 1920x1080x4 B frame into mapped pinned host
// memory, as a function of the store pattern, against a DMA copy of the same bytes?
// (development aid; results recorded in DESIGN.md)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf ("%s: %s\n", #x, cudaGetErrorString (e)); exit (1); } } while (0)
#define W 1920
#define H 1080

// warp = 8x4 tile, 32 B row segments (what k_render does)
__global__ void st_tile8x4 (unsigned *pix, unsigned v)
{
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int tiles_x = W / 8;
    if (warp >= tiles_x * (H / 4)) return;
    int x = (warp % tiles_x) * 8 + (lane & 7), y = (warp / tiles_x) * 4 + (lane >> 3);
    pix[(size_t) y * W + x] = v + x;
}
// warp = 16x2 tile, 64 B segments
__global__ void st_tile16x2 (unsigned *pix, unsigned v)
{
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int tiles_x = W / 16;
    if (warp >= tiles_x * (H / 2)) return;
    int x = (warp % tiles_x) * 16 + (lane & 15), y = (warp / tiles_x) * 2 + (lane >> 4);
    pix[(size_t) y * W + x] = v + x;
}
// warp = 32x1, 128 B lines
__global__ void st_row32 (unsigned *pix, unsigned v)
{
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t) W * H) pix[i] = v + (unsigned) i;
}
// warp = 32x4 region, each lane stores one uint4 (16 B): 4 x 128 B lines per store instruction
__global__ void st_32x4_v4 (unsigned *pix, unsigned v)
{
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int reg_x = W / 32;
    if (warp >= reg_x * (H / 4)) return;
    int x = (warp % reg_x) * 32 + (lane & 7) * 4, y = (warp / reg_x) * 4 + (lane >> 3);
    *(uint4*) (pix + (size_t) y * W + x) = make_uint4 (v, v + 1, v + 2, v + 3);
}
// fully linear uint4 stores (512 B per warp instruction)
__global__ void st_lin_v4 (unsigned *pix, unsigned v)
{
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t) W * H / 4) ((uint4*) pix)[i] = make_uint4 (v, v + 1, v + 2, v + 3);
}

template <typename F> static float timeit (F f, int reps)
{
    cudaEvent_t a, b; cudaEventCreate (&a); cudaEventCreate (&b);
    f (); CK (cudaDeviceSynchronize ());
    float best = 1e9;
    for (int r=0; r<reps; r++)
    {
        cudaEventRecord (a); f (); cudaEventRecord (b); CK (cudaEventSynchronize (b));
        float ms; cudaEventElapsedTime (&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main ()
{
    size_t bytes = (size_t) W * H * 4;
    unsigned *h, *dh, *d;
    CK (cudaHostAlloc (&h, bytes, cudaHostAllocMapped));
    CK (cudaHostGetDevicePointer (&dh, h, 0));
    CK (cudaMalloc (&d, bytes));
    int n = W * H;
    struct { const char *name; float ms; } r[16]; int k = 0;
    for (int target=0; target<2; target++)
    {
        unsigned *p = target ? dh : d;
        const char *t = target ? "mapped host" : "device";
        printf ("--- stores into %s memory (%.1f MB)\n", t, bytes / 1e6);
        float ms;
        ms = timeit ([&]{ st_tile8x4<<<(n + 127) / 128, 128>>> (p, 1); }, 20);  printf ("tile 8x4 (32 B segs)   %.4f ms %.1f GB/s\n", ms, bytes / ms / 1e6);
        ms = timeit ([&]{ st_tile16x2<<<(n + 127) / 128, 128>>> (p, 2); }, 20); printf ("tile 16x2 (64 B segs)  %.4f ms %.1f GB/s\n", ms, bytes / ms / 1e6);
        ms = timeit ([&]{ st_row32<<<(n + 127) / 128, 128>>> (p, 3); }, 20);    printf ("row 32x1 (128 B)       %.4f ms %.1f GB/s\n", ms, bytes / ms / 1e6);
        ms = timeit ([&]{ st_32x4_v4<<<(n / 4 + 127) / 128, 128>>> (p, 4); }, 20); printf ("32x4 uint4 (4x128 B)   %.4f ms %.1f GB/s\n", ms, bytes / ms / 1e6);
        ms = timeit ([&]{ st_lin_v4<<<(n / 4 + 127) / 128, 128>>> (p, 5); }, 20);  printf ("linear uint4 (512 B)   %.4f ms %.1f GB/s\n", ms, bytes / ms / 1e6);
    }
    float ms = timeit ([&]{ cudaMemcpyAsync (h, d, bytes, cudaMemcpyDeviceToHost, 0); }, 20);
    printf ("--- cudaMemcpyAsync D2H whole frame %.4f ms %.1f GB/s\n", ms, bytes / ms / 1e6);
    for (int parts=2; parts<=8; parts*=2)
    {
        ms = timeit ([&]{ for (int i=0; i<parts; i++) cudaMemcpyAsync ((char*) h + i * (bytes / parts), (char*) d + i * (bytes / parts), bytes / parts, cudaMemcpyDeviceToHost, 0); }, 20);
        printf ("    in %d parts %.4f ms %.1f GB/s\n", parts, ms, bytes / ms / 1e6);
    }
    return 0;
}
