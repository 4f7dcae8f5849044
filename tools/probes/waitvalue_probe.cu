// Probe: can a copy stream follow the progress of a RUNNING kernel?  A kernel "finishes" S strips
// one after another (spinning `gap` ns between them, counter released after each strip's stores);
// a second stream holds S x { cuStreamWaitValue32 (counter[s] >= target) ; cudaMemcpyAsync (strip) }.
// Reports when (relative to the kernel's start) each strip arrived in pinned host memory.
// The driver entry point is fetched through the runtime (no link against libcuda).
// (development aid; results recorded in DESIGN.md)
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf ("%s: %s\n", #x, cudaGetErrorString (e)); exit (1); } } while (0)
#define S 16
#define STRIP_WORDS (1920 * 68)

__device__ unsigned long long gtimer () { unsigned long long t; asm volatile ("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void producer (unsigned *fb, unsigned *counter, unsigned target, unsigned long long gap_ns, unsigned v, unsigned long long *t_start)
{
    // one CTA per strip chunk; CTA b of strip s waits until s*gap after the kernel's start
    const int s = blockIdx.x / 8, part = blockIdx.x % 8;
    if (blockIdx.x == 0 && threadIdx.x == 0) *t_start = gtimer ();
    const unsigned long long t0 = gtimer ();
    while (gtimer () - t0 < gap_ns * (unsigned long long) (s + 1)) ;
    const int per = STRIP_WORDS / 8;
    for (int i = threadIdx.x; i < per; i += blockDim.x) fb[(size_t) s * STRIP_WORDS + part * per + i] = v + s;
    __syncthreads ();
    if (threadIdx.x == 0) { __threadfence (); atomicAdd (counter + s, 1u); }
}

int main (int argc, char **argv)
{
    const int NS = argc > 1 ? atoi (argv[1]) : 1;
    const unsigned long long gap = argc > 2 ? strtoull (argv[2], 0, 10) : 10000ull;
    const int use_graph = argc > 3 ? atoi (argv[3]) : 0;
    typedef CUresult (*wait_fn) (CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    wait_fn waitv = NULL;
    cudaDriverEntryPointQueryResult qr;
    CK (cudaGetDriverEntryPoint ("cuStreamWaitValue32", (void**) &waitv, cudaEnableDefault, &qr));
    printf ("cuStreamWaitValue32 entry point: %p (query %d)\n", (void*) waitv, (int) qr);
    if (!waitv) return 1;
    int memops = 0;
    cudaDeviceGetAttribute (&memops, (cudaDeviceAttr) 0, 0);
    unsigned *fb, *counter, *host;
    unsigned long long *t_start;
    CK (cudaMalloc (&fb, (size_t) S * STRIP_WORDS * 4));
    CK (cudaMalloc (&counter, S * 4));
    CK (cudaMemset (counter, 0, S * 4));
    CK (cudaMalloc (&t_start, 8));
    CK (cudaHostAlloc (&host, (size_t) S * STRIP_WORDS * 4, cudaHostAllocDefault));
    cudaStream_t sk, scs[16];
    CK (cudaStreamCreateWithFlags (&sk, cudaStreamNonBlocking));
    for (int i=0; i<NS; i++) CK (cudaStreamCreateWithFlags (&scs[i], cudaStreamNonBlocking));
    cudaEvent_t ed[16];
    for (int i=0; i<NS; i++) CK (cudaEventCreate (&ed[i]));
    cudaEvent_t e0, e1, ek, es[S];
    CK (cudaEventCreate (&e0)); CK (cudaEventCreate (&e1)); CK (cudaEventCreate (&ek));
    for (int s=0; s<S; s++) CK (cudaEventCreate (&es[s]));
    if (use_graph)
    {
        // the whole frame as ONE graph: memset(counters) -> kernel, and NS branches of {wait; copy} pairs
        cudaGraph_t graph; cudaGraphExec_t exec;
        cudaEvent_t fork, join[16];
        CK (cudaEventCreate (&fork));
        for (int i=0; i<NS; i++) CK (cudaEventCreate (&join[i]));
        CK (cudaStreamBeginCapture (sk, cudaStreamCaptureModeGlobal));
        CK (cudaMemsetAsync (counter, 0, S * 4, sk));
        CK (cudaEventRecord (fork, sk));
        for (int i=0; i<NS; i++) CK (cudaStreamWaitEvent (scs[i], fork, 0));
        producer<<<S * 8, 256, 0, sk>>> (fb, counter, 0, gap, 7000u, t_start);
        for (int s=0; s<S; s++)
        {
            cudaStream_t sc = scs[s % NS];
            CUresult r = waitv ((CUstream) sc, (CUdeviceptr) (counter + s), 8u, CU_STREAM_WAIT_VALUE_GEQ);
            if (r != CUDA_SUCCESS) { printf ("capture of cuStreamWaitValue32 failed: %d\n", (int) r); return 1; }
            CK (cudaMemcpyAsync (host + (size_t) s * STRIP_WORDS, fb + (size_t) s * STRIP_WORDS, STRIP_WORDS * 4, cudaMemcpyDeviceToHost, sc));
        }
        for (int i=0; i<NS; i++) { CK (cudaEventRecord (join[i], scs[i])); CK (cudaStreamWaitEvent (sk, join[i], 0)); }
        CK (cudaStreamEndCapture (sk, &graph));
        CK (cudaGraphInstantiate (&exec, graph, 0));
        for (int frame = 1; frame <= 5; frame++)
        {
            memset (host, 0, (size_t) S * STRIP_WORDS * 4);
            CK (cudaEventRecord (e0, sk));
            CK (cudaGraphLaunch (exec, sk));
            CK (cudaEventRecord (e1, sk));
            CK (cudaEventSynchronize (e1));
            float t1; CK (cudaEventElapsedTime (&t1, e0, e1));
            int bad = 0;
            for (int s=0; s<S; s++) for (int i=0; i<STRIP_WORDS; i += 997) bad += host[(size_t) s * STRIP_WORDS + i] != 7000u + s;
            printf ("graph frame %d: all strips in host memory at %.1f us (kernel alone ~%.0f us), bad words %d\n", frame, t1 * 1e3, (double) gap * S / 1e3, bad);
        }
        return 0;
    }
    for (int frame = 1; frame <= 4; frame++)
    {
        memset (host, 0, (size_t) S * STRIP_WORDS * 4);
        CK (cudaEventRecord (e0, sk));
        for (int i=0; i<NS; i++) CK (cudaStreamWaitEvent (scs[i], e0, 0));
        producer<<<S * 8, 256, 0, sk>>> (fb, counter, 0, gap, 1000u * frame, t_start);
        CK (cudaEventRecord (ek, sk));
        for (int s=0; s<S; s++)
        {
            cudaStream_t sc = scs[s % NS];
            CUresult r = waitv ((CUstream) sc, (CUdeviceptr) (counter + s), (cuuint32_t) (8 * frame), CU_STREAM_WAIT_VALUE_GEQ);
            if (r != CUDA_SUCCESS) { printf ("cuStreamWaitValue32 failed: %d\n", (int) r); return 1; }
            CK (cudaMemcpyAsync (host + (size_t) s * STRIP_WORDS, fb + (size_t) s * STRIP_WORDS, STRIP_WORDS * 4, cudaMemcpyDeviceToHost, sc));
            CK (cudaEventRecord (es[s], sc));
        }
        for (int i=0; i<NS; i++) { CK (cudaEventRecord (ed[i], scs[i])); CK (cudaStreamWaitEvent (sk, ed[i], 0)); }
        CK (cudaEventRecord (e1, sk));
        CK (cudaEventSynchronize (e1));
        CK (cudaStreamSynchronize (sk));
        float tk, t1;
        CK (cudaEventElapsedTime (&tk, e0, ek)); CK (cudaEventElapsedTime (&t1, e0, e1));
        int bad = 0;
        for (int s=0; s<S; s++) for (int i=0; i<STRIP_WORDS; i += 997) bad += host[(size_t) s * STRIP_WORDS + i] != 1000u * frame + s;
        printf ("frame %d: kernel done at %.1f us, last strip in host memory at %.1f us, bad words %d; strip arrivals (us):", frame, tk * 1e3, t1 * 1e3, bad);
        for (int s=0; s<S; s++) { float t; CK (cudaEventElapsedTime (&t, e0, es[s])); printf (" %.0f", t * 1e3); }
        printf ("\n");
    }
    return 0;
}
