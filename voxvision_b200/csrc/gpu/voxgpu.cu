/*
 * voxgpu.cu -- CUDA kernels (sm_100a) and the thin C-ABI of include/voxgpu.h.
 *
 *   K1 k_render_persistent / k_render_blocks (render_kernels.inl):
 *                 camera ray generation (simple-camera.c:21-31 / doom-camera.c:36-57)
 *                 + first-hit octree search (search.c:8-153) + get_color
 *                 (renderer.c:27-52) + linear framebuffer store, one thread per
 *                 pixel, one warp per 8x4 pixel tile.  Replaces the 4x4-block
 *                 dispatch_apply loop of renderer.c:214-318 and copy_squares
 *                 (copy-helper.c:5-40).
 *   K2 k_rays   : the same search for an array of independent rays.
 *   K3 k_balls  : vox_tree_ball_collidep (search.c:155-183) for an array of balls.
 *
 * Compile with -fmad=false (see traverse.cuh for why).
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <math.h>
#include <omp.h>
#include <atomic>
#include <mutex>
#include <thread>
#include <condition_variable>

#include "voxgpu.h"
#include "traverse.cuh"

/* ------------------------------------------------------------------ errors */

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

static int fail (const char *fmt, ...)
{
    va_list ap;
    va_start (ap, fmt);
    vsnprintf (g_err, sizeof g_err, fmt, ap);
    va_end (ap);
    return -1;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail ("%s failed: %s (%s:%d)", #call, cudaGetErrorString (e_), __FILE__, __LINE__); } while (0)
#define CU_NULL(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fail ("%s failed: %s (%s:%d)", #call, cudaGetErrorString (e_), __FILE__, __LINE__); return NULL; } } while (0)

extern "C" const char* voxgpu_last_error (void) { return g_err; }
extern "C" unsigned long long voxgpu_launch_count (void) { return g_launches.load (); }

/* ------------------------------------------------------------------ device */

extern "C" int voxgpu_device_count (void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount (&n);
    if (e != cudaSuccess) { fail ("cudaGetDeviceCount: %s", cudaGetErrorString (e)); return 0; }
    return n;
}
extern "C" int voxgpu_set_device (int device) { CU (cudaSetDevice (device)); return 0; }
extern "C" int voxgpu_get_device (void) { int d = -1; if (cudaGetDevice (&d) != cudaSuccess) return -1; return d; }
extern "C" int voxgpu_sm_count (void)
{
    int d = 0, n = 0;
    CU (cudaGetDevice (&d));
    CU (cudaDeviceGetAttribute (&n, cudaDevAttrMultiProcessorCount, d));
    return n;
}
extern "C" int voxgpu_synchronize (voxgpu_stream s)
{
    if (s) CU (cudaStreamSynchronize ((cudaStream_t) s));
    else CU (cudaDeviceSynchronize ());
    return 0;
}

extern "C" void* voxgpu_malloc (size_t bytes)
{
    void *p = NULL;
    CU_NULL (cudaMalloc (&p, bytes ? bytes : 16));
    return p;
}
extern "C" void voxgpu_free (void *p) { if (p) cudaFree (p); }
extern "C" void* voxgpu_host_alloc (size_t bytes)
{
    void *p = NULL;
    CU_NULL (cudaHostAlloc (&p, bytes ? bytes : 16, cudaHostAllocMapped | cudaHostAllocPortable));
    return p;
}
extern "C" void voxgpu_host_free (void *p) { if (p) cudaFreeHost (p); }
extern "C" int voxgpu_host_register (void *p, size_t bytes)
{
    CU (cudaHostRegister (p, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    return 0;
}
extern "C" int voxgpu_host_unregister (void *p) { CU (cudaHostUnregister (p)); return 0; }
extern "C" void* voxgpu_host_device_pointer (void *p)
{
    void *d = NULL;
    CU_NULL (cudaHostGetDevicePointer (&d, p, 0));
    return d;
}
extern "C" int voxgpu_memcpy_h2d (void *d, const void *h, size_t n, voxgpu_stream s)
{ CU (cudaMemcpyAsync (d, h, n, cudaMemcpyHostToDevice, (cudaStream_t) s)); return 0; }
extern "C" int voxgpu_memcpy_d2h (void *h, const void *d, size_t n, voxgpu_stream s)
{ CU (cudaMemcpyAsync (h, d, n, cudaMemcpyDeviceToHost, (cudaStream_t) s)); return 0; }
extern "C" int voxgpu_memset (void *d, int v, size_t n, voxgpu_stream s)
{ CU (cudaMemsetAsync (d, v, n, (cudaStream_t) s)); return 0; }
extern "C" voxgpu_stream voxgpu_stream_create (void)
{
    cudaStream_t s = NULL;
    CU_NULL (cudaStreamCreateWithFlags (&s, cudaStreamNonBlocking));
    return (voxgpu_stream) s;
}
extern "C" void voxgpu_stream_destroy (voxgpu_stream s) { if (s) cudaStreamDestroy ((cudaStream_t) s); }
extern "C" int voxgpu_stream_join (voxgpu_stream waiter, voxgpu_stream other)
{
    /* everything queued on `other` so far must finish before what is queued on `waiter` next */
    cudaEvent_t e;
    CU (cudaEventCreateWithFlags (&e, cudaEventDisableTiming));
    cudaError_t rc = cudaEventRecord (e, (cudaStream_t) other);
    if (rc == cudaSuccess) rc = cudaStreamWaitEvent ((cudaStream_t) waiter, e, 0);
    cudaEventDestroy (e);
    if (rc != cudaSuccess) return fail ("voxgpu_stream_join: %s", cudaGetErrorString (rc));
    return 0;
}

/* ------------------------------------------------ peer memory (multi-GPU) */

/* One process per GPU: rank 0 exports a cudaMalloc'ed buffer, the other ranks map it and
   hand the mapped address to voxgpu_render as the framebuffer -- the render kernel then
   stores its pixels straight into GPU 0's memory over NVLink (the "gather" of the
   multi-GPU path, fused into K1's store). */
extern "C" int voxgpu_ipc_export (const void *dptr, unsigned char handle[VOXGPU_IPC_HANDLE_BYTES])
{
    static_assert (sizeof (cudaIpcMemHandle_t) == VOXGPU_IPC_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    CU (cudaIpcGetMemHandle (&h, (void*) dptr));
    memcpy (handle, &h, sizeof h);
    return 0;
}
extern "C" void* voxgpu_ipc_open (const unsigned char handle[VOXGPU_IPC_HANDLE_BYTES])
{
    cudaIpcMemHandle_t h;
    void *p = NULL;
    memcpy (&h, handle, sizeof h);
    CU_NULL (cudaIpcOpenMemHandle (&p, h, cudaIpcMemLazyEnablePeerAccess));
    return p;
}
extern "C" int voxgpu_ipc_close (void *dptr) { CU (cudaIpcCloseMemHandle (dptr)); return 0; }
/* One process, several devices: let kernels of the CURRENT device store into memory of `peer` (the same
   NVLink path as the IPC mapping above, without a second process). */
extern "C" int voxgpu_enable_peer (int peer)
{
    cudaError_t e = cudaDeviceEnablePeerAccess (peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError (); return 0; }
    if (e != cudaSuccess) return fail ("cudaDeviceEnablePeerAccess (%d): %s", peer, cudaGetErrorString (e));
    return 0;
}

/* ------------------------------------------------------------------- timer */

struct voxgpu_timer { cudaEvent_t a, b; };
extern "C" voxgpu_timer* voxgpu_timer_create (void)
{
    voxgpu_timer *t = (voxgpu_timer*) calloc (1, sizeof *t);
    if (cudaEventCreate (&t->a) != cudaSuccess || cudaEventCreate (&t->b) != cudaSuccess)
    { fail ("cudaEventCreate failed"); free (t); return NULL; }
    return t;
}
extern "C" void voxgpu_timer_destroy (voxgpu_timer *t)
{ if (!t) return; cudaEventDestroy (t->a); cudaEventDestroy (t->b); free (t); }
extern "C" int voxgpu_timer_start (voxgpu_timer *t, voxgpu_stream s) { CU (cudaEventRecord (t->a, (cudaStream_t) s)); return 0; }
extern "C" int voxgpu_timer_stop (voxgpu_timer *t, voxgpu_stream s) { CU (cudaEventRecord (t->b, (cudaStream_t) s)); return 0; }
extern "C" float voxgpu_timer_elapsed_ms (voxgpu_timer *t)
{
    float ms = -1.0f;
    if (cudaEventSynchronize (t->b) != cudaSuccess) { fail ("cudaEventSynchronize failed"); return -1.0f; }
    if (cudaEventElapsedTime (&ms, t->a, t->b) != cudaSuccess) { fail ("cudaEventElapsedTime failed"); return -1.0f; }
    return ms;
}

#include "tree.inl"
#include "sort.inl"
#include "queries.inl"
#include "volume.inl"
#include "render_kernels.inl"
#include "render_launch.inl"
#include "present.inl"
