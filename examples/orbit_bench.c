/*
 * orbit_bench.c -- the end-to-end leg of bench.py, in C against the reference-facing API only
 * (vox_context_set_camera / vox_render / vox_redraw, src/demo/main.c:226-237): a caller that renders a
 * camera path into host surfaces and looks at every finished frame.
 *
 * `contexts` render contexts (each with its own page-locked surface: a swap chain) take the frames of the
 * path round-robin; a context is redrawn, and the centre pixel of its surface read, right before it is
 * given its next frame.  contexts = 1 is the reference demo's loop: render, redraw, render, redraw.
 * `devices` goes to vox_context_set_devices (1 = the current device, 0 = all devices of this process).
 * One untimed pass over the path, then `steps` timed passes (CLOCK_MONOTONIC around them; vox_redraw of the
 * last frames is inside).  Built as a shared object so that bench.py can call it in-process
 * (voxvision_b200/Makefile: lib/liborbitbench.so); nothing here touches CUDA or libvoxgpu directly.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <voxtrees.h>
#include <voxrnd.h>

struct orbit_bench_result
{
    double seconds;                 /* wall time of the timed passes */
    unsigned long long checksum;    /* xor of the centre pixels read: the frames were looked at */
    long frames;                    /* frames rendered in the timed passes */
    int devices;                    /* what vox_context_set_devices answered */
    double device_ms_mean;          /* contexts == 1: mean vox_context_last_kernel_ms over one extra, untimed pass
                                       (first launch of a frame -> last byte of it in host memory, CUDA events) */
    char error[256];
};

#define MAX_CONTEXTS 64

static double now (void)
{
    struct timespec ts;
    clock_gettime (CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* poses: 6 floats per frame (position xyz, look-at xyz) */
__attribute__((visibility ("default")))
int orbit_bench_run (struct vox_node *tree, int w, int h, int nframes, const float *poses, float fov,
                     int contexts, int devices, unsigned int quality, int steps, struct orbit_bench_result *out)
{
    struct vox_rnd_ctx *ctx[MAX_CONTEXTS] = {0};
    struct vox_camera *cam[MAX_CONTEXTS] = {0};
    int rc = -1;
    memset (out, 0, sizeof *out);
    if (contexts < 1 || contexts > MAX_CONTEXTS || nframes < 1 || steps < 1) { snprintf (out->error, sizeof out->error, "bad arguments"); return -1; }
    struct vox_camera_interface *methods = vox_camera_methods ("simple-camera");
    if (!methods) { snprintf (out->error, sizeof out->error, "no simple-camera module: %s", vox_last_error ()); return -1; }
    for (int i=0; i<contexts; i++)
    {
        ctx[i] = vox_make_context_and_window (w, h);
        cam[i] = methods->construct_camera (NULL);
        if (!ctx[i] || !cam[i]) { snprintf (out->error, sizeof out->error, "no context / camera"); goto done; }
        out->devices = vox_context_set_devices (ctx[i], devices);
        vox_context_set_scene (ctx[i], tree);
        if (!vox_context_set_quality (ctx[i], quality)) { snprintf (out->error, sizeof out->error, "bad quality %u", quality); goto done; }
        cam[i]->iface->set_property_number (cam[i], "fov", fov);
    }
    const size_t centre = (size_t) (h / 2) * w + w / 2;
    double t0 = 0;
    for (int pass = 0; pass <= steps; pass++)
    {
        if (pass == 1) t0 = now ();
        unsigned long long acc = 0;
        for (int f = 0; f < nframes + contexts; f++)
        {
            const int k = f % contexts;
            if (f >= contexts)
            {
                vox_redraw (ctx[k]);                                            /* frame f - contexts is in its surface */
                acc ^= ((const Uint32*) ctx[k]->surface->pixels)[centre];
            }
            if (f >= nframes) continue;
            vox_dot pos, look;
            vox_dot_set (pos, poses[6*f], poses[6*f+1], poses[6*f+2]);
            vox_dot_set (look, poses[6*f+3], poses[6*f+4], poses[6*f+5]);
            cam[k]->iface->set_property_dot (cam[k], "position", pos);
            cam[k]->iface->look_at (cam[k], look);
            vox_context_set_camera (ctx[k], cam[k]);
            vox_render (ctx[k]);
            if (vox_last_error ()[0]) { snprintf (out->error, sizeof out->error, "vox_render: %s", vox_last_error ()); goto done; }
        }
        if (pass) { out->checksum ^= acc; out->frames += nframes; }
    }
    out->seconds = now () - t0;
    if (contexts == 1)
    {
        double sum = 0;
        for (int f = 0; f < nframes; f++)
        {
            vox_dot pos, look;
            vox_dot_set (pos, poses[6*f], poses[6*f+1], poses[6*f+2]);
            vox_dot_set (look, poses[6*f+3], poses[6*f+4], poses[6*f+5]);
            cam[0]->iface->set_property_dot (cam[0], "position", pos);
            cam[0]->iface->look_at (cam[0], look);
            vox_context_set_camera (ctx[0], cam[0]);
            vox_render (ctx[0]);
            vox_redraw (ctx[0]);
            sum += vox_context_last_kernel_ms (ctx[0]);
        }
        out->device_ms_mean = sum / nframes;
    }
    rc = 0;
done:
    for (int i=0; i<contexts; i++)
    {
        if (ctx[i]) vox_destroy_context (ctx[i]);
        if (cam[i]) cam[i]->iface->destroy_camera (cam[i]);
    }
    return rc;
}
