/*
 * orbit.c -- the reference's scene setup (src/demo/main.c:163-237) written against this
 * repository's headers, in C, with nothing but the reference's own API: build a tree, make a
 * camera through the module interface, make a context, render a few frames of an orbit and
 * print, per frame, the number of lit pixels and an FNV-1a hash of the surface.
 *
 *   gcc -std=gnu11 -Iinclude examples/orbit.c -Lvoxvision_b200/lib -lvoxrnd -lvoxtrees \
 *       -Wl,-rpath,$PWD/voxvision_b200/lib -o orbit && ./orbit [frames]
 *
 * The scene is the reference's own test scene (a ball of voxels, tests/tests.c:214-237, radius 20
 * here) so that no data file is needed.  tests/test_gpu_parity.py::test_c_example_matches_the_oracle
 * builds and runs this program and compares the hashes with the oracle's frames.
 */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <voxtrees.h>
#include <voxrnd.h>

int main (int argc, char *argv[])
{
    int frames = argc > 1 ? atoi (argv[1]) : 4;
    const int radius = 20, w = 256, h = 128;

    size_t n = 0, cap = 8 * radius * radius * radius;
    vox_dot *set = vox_alloc (cap * sizeof (vox_dot));
    for (int x=-radius; x<radius; x++)
        for (int y=-radius; y<radius; y++)
            for (int z=-radius; z<radius; z++)
                if (x*x + y*y + z*z < radius*radius) { vox_dot_set (set[n], x, y, z); n++; }
    struct vox_node *tree = vox_make_tree (set, n);
    free (set);
    if (!tree) { fprintf (stderr, "vox_make_tree failed: %s\n", vox_last_error ()); return 1; }

    struct vox_camera_interface *methods = vox_camera_methods ("simple-camera");
    if (!methods) { fprintf (stderr, "no simple-camera module: %s\n", vox_last_error ()); return 1; }
    struct vox_camera *camera = methods->construct_camera (NULL);
    struct vox_rnd_ctx *ctx = vox_make_context_and_window (w, h);
    if (!ctx) { fprintf (stderr, "no context\n"); return 1; }
    vox_context_set_scene (ctx, tree);
    vox_context_set_camera (ctx, camera);
    if (!vox_context_set_quality (ctx, VOX_QUALITY_BEST)) return 1;

    for (int f=0; f<frames; f++)
    {
        double th = f * 2.0 * M_PI / 16;
        vox_dot pos, centre;
        vox_dot_set (pos, (float) (70.0 * sin (th)), (float) (-70.0 * cos (th)), 10.0f);
        vox_dot_set (centre, 0, 0, 0);
        camera->iface->set_property_dot (camera, "position", pos);
        camera->iface->set_property_number (camera, "fov", 1.0f);
        camera->iface->look_at (camera, centre);
        vox_context_set_camera (ctx, camera);      /* set_window_size, as the demo does after moving */
        vox_render (ctx);
        vox_redraw (ctx);
        if (vox_last_error ()[0]) { fprintf (stderr, "render failed: %s\n", vox_last_error ()); return 2; }
        const Uint32 *pix = ctx->surface->pixels;
        unsigned long long hash = 1469598103934665603ull;
        long lit = 0;
        for (int i=0; i<w*h; i++)
        {
            lit += pix[i] != 0;
            for (int b=0; b<4; b++) { hash ^= (pix[i] >> (8*b)) & 0xff; hash *= 1099511628211ull; }
        }
        printf ("frame %d lit %ld hash %016llx\n", f, lit, hash);
    }
    vox_destroy_context (ctx);
    camera->iface->destroy_camera (camera);
    vox_destroy_tree (tree);
    return 0;
}
