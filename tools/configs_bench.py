#!/usr/bin/env python
"""Throughput of the configs bench.py does not cover: C1 (bonsai 800x600 frame),
C3 (16 Mi random rays vs the ~10 M-voxel random tree, K2 k_rays) and C4 (4 Mi balls vs the skull
tree, K3 k_balls).  One GPU.  For each: device-resident rate (CUDA events, buffers in HBM), the
rate through the host-buffer C API (copies inside the timed region), algorithmic bytes per query
from the instrumented traversal, and the reference's CPU code on a bounded sample on all host
threads.  Prints one JSON line per config.

    python tools/configs_bench.py [c1] [c3] [c4] [modes]
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import voxvision_b200 as vv  # noqa: E402
from voxvision_b200 import api  # noqa: E402
from tools import datasets as ds  # noqa: E402
from oracle import orc  # noqa: E402  (CPU baseline + visit counts only)

BYTES_INNER, BYTES_LEAF, BYTES_DENSE, BYTES_DOT = 48, 32, 32, 16
THREADS = os.cpu_count() or 1
g = api.libs.gpu


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6550.0


def cpu_oracle():
    for which in ("ref_sse", "ref_fast", "ref", "port"):
        if orc.available(which):
            return orc.Oracle(which)
    return orc.Oracle("port")


def timed(fn, reps, flush=None):
    t = vv.Timer()
    ms = []
    for _ in range(reps):
        if flush:
            flush()
        t.start()
        fn()
        t.stop()
        ms.append(t.elapsed_ms())
    return float(np.mean(ms)), float(np.min(ms))


def c1():
    W, H = 800, 600
    vox = ds.volume_voxels("bonsai")
    tree = vv.Tree.build(vox)
    dev = tree.device_handle()
    ctx = vv.Context(W, H)
    cam = ds.setup_camera(vv.Camera(), (128, 128, 400), (128, 128, 128), 1.2, W, H)
    gc = cam.gpu_camera()
    d_tex, d_pix = vv.DeviceBuffer.from_array(ctx.texture()), vv.DeviceBuffer(W * H * 4)
    fr = api.GpuFrame(W, H, 0, H, W, 0, ctx.pixfmt(), None)
    render = lambda: g.voxgpu_render(dev, C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, None, None, None)
    for _ in range(100):        # more than the 16 mask / counter slots: the first use of a slot allocates
        render()
    mean, best = timed(render, 200)
    ctx.set_scene(tree); ctx.set_camera(cam); ctx.set_quality(0)
    ctx.render(); ctx.redraw()
    t0 = time.perf_counter()
    for _ in range(200):
        ctx.render(); ctx.redraw()
    e2e = (time.perf_counter() - t0) / 200
    O = cpu_oracle()
    ot = O.tree(vox)
    ocam = ds.setup_camera(O.camera(0), (128, 128, 400), (128, 128, 128), 1.2, W, H)
    ot.render(ocam, W, H, 0, threads=THREADS, want_hits=False)
    cpu = min(_t(lambda: ot.render(ocam, W, H, 0, threads=THREADS, want_hits=False)) for _ in range(5))
    return {"config": "C1 bonsai 800x600, static camera (128,128,400), quality BEST", "unit": "Mrays/s",
            "device": round(W * H / mean / 1e3, 1), "kernel_ms": round(mean, 4), "kernel_ms_min": round(best, 4),
            "e2e_vox_render_redraw": round(W * H / e2e / 1e6, 1),
            "cpu": {"value": round(W * H / cpu / 1e6, 1), "threads": THREADS, "oracle": O.name, "sample": "the same frame, best of 5"}}


def modes():
    """C2 frames in the reference's other quality modes (ADAPTIVE is its default, renderer.c:78)."""
    W, H = 1920, 1080
    vox = ds.volume_voxels("skull")
    tree = vv.Tree.build(vox)
    dev = tree.device_handle()
    ctx = vv.Context(W, H)
    d_tex, d_pix = vv.DeviceBuffer.from_array(ctx.texture()), vv.DeviceBuffer(W * H * 4)
    cam = vv.Camera()
    O = cpu_oracle()
    ot = O.tree(vox)
    frames = [0, 90, 180, 270]
    out = {"config": "C2 frames 0/90/180/270 per quality mode (device-resident kernel; CPU = same mode, all threads)", "unit": "Mrays/s"}
    for name, q in (("best", 0), ("fast", 1), ("adaptive", 2), ("adaptive+merge", 6)):
        ms, cpu = [], []
        for f in frames:
            ds.setup_camera(cam, *ds.orbit_pose(f), 1.2, W, H)
            gc = cam.gpu_camera()
            fr = api.GpuFrame(W, H, 0, H, W, q, ctx.pixfmt(), None)
            run = lambda: g.voxgpu_render(dev, C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, None, None, None)
            run(); run()
            ms.append(timed(run, 9)[0])
            ocam = ds.setup_camera(O.camera(0), *ds.orbit_pose(f), 1.2, W, H)
            ot.render(ocam, W, H, q, threads=THREADS, want_hits=False)
            cpu.append(min(_t(lambda: ot.render(ocam, W, H, q, threads=THREADS, want_hits=False)) for _ in range(3)))
        out[name] = {"device": round(W * H / np.mean(ms) / 1e3, 1), "kernel_ms": round(float(np.mean(ms)), 4),
                     "cpu": round(W * H / np.mean(cpu) / 1e6, 1)}
    return out


def _t(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def c3():
    n = 1 << 24
    vox, used = ds.random_voxels()
    tree = vv.Tree.build(vox)
    dev = tree.device_handle()
    o, d = ds.random_rays(n, side=512, start_index=used)
    o4, d4 = api.dots4(o), api.dots4(d)
    d_o, d_d = vv.DeviceBuffer.from_array(o4), vv.DeviceBuffer.from_array(d4)
    d_leaf, d_pts, d_vox = vv.DeviceBuffer(n * 4), vv.DeviceBuffer(n * 16), vv.DeviceBuffer(n * 4)
    run = lambda: g.voxgpu_rays_device(dev, 0, d_o.ptr, d_d.ptr, n, d_leaf.ptr, d_pts.ptr, d_vox.ptr, None)
    # the batch as given (a warp = 32 unrelated rays) and in entry order (the default for batches this size;
    # the time includes the key kernel and the radix sort of the ray indices)
    g.voxgpu_set_rays_sort(0)
    for _ in range(3):
        run()
    plain_mean, plain_best = timed(run, 10)
    plain_leaf = d_leaf.download(np.int32, (n,))
    g.voxgpu_set_rays_sort(-1)
    for _ in range(3):
        run()
    mean, best = timed(run, 10)
    leaf_sorted = d_leaf.download(np.int32, (n,))
    assert np.array_equal(plain_leaf, leaf_sorted)
    hits = int((leaf_sorted >= 0).sum())
    gs = api.GpuStats()
    ns = 1 << 20
    assert g.voxgpu_rays_stats(dev, 0, d_o.ptr, d_d.ptr, ns, None, None, None, C.byref(gs)) == 0
    st = gs.as_dict()
    bpr = (st["inner"] * BYTES_INNER + st["leaf"] * BYTES_LEAF + st["dense"] * BYTES_DENSE + st["dots_tested"] * BYTES_DOT) / ns + 32 + 24
    leaf = np.empty(n, np.int32); pts = np.zeros((n, 4), np.float32); vx = np.empty(n, np.int32)
    host = lambda: api.libs.trees.vox_ray_tree_intersection_batch(tree.h, o4.ctypes.data, d4.ctypes.data, n, leaf.ctypes.data,
                                                                   pts.ctypes.data, vx.ctypes.data)
    host()
    e2e = min(_t(host) for _ in range(3))
    O = cpu_oracle()
    ot = O.tree(vox)
    m = 1 << 20
    ot.rays(o[:65536], d[:65536], threads=THREADS)
    cpu = min(_t(lambda: ot.rays(o[:m], d[:m], threads=THREADS)) for _ in range(3))
    return {"config": "C3 16 Mi random rays vs 10 063 815-voxel random tree (vox_ray_tree_intersection batch, k_rays)", "unit": "Mrays/s",
            "device": round(n / mean / 1e3, 1), "kernel_ms": round(mean, 3), "kernel_ms_min": round(best, 3), "hits": hits,
            "order": "entry order (Morton code of the root-box entry point; key kernel + 2-pass radix sort inside the timed region)",
            "as_given": {"device": round(n / plain_mean / 1e3, 1), "kernel_ms": round(plain_mean, 3),
                         "frac": round(bpr * n / plain_mean / 1e6 / peak(), 4)},
            "tree_device_bytes": int(g.voxgpu_tree_device_bytes(dev)),
            "e2e_host_buffers": round(n / e2e / 1e6, 1), "h2d_bytes": n * 32, "d2h_bytes": n * 24,
            "roofline": {"bound": "hbm", "bytes_per_ray": round(bpr, 1), "achieved_gbs": round(bpr * n / mean / 1e6, 1),
                         "peak_gbs": peak(), "frac": round(bpr * n / mean / 1e6 / peak(), 4)},
            "cpu": {"value": round(m / cpu / 1e6, 2), "threads": THREADS, "oracle": O.name, "sample": f"the first {m} rays, best of 3"}}


def c4():
    n = 1 << 22
    vox = ds.volume_voxels("skull")
    tree = vv.Tree.build(vox)
    dev = tree.device_handle()
    c, r = ds.random_balls(n, *tree.bounding_box())
    c4_, r_ = api.dots4(c), np.ascontiguousarray(r, np.float32)
    d_c, d_r, d_out = vv.DeviceBuffer.from_array(c4_), vv.DeviceBuffer.from_array(r_), vv.DeviceBuffer(n)
    run = lambda: g.voxgpu_balls_device(dev, d_c.ptr, d_r.ptr, n, d_out.ptr, None)
    for _ in range(3):
        run()
    mean, best = timed(run, 20)
    hits = int(d_out.download(np.uint8, (n,)).sum())
    out = np.empty(n, np.uint8)
    host = lambda: api.libs.trees.vox_tree_ball_collidep_batch(tree.h, c4_.ctypes.data, r_.ctypes.data, n, out.ctypes.data)
    host()
    e2e = min(_t(host) for _ in range(3))
    O = cpu_oracle()
    ot = O.tree(vox)
    m = 1 << 20
    ot.balls(c[:65536], r[:65536], threads=THREADS)
    cpu = min(_t(lambda: ot.balls(c[:m], r[:m], threads=THREADS)) for _ in range(3))
    return {"config": "C4 4 Mi random balls (radius 0.5..8) vs the skull tree (vox_tree_ball_collidep batch, k_balls)", "unit": "Mballs/s",
            "device": round(n / mean / 1e3, 1), "kernel_ms": round(mean, 4), "kernel_ms_min": round(best, 4), "collisions": hits,
            "e2e_host_buffers": round(n / e2e / 1e6, 1), "h2d_bytes": n * 20, "d2h_bytes": n,
            "cpu": {"value": round(m / cpu / 1e6, 2), "threads": THREADS, "oracle": O.name, "sample": f"the first {m} balls, best of 3"}}


if __name__ == "__main__":
    if not vv.gpu_available():
        raise SystemExit("configs_bench.py: no CUDA device")
    vv.set_device(0)
    which = [a.lower() for a in sys.argv[1:]] or ["c1", "c3", "c4"]
    for name in which:
        print(json.dumps({"c1": c1, "c3": c3, "c4": c4, "modes": modes}[name]()), flush=True)
