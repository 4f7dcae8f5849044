#!/usr/bin/env python
"""Per-tile timeline of the render kernel (development aid): how many warps are
busy over the kernel's lifetime, how long the tiles take, how long the tail is."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxvision_b200 as vv
from voxvision_b200 import api
from tools import datasets as ds

W, H = 1920, 1080
g = api.libs.gpu
g.voxgpu_set_trace.argtypes = [C.c_void_p]
g.voxgpu_set_variant.argtypes = [C.c_int]
tree = vv.Tree.build(ds.volume_voxels("skull"))
dev = tree.device_handle()
ctx = vv.Context(W, H)
d_tex = vv.DeviceBuffer.from_array(ctx.texture())
d_pix = vv.DeviceBuffer(W * H * 4)
ntiles = (W // 8) * (H // 4)
d_tr = vv.DeviceBuffer(ntiles * 24)
cam = vv.Camera()
fr = api.GpuFrame(W, H, 0, H, W, 0, ctx.pixfmt())
for variant in [int(a) for a in sys.argv[1:]] or [1]:
    g.voxgpu_set_variant(variant)
    for f in (0, 90):
        ds.setup_camera(cam, *ds.orbit_pose(f), 1.2, W, H)
        gc = cam.gpu_camera()
        sched = g.voxgpu_sched_create() if os.environ.get("TRACE_SCHED") else None
        # with a schedule: warm it up on the frame 3 steps earlier on the orbit, like an animation would
        ds.setup_camera(cam, *ds.orbit_pose(f - 1), 1.2, W, H)
        gp = cam.gpu_camera()
        for _ in range(3):
            g.voxgpu_set_trace(None)
            g.voxgpu_render_sched(dev, C.byref(gp), C.byref(fr), d_tex.ptr, d_pix.ptr, None, None, None, sched)
        g.voxgpu_set_trace(d_tr.ptr)
        g.voxgpu_render_sched(dev, C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, None, None, None, sched)
        g.voxgpu_set_trace(None)
        tr = d_tr.download(np.uint64, (ntiles, 3))
        t0, t1 = tr[:, 0].astype(np.int64), tr[:, 1].astype(np.int64)
        base = t0.min()
        t0 -= base; t1 -= base
        dur = (t1 - t0) / 1e3
        total = t1.max() / 1e3
        print(f"variant {variant} frame {f}: kernel span {total:.1f} us, tiles {ntiles}, tile us: median {np.median(dur):.2f} "
              f"p90 {np.percentile(dur, 90):.1f} p99 {np.percentile(dur, 99):.1f} max {dur.max():.1f}; sum {dur.sum()/1e3:.1f} ms")
        # busy warps over time
        edges = np.linspace(0, t1.max(), 41)
        busy = []
        for a, b in zip(edges[:-1], edges[1:]):
            ov = np.clip(np.minimum(t1, b) - np.maximum(t0, a), 0, None).sum() / (b - a)
            busy.append(ov)
        print("   busy warps per 2.5% slice:", " ".join(f"{x:.0f}" for x in busy))
        it = (tr[:, 2] & np.uint64(0xffffffff)).astype(np.int64)
        print("   iterations per tile: median %d p90 %d p95 %d p99 %d max %d; tiles >= 24: %d, >= 40: %d, >= 64: %d, >= 100: %d" % (
            np.median(it), np.percentile(it, 90), np.percentile(it, 95), np.percentile(it, 99), it.max(),
            (it >= 24).sum(), (it >= 40).sum(), (it >= 64).sum(), (it >= 100).sum()))
        late = np.argsort(t1)[-6:]
        for i in late:
            print(f"   late tile {i} (x {i % (W//8)}, y {i // (W//8)}): start {t0[i]/1e3:.1f} end {t1[i]/1e3:.1f} us dur {dur[i]:.1f} iters {it[i]}")
        for lo, hi in ((1, 2), (2, 8), (8, 16), (16, 24), (24, 32), (32, 40), (40, 64), (64, 1000)):
            sel = (it >= lo) & (it < hi)
            if sel.any():
                print(f"   iters [{lo},{hi}): {int(sel.sum()):6d} tiles, dur us median {np.median(dur[sel]):7.1f} p90 {np.percentile(dur[sel], 90):7.1f} max {dur[sel].max():7.1f}, "
                      f"median start {np.median(t0[sel])/1e3:6.1f} us, share of warp time {100*dur[sel].sum()/dur.sum():4.1f}%")
