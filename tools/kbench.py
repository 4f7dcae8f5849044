#!/usr/bin/env python
"""Kernel A/B harness: for each variant, render the eight golden C2 frames,
check them against the golden digests, and print CUDA-event kernel times with
the frame going (a) to device memory and (b) to mapped pinned host memory.

    python tools/kbench.py [variant ...]      variants: see voxgpu_set_variant
"""
import ctypes as C
import json
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxvision_b200 as vv
from voxvision_b200 import api
from tools import datasets as ds

W, H = 1920, 1080
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
gold = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "golden.json")))
variants = [int(a) for a in sys.argv[1:]] or [1]
g = api.libs.gpu
g.voxgpu_set_variant.argtypes = [C.c_int]
tree = vv.Tree.build(ds.volume_voxels("skull"))
dev = tree.device_handle()
ctx = vv.Context(W, H)
d_tex = vv.DeviceBuffer.from_array(ctx.texture())
d_pix = vv.DeviceBuffer(W * H * 4)
d_leaf = vv.DeviceBuffer(W * H * 4)
d_pts = vv.DeviceBuffer(W * H * 16)
h_pix = g.voxgpu_host_alloc(W * H * 4)
m_pix = g.voxgpu_host_device_pointer(h_pix)
h_view = np.ctypeslib.as_array(C.cast(h_pix, C.POINTER(C.c_uint32)), (H, W))
cam = vv.Camera()
fr = api.GpuFrame(W, H, 0, H, W, 0, ctx.pixfmt())
t = vv.Timer()
frames = sorted(int(k) for k in gold["c2"])


sched = g.voxgpu_sched_create()


def timed(gc, target, reps=9, sc=None):
    ms = []
    for _ in range(reps):
        t.start(); g.voxgpu_render_sched(dev, C.byref(gc), C.byref(fr), d_tex.ptr, target, None, None, None, sc); t.stop()
        ms.append(t.elapsed_ms())
    return sorted(ms)


for v in variants:
    g.voxgpu_set_variant(v)
    tot_d = tot_m = tot_s = 0.0
    for f in frames:
        ds.setup_camera(cam, *ds.orbit_pose(f), 1.2, W, H)
        gc = cam.gpu_camera()
        assert g.voxgpu_render(dev, C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, d_leaf.ptr, d_pts.ptr, None) == 0
        leaf = d_leaf.download(np.int32, (H, W)); pts = d_pts.download(np.float32, (H, W, 4))[:, :, :3]
        pix = d_pix.download(np.uint32, (H, W))
        want = gold["c2"][str(f)]
        m = leaf >= 0
        ok = [sha(leaf) == want["leaf_sha"], sha(pts[m]) == want["hit_sha"], sha(pix) == want["pix_sha"]]
        md = timed(gc, d_pix.ptr)
        ok.append(sha(d_pix.download(np.uint32, (H, W))) == want["pix_sha"])
        # scheduled: as in an animation, every timed launch follows a launch of the frame one orbit step before
        ds.setup_camera(cam, *ds.orbit_pose(f - 1), 1.2, W, H)
        gprev = cam.gpu_camera()
        ms_ = []
        for _ in range(9):
            g.voxgpu_render_sched(dev, C.byref(gprev), C.byref(fr), d_tex.ptr, d_pix.ptr, None, None, None, sched)
            ms_ += timed(gc, d_pix.ptr, reps=1, sc=sched)
        ms_.sort()
        ok.append(sha(d_pix.download(np.uint32, (H, W))) == want["pix_sha"])
        tot_s += ms_[len(ms_) // 2]
        h_view[:] = 0xdeadbeef
        mm = timed(gc, m_pix)
        g.voxgpu_synchronize(None)
        ok.append(sha(h_view) == want["pix_sha"])
        tot_d += md[len(md) // 2]; tot_m += mm[len(mm) // 2]
        print(f"variant {v} frame {f:3d} parity {'OK' if all(ok) else ok} hits {int(m.sum())} "
              f"device min {md[0]:.4f} med {md[len(md)//2]:.4f} | scheduled min {ms_[0]:.4f} med {ms_[len(ms_)//2]:.4f} | mapped min {mm[0]:.4f} med {mm[len(mm)//2]:.4f} ms", flush=True)
    n = len(frames)
    print(f"variant {v}: mean of per-frame medians: device {tot_d/n:.4f} ms -> {W*H/(tot_d/n)/1e3:.0f} Mrays/s | "
          f"scheduled {tot_s/n:.4f} ms -> {W*H/(tot_s/n)/1e3:.0f} Mrays/s | "
          f"mapped {tot_m/n:.4f} ms -> {W*H/(tot_m/n)/1e3:.0f} Mrays/s", flush=True)
