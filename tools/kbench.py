#!/usr/bin/env python
"""Kernel A/B harness: for each variant, render C2 frames into device memory,
check them against the golden digests, and print CUDA-event kernel times."""
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
variants = [int(a) for a in sys.argv[1:]] or [0, 1]
g = api.libs.gpu
g.voxgpu_set_variant.argtypes = [C.c_int]
tree = vv.Tree.build(ds.volume_voxels("skull"))
dev = tree.device_handle()
ctx = vv.Context(W, H)
d_tex = vv.DeviceBuffer.from_array(ctx.texture())
d_pix = vv.DeviceBuffer(W * H * 4)
d_leaf = vv.DeviceBuffer(W * H * 4)
d_pts = vv.DeviceBuffer(W * H * 16)
cam = vv.Camera()
fr = api.GpuFrame(W, H, 0, H, W, 0, ctx.pixfmt())
t = vv.Timer()
frames = sorted(int(k) for k in gold["c2"])
for v in variants:
    g.voxgpu_set_variant(v)
    total = 0.0
    for f in frames:
        ds.setup_camera(cam, *ds.orbit_pose(f), 1.2, W, H)
        gc = cam.gpu_camera()
        assert g.voxgpu_render(dev, C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, d_leaf.ptr, d_pts.ptr, None) == 0
        leaf = d_leaf.download(np.int32, (H, W)); pts = d_pts.download(np.float32, (H, W, 4))[:, :, :3]
        pix = d_pix.download(np.uint32, (H, W))
        want = gold["c2"][str(f)]
        m = leaf >= 0
        ok = (sha(leaf) == want["leaf_sha"], sha(pts[m]) == want["hit_sha"], sha(pix) == want["pix_sha"])
        ms = []
        for _ in range(7):
            t.start(); g.voxgpu_render(dev, C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, None, None, None); t.stop()
            ms.append(t.elapsed_ms())
        pix2 = d_pix.download(np.uint32, (H, W))
        ok += (sha(pix2) == want["pix_sha"],)
        total += min(ms)
        print(f"variant {v} frame {f:3d} parity {ok} hits {int(m.sum())} ms min {min(ms):.4f} med {sorted(ms)[3]:.4f}", flush=True)
    print(f"variant {v}: mean of per-frame minima {total/len(frames):.4f} ms -> {W*H/(total/len(frames))/1e3:.0f} Mrays/s", flush=True)
