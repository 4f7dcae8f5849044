#!/usr/bin/env python
"""Kernel A/B harness: render the eight golden C2 frames, check them against the golden digests, and print
CUDA-event times of LONE launches (the figure bench.py's roofline is computed from) for each tile cost feedback
threshold given (voxgpu_set_tile_feedback; 0 = feedback off).

    python tools/kbench.py [threshold ...]           default: the library's
    KBENCH_CULL=0|1|2                                voxgpu_set_cull for all runs (default 1)
"""
import ctypes as C
import hashlib
import json
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
thresholds = [int(a) for a in sys.argv[1:]] or [-1]
g = api.libs.gpu
g.voxgpu_set_cull(int(os.environ.get("KBENCH_CULL", "1")))
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


def timed(gc, target, reps=15):
    ms = []
    for _ in range(reps):
        t.start(); g.voxgpu_render(dev, C.byref(gc), C.byref(fr), d_tex.ptr, target, None, None, None); t.stop()
        ms.append(t.elapsed_ms())
    return sorted(ms)


for thr in thresholds:
    if thr >= 0:
        g.voxgpu_set_tile_feedback(thr)
    tot = 0.0
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
        tot += md[len(md) // 2]
        print(f"feedback {thr:3d} frame {f:3d} parity {'OK' if all(ok) else ok} hits {int(m.sum())} "
              f"lone launch min {md[0]:.4f} med {md[len(md)//2]:.4f} max {md[-1]:.4f} ms", flush=True)
    n = len(frames)
    print(f"feedback {thr}: mean of per-frame medians {tot/n:.4f} ms -> {W*H/(tot/n)/1e3:.0f} Mrays/s", flush=True)
