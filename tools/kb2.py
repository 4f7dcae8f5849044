#!/usr/bin/env python
"""Quick A/B of render kernel variants on the golden C2 frames: parity digest + CUDA-event times.
    python tools/kb2.py v[:refill_min] ...   e.g.  1 3:1 3:8 3:16 3:32"""
import ctypes as C, json, hashlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxvision_b200 as vv
from voxvision_b200 import api
from tools import datasets as ds

W, H = 1920, 1080
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
gold = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "golden.json")))
g = api.libs.gpu
g.voxgpu_set_variant.argtypes = [C.c_int]
g.voxgpu_set_refill_min.argtypes = [C.c_int]
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
for spec in sys.argv[1:] or ["1"]:
    v, _, rm = spec.partition(":")
    g.voxgpu_set_variant(int(v))
    if rm: g.voxgpu_set_refill_min(int(rm))
    tot = 0.0; allok = True
    per = []
    for f in frames:
        ds.setup_camera(cam, *ds.orbit_pose(f), 1.2, W, H)
        gc = cam.gpu_camera()
        assert g.voxgpu_render(dev, C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, d_leaf.ptr, d_pts.ptr, None) == 0
        leaf = d_leaf.download(np.int32, (H, W)); pts = d_pts.download(np.float32, (H, W, 4))[:, :, :3]
        pix = d_pix.download(np.uint32, (H, W))
        want = gold["c2"][str(f)]
        m = leaf >= 0
        ok = sha(leaf) == want["leaf_sha"] and sha(pts[m]) == want["hit_sha"] and sha(pix) == want["pix_sha"]
        ms = []
        for _ in range(9):
            t.start(); g.voxgpu_render(dev, C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, None, None, None); t.stop()
            ms.append(t.elapsed_ms())
        ok = ok and sha(d_pix.download(np.uint32, (H, W))) == want["pix_sha"]
        allok &= ok
        ms.sort(); tot += ms[4]; per.append(ms[4])
    n = len(frames)
    print(f"variant {spec}: parity {'OK' if allok else 'FAIL'}; median ms/frame mean {tot/n:.4f} -> {W*H/(tot/n)/1e3:.0f} Mrays/s; per-frame " + " ".join(f"{x:.3f}" for x in per), flush=True)
