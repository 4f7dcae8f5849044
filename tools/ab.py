#!/usr/bin/env python
"""A/B of the installed libvoxgpu.so (tools/ab_build.sh swaps build-time variants in): golden parity of the
eight C2 frames, lone-launch time, throughput with 4 frames in flight, and optionally C3 (K2) and C5 (8K frame of
the 100 M-voxel volume, HBM-bound).  Render variants: `V[:F]` = voxgpu_set_variant (V) (1 = default) with the
tile cost feedback threshold F (voxgpu_set_tile_feedback; 0 = off; default: the library's).

    python tools/ab.py [--c3] [--c5] [variant ...]        default variants: 1
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

args = sys.argv[1:]
want_c3, want_c5 = "--c3" in args, "--c5" in args
specs = [a for a in args if not a.startswith("--")] or ["1"]
W, H = 1920, 1080
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
gold = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "golden.json")))
g = api.libs.gpu
tree = vv.Tree.build(ds.volume_voxels("skull"))
dev = tree.device_handle()
ctx = vv.Context(W, H)
fmt = ctx.pixfmt()
d_tex = vv.DeviceBuffer.from_array(ctx.texture())
d_pix = [vv.DeviceBuffer(W * H * 4) for _ in range(4)]
d_leaf, d_pts = vv.DeviceBuffer(W * H * 4), vv.DeviceBuffer(W * H * 16)
cam = vv.Camera()
fr = api.GpuFrame(W, H, 0, H, W, 0, fmt)
t = vv.Timer()
frames = sorted(int(k) for k in gold["c2"])
orbit = []
for f in range(0, 360, 2):
    ds.setup_camera(cam, *ds.orbit_pose(f), 1.2, W, H)
    orbit.append(cam.gpu_camera())
streams = [C.c_void_p(g.voxgpu_stream_create()) for _ in range(4)]
tag = os.environ.get("AB_TAG", "installed")

for spec in specs:
    v, _, fbk = spec.partition(":")
    g.voxgpu_set_variant(int(v))
    if fbk:
        g.voxgpu_set_tile_feedback(int(fbk))
    ok, lone = 0, []
    for f in frames:
        ds.setup_camera(cam, *ds.orbit_pose(f), 1.2, W, H)
        gc = cam.gpu_camera()
        assert g.voxgpu_render(dev, C.byref(gc), C.byref(fr), d_tex.ptr, d_pix[0].ptr, d_leaf.ptr, d_pts.ptr, None) == 0
        leaf = d_leaf.download(np.int32, (H, W)); pts = d_pts.download(np.float32, (H, W, 4))[:, :, :3]
        want = gold["c2"][str(f)]
        good = sha(leaf) == want["leaf_sha"] and sha(pts[leaf >= 0]) == want["hit_sha"] and sha(d_pix[0].download(np.uint32, (H, W))) == want["pix_sha"]
        ms = []
        for _ in range(11):
            t.start(); g.voxgpu_render(dev, C.byref(gc), C.byref(fr), d_tex.ptr, d_pix[0].ptr, None, None, None); t.stop()
            ms.append(t.elapsed_ms())
        good = good and sha(d_pix[0].download(np.uint32, (H, W))) == want["pix_sha"]
        ok += good
        lone.append(sorted(ms)[5])
    # lone launches along the orbit, 4 degrees apart, each frame once (what bench.py's roofline times)
    seq = []
    for rep in range(2):
        seq = []
        for gc in orbit[::2]:
            t.start(); g.voxgpu_render(dev, C.byref(gc), C.byref(fr), d_tex.ptr, d_pix[0].ptr, None, None, None); t.stop()
            seq.append(t.elapsed_ms())
    flight = []
    for rep in range(3):
        for st in streams[1:]:
            g.voxgpu_stream_join(st, streams[0])
        t.start(streams[0])
        for i, gc in enumerate(orbit):
            g.voxgpu_render(dev, C.byref(gc), C.byref(fr), d_tex.ptr, d_pix[i % 4].ptr, None, None, streams[i % 4])
        for st in streams[1:]:
            g.voxgpu_stream_join(streams[0], st)
        t.stop(streams[0])
        flight.append(t.elapsed_ms() / len(orbit))
    print(f"AB {tag} render {spec}: parity {ok}/8 | lone launch, same frame again {np.mean(lone):.4f} ms | along the orbit {np.mean(seq):.4f} ms ({W*H/np.mean(seq)/1e3:.0f} Mrays/s) | "
          f"4 in flight {min(flight):.4f} ms/frame ({W*H/min(flight)/1e3:.0f} Mrays/s)", flush=True)
g.voxgpu_set_variant(1)

if want_c3:
    n = 1 << 24
    vox, used = ds.random_voxels()
    t3 = vv.Tree.build(vox)
    o, d = ds.random_rays(n, side=512, start_index=used)
    d_o, d_d = vv.DeviceBuffer.from_array(api.dots4(o)), vv.DeviceBuffer.from_array(api.dots4(d))
    dl, dp, dv = vv.DeviceBuffer(n * 4), vv.DeviceBuffer(n * 16), vv.DeviceBuffer(n * 4)
    res = []
    for mode, refill in ((0, 0), (1, 0), (1, 8)):
        g.voxgpu_set_rays_sort(mode); g.voxgpu_set_rays_refill(refill)
        ms = []
        for _ in range(5):
            t.start(); g.voxgpu_rays_device(t3.device_handle(), 0, d_o.ptr, d_d.ptr, n, dl.ptr, dp.ptr, dv.ptr, None); t.stop()
            ms.append(t.elapsed_ms())
        res.append(f"sort {mode} refill {refill}: {min(ms):.3f} ms")
    print(f"AB {tag} C3: " + " | ".join(res), flush=True)
    del t3, d_o, d_d, dl, dp, dv

if want_c5:
    t5 = vv.Tree.build(ds.ball_union_voxels(1024, 4096))
    W5, H5 = 7680, 4320
    buf = vv.DeviceBuffer(W5 * H5 * 4)
    cams = []
    for k in range(4):
        ds.setup_camera(cam, *ds.orbit_pose(k * 90.0, 360, (512.0,) * 3, 328.0 * 4), 1.2, W5, H5)
        cams.append(cam.gpu_camera())
    fr5 = api.GpuFrame(W5, H5, 0, H5, W5, 0, fmt)
    ms = []
    for rep in range(8):            # the first rounds touch every render slot once (first use allocates its mask)
        t.start()
        for gc in cams:
            assert g.voxgpu_render(t5.device_handle(), C.byref(gc), C.byref(fr5), d_tex.ptr, buf.ptr, None, None, None) == 0
        t.stop()
        ms.append(t.elapsed_ms() / len(cams))
    print(f"AB {tag} C5: {min(ms):.3f} ms per 8K frame ({W5*H5/min(ms)/1e3:.0f} Mrays/s), digest {sha(buf.download(np.uint32, (H5, W5)))[:12]}", flush=True)
