#!/usr/bin/env python
"""Minimal driver for ncu: renders a few skull 1080p frames into device memory."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxvision_b200 as vv
from voxvision_b200 import api
from tools import datasets as ds

W, H = 1920, 1080
frames = [int(a) for a in sys.argv[1:]] or [0, 0, 0, 0, 90, 90]
tree = vv.Tree.build(ds.volume_voxels("skull"))
ctx = vv.Context(W, H)
d_tex = vv.DeviceBuffer.from_array(ctx.texture())
d_pix = vv.DeviceBuffer(W * H * 4)
cam = vv.Camera()
fr = api.GpuFrame(W, H, 0, H, W, 0, ctx.pixfmt())
t = vv.Timer()
for f in frames:
    ds.setup_camera(cam, *ds.orbit_pose(f), 1.2, W, H)
    gc = cam.gpu_camera()
    t.start(); rc = api.libs.gpu.voxgpu_render(tree.device_handle(), C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, None, None, None); t.stop()
    assert rc == 0
    print("frame", f, "ms", t.elapsed_ms())
