#!/usr/bin/env python
"""Minimal driver for ncu / A-B: 7680x4320 frames of the C5 volume (union of 4096 balls in 1024^3, ~106 M voxels,
2.6 GB device tree: streams from HBM) through K1.

    python tools/profile_c5.py [launches] [side] [balls] [interleave world] [interleave rank]
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxvision_b200 as vv
from voxvision_b200 import api
from tools import datasets as ds

launches = int(sys.argv[1]) if len(sys.argv) > 1 else 24
side = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
balls = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
W, H = 7680, 4320
g = api.libs.gpu
tree = vv.Tree.build(ds.ball_union_voxels(side, balls))
dev = tree.device_handle()
ctx = vv.Context(16, 4)
d_tex = vv.DeviceBuffer.from_array(ctx.texture())
buf = vv.DeviceBuffer(W * H * 4)
cam = vv.Camera()
cams = []
for k in range(4):
    ds.setup_camera(cam, *ds.orbit_pose(k * 90.0, 360, (side / 2.0,) * 3, 328.0 * side / 256.0), 1.2, W, H)
    cams.append(cam.gpu_camera())
il_world = int(sys.argv[4]) if len(sys.argv) > 4 else 0
il_rank = int(sys.argv[5]) if len(sys.argv) > 5 else 0
fr = api.GpuFrame(W, H, 0, H, W, 0, ctx.pixfmt(), None, il_world, il_rank)
t = vv.Timer()
ms = []
for i in range(launches):
    t.start()
    assert g.voxgpu_render(dev, C.byref(cams[i % 4]), C.byref(fr), d_tex.ptr, buf.ptr, None, None, None) == 0
    t.stop()
    ms.append(t.elapsed_ms())
print("C5 ms per 8K frame, launches", launches, ": first 16 (slots warming)", np.round(ms[:16], 3).tolist())
print("C5 warm: mean %.3f ms min %.3f ms -> %.0f Mrays/s" % (np.mean(ms[16:]), np.min(ms[16:]), W * H / np.mean(ms[16:]) / 1e3) if launches > 16 else "")
