#!/usr/bin/env python
"""Minimal driver for ncu: the C3 batch (16 Mi random rays vs the ~10 M-voxel random tree) through K2.

    python tools/profile_rays.py [sort mode: 0 as given, 1 entry order] [launches] [log2 of the batch]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxvision_b200 as vv
from voxvision_b200 import api
from tools import datasets as ds

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 24)
g = api.libs.gpu
vox, used = ds.random_voxels()
tree = vv.Tree.build(vox)
dev = tree.device_handle()
o, d = ds.random_rays(n, side=512, start_index=used)
d_o, d_d = vv.DeviceBuffer.from_array(api.dots4(o)), vv.DeviceBuffer.from_array(api.dots4(d))
d_leaf, d_pts, d_vox = vv.DeviceBuffer(n * 4), vv.DeviceBuffer(n * 16), vv.DeviceBuffer(n * 4)
g.voxgpu_set_rays_sort(mode)
t = vv.Timer()
for i in range(launches):
    t.start()
    assert g.voxgpu_rays_device(dev, 0, d_o.ptr, d_d.ptr, n, d_leaf.ptr, d_pts.ptr, d_vox.ptr, None) == 0
    t.stop()
    print("mode", mode, "launch", i, "ms", t.elapsed_ms(), flush=True)
print("hits", int((d_leaf.download(np.int32, (n,)) >= 0).sum()))
