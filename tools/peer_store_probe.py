#!/usr/bin/env python
"""NVLink evidence for the fused gather (needs 2 GPUs, ONE process so that ncu can watch it): GPU 1 renders
C2 frames with its pixel stores going straight into a frame buffer in GPU 0's memory (peer access), which is
exactly what the ranks of bench.py / the C5 leg do through a CUDA-IPC mapping of the same kind of buffer.

    python tools/peer_store_probe.py [frames]
    ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum -k regex:k_render python tools/peer_store_probe.py 8

Prints the frame rate and the achieved GB/s over the link for local and for peer stores, and checks that the
frame that arrived in GPU 0's memory is the frame GPU 0 renders itself."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxvision_b200 as vv
from voxvision_b200 import api
from tools import datasets as ds

W, H = 1920, 1080
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 120
g = api.libs.gpu
if vv.device_count() < 2:
    raise SystemExit("peer_store_probe.py needs two CUDA devices")
tree = vv.Tree.build(ds.volume_voxels("skull"))
ctx = vv.Context(W, H)
tex, fmt = ctx.texture(), ctx.pixfmt()
cam = vv.Camera()
gcams = []
for f in range(frames):
    ds.setup_camera(cam, *ds.orbit_pose(f * 360 // frames), 1.2, W, H)
    gcams.append(cam.gpu_camera())
fr = api.GpuFrame(W, H, 0, H, W, 0, fmt)

vv.set_device(0)
dev0 = tree.device_handle()
tex0 = vv.DeviceBuffer.from_array(tex)
fb0 = [vv.DeviceBuffer(W * H * 4) for _ in range(4)]          # the gather target, in GPU 0's memory
ref = vv.DeviceBuffer(W * H * 4)
vv.set_device(1)
assert g.voxgpu_enable_peer(0) == 0, g.voxgpu_last_error()
dev1 = tree.device_handle()
tex1 = vv.DeviceBuffer.from_array(tex)
fb1 = [vv.DeviceBuffer(W * H * 4) for _ in range(4)]
streams = [C.c_void_p(g.voxgpu_stream_create()) for _ in range(4)]
t = vv.Timer()
out = {}
for name, targets in (("local", fb1), ("peer", fb0)):
    for rep in range(2):
        for st in streams[1:]:
            g.voxgpu_stream_join(st, streams[0])
        t.start(streams[0])
        for f in range(frames):
            assert g.voxgpu_render(dev1, C.byref(gcams[f]), C.byref(fr), tex1.ptr, targets[f % 4].ptr, None, None, streams[f % 4]) == 0
        for st in streams[1:]:
            g.voxgpu_stream_join(streams[0], st)
        t.stop(streams[0])
        ms = t.elapsed_ms()
    out[name] = {"ms_per_frame": round(ms / frames, 4), "mrays_per_s": round(frames * W * H / ms / 1e3, 1),
                 "gbs_pixels": round(frames * W * H * 4 / ms / 1e6, 2)}
g.voxgpu_synchronize(None)
# the last frame each of the four targets received, against GPU 0's own rendering of it
vv.set_device(0)
ok = True
for f in range(frames - 4, frames):
    assert g.voxgpu_render(dev0, C.byref(gcams[f]), C.byref(fr), tex0.ptr, ref.ptr, None, None, None) == 0
    ok &= bool(np.array_equal(ref.download(np.uint32, (H, W)), fb0[f % 4].download(np.uint32, (H, W))))
out["frames_in_gpu0_equal_gpu0s_own"] = ok
out["what"] = "GPU 1 renders, pixels stored into GPU 0's memory over NVLink by the render kernel itself (4 frames in flight)"
print(json.dumps(out))
