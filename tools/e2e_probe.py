#!/usr/bin/env python
"""Where does an end-to-end frame go?  wall time per vox_render+vox_redraw vs the presenter's device time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxvision_b200 as vv
from tools import datasets as ds
W, H, FOV = 1920, 1080, 1.2
tree = vv.Tree.build(ds.volume_voxels("skull"))
ctx = vv.Context(W, H); ctx.set_scene(tree); ctx.set_quality(vv.QUALITY_BEST)
cam = vv.Camera("simple-camera")
poses = [ds.orbit_pose(f) for f in range(360)]
ds.setup_camera(cam, *poses[0], FOV, W, H); ctx.set_camera(cam)
def loop(setcam=True, n=360):
    dev = []
    t0 = time.perf_counter()
    tr = trd = 0.0
    for f in range(n):
        if setcam:
            pos, look = poses[f]
            cam.set_position(pos); cam.look_at(look); ctx.set_camera(cam)
        a = time.perf_counter(); ctx.render(); b = time.perf_counter(); ctx.redraw(); c = time.perf_counter()
        tr += b - a; trd += c - b
        dev.append(ctx.kernel_ms())
    wall = (time.perf_counter() - t0) / n * 1e3
    return wall, tr / n * 1e3, trd / n * 1e3, float(np.mean(dev))
loop(); loop()
for sc in (True, False):
    w, r, rd, d = loop(sc)
    print(f"setcam={sc}: wall {w:.4f} ms/frame (render call {r:.4f}, redraw call {rd:.4f}), presenter device time {d:.4f} ms")
