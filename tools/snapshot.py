#!/usr/bin/env python
"""Render a skull frame (unlit, and lit by the scripted lights) on the GPU and save it with
vox_screenshot into gpurun_out/ -- something to look at.   python tools/snapshot.py [frame]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import voxvision_b200 as vv  # noqa: E402
from tools import datasets as ds  # noqa: E402

frame = int(sys.argv[1]) if len(sys.argv) > 1 else 45
out = os.path.join(ROOT, "gpurun_out", "shots")
os.makedirs(out, exist_ok=True)
R = vv.libs.rnd
R.vox_screenshot.argtypes = [C.c_void_p, C.c_char_p]
tree = vv.Tree.build(ds.volume_voxels("skull"))
w, h = 960, 540
cam = ds.setup_camera(vv.Camera(), *ds.orbit_pose(frame), 1.2, w, h)
ctx = vv.Context(w, h)
ctx.set_scene(tree); ctx.set_camera(cam); ctx.set_quality(vv.QUALITY_BEST)
ctx.render(); ctx.redraw()
assert R.vox_screenshot(ctx.h, out.encode()) == 1, vv.last_error()
lm = vv.LightManager()
ds.apply_lights(lm, ds.light_script())
ctx.set_light_manager(lm)
ctx.render(); ctx.redraw()
assert R.vox_screenshot(ctx.h, out.encode()) == 1, vv.last_error()
print("saved", sorted(os.listdir(out)))
