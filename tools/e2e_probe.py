#!/usr/bin/env python
"""One-context end-to-end time per C2 frame (vox_render + vox_redraw from C, examples/orbit_bench.c) for several band
counts of the presenter ($VOXGPU_BANDS is read when a context is first rendered), in BEST and ADAPTIVE.

    python tools/e2e_probe.py [bands ...]         default: 1 2 3 4 6 8; 0 = the presenter's own choice
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxvision_b200 as vv
from voxvision_b200 import api
from tools import datasets as ds
import bench

W, H, FRAMES = 1920, 1080, 360
DEVICES = int(os.environ.get("PROBE_DEVICES", "1"))  # > 1: every frame split over that many GPUs of this process
STEPS = int(os.environ.get("PROBE_STEPS", "2"))      # orbits per leg ($PROBE_STEPS > 2: BEST only, for runs side by side)
orbit = C.CDLL(os.path.join(api.LIB_DIR, "liborbitbench.so"))
orbit.orbit_bench_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_int,
                                  C.c_uint, C.c_int, C.POINTER(bench.OrbitBenchResult)]
vv.set_device(0)
tree = vv.Tree.build(ds.volume_voxels("skull"))
poses = np.ascontiguousarray([list(p) + list(l) for p, l in (ds.orbit_pose(f, FRAMES) for f in range(FRAMES))], np.float32)
for bands in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 6, 8]:
    if bands > 0:
        os.environ["VOXGPU_BANDS"] = str(bands)
    else:
        os.environ.pop("VOXGPU_BANDS", None)
    line = [f"bands {bands}:"]
    for qname, q in (("BEST", 0), ("ADAPTIVE", 2))[:1 if STEPS != 2 or DEVICES > 1 else 2]:
        for contexts in (1, 6):
            res = bench.OrbitBenchResult()
            rc = orbit.orbit_bench_run(tree.h, W, H, FRAMES, poses.ctypes.data, 1.2, contexts, DEVICES, q, STEPS, C.byref(res))
            assert rc == 0, res.error
            wall = res.seconds / res.frames * 1e3
            line.append(f"{qname} x{contexts}: wall {wall:.4f} ms/frame ({W*H/wall/1e3:.0f} Mrays/s)" +
                        (f" device {res.device_ms_mean:.4f}" if contexts == 1 else ""))
    print(" | ".join(line), flush=True)
