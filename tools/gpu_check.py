#!/usr/bin/env python
"""Quick on-GPU parity + timing check (development aid, prints details)."""
import ctypes as C
import faulthandler
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxvision_b200 as vv
from voxvision_b200 import api
from oracle import orc
from tools import datasets as ds


def render_device(tree, cam, w, h, tex, fmt, want_hits=True, iters=1, stream=None):
    g = api.libs.gpu
    dev = tree.device_handle()
    d_tex = vv.DeviceBuffer.from_array(tex)
    d_pix = vv.DeviceBuffer(w * h * 4)
    d_leaf = vv.DeviceBuffer(w * h * 4) if want_hits else None
    d_pts = vv.DeviceBuffer(w * h * 16) if want_hits else None
    gc = cam.gpu_camera()
    fr = api.GpuFrame(w, h, 0, h, w, 0, fmt)
    t = vv.Timer()
    ms = []
    for _ in range(iters):
        t.start(stream)
        rc = g.voxgpu_render(dev, C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr,
                             d_leaf.ptr if want_hits else None, d_pts.ptr if want_hits else None, stream)
        t.stop(stream)
        if rc:
            raise vv.VoxError(g.voxgpu_last_error().decode())
        ms.append(t.elapsed_ms())
    pix = d_pix.download(np.uint32, (h, w))
    if want_hits:
        return pix, d_leaf.download(np.int32, (h, w)), d_pts.download(np.float32, (h, w, 4))[:, :, :3], ms
    return pix, None, None, ms


def main():
    faulthandler.dump_traceback_later(int(os.environ.get("CHECK_TIMEOUT", "120")), exit=True)
    print("devices:", vv.device_count())
    P = orc.Oracle("ref" if orc.available("ref") else "port")
    print("oracle:", P.name)
    W, H = 1920, 1080
    vox = ds.volume_voxels("skull")
    t0 = time.time(); T = vv.Tree.build(vox); print("build %.3fs" % (time.time() - t0))
    TP = P.tree(vox)
    ti = T.info(); print({k: getattr(ti, k) for k, _ in ti._fields_})

    ctx = vv.Context(W, H)
    tex = ctx.texture()
    for frame in (0, 90):
        pos, look = ds.orbit_pose(frame)
        cam = ds.setup_camera(vv.Camera("simple-camera"), pos, look, 1.2, W, H)
        ocam = ds.setup_camera(P.camera(0), pos, look, 1.2, W, H)
        assert np.array_equal(cam.state()[:10].view(np.uint32), ocam.state()[:10].view(np.uint32))
        t0 = time.time(); opix, oleaf, ohit = TP.render(ocam, W, H, 0, tex, threads=os.cpu_count()); t_cpu = time.time() - t0
        pix, leaf, pts, ms = render_device(T, cam, W, H, tex, ctx.pixfmt(), True, 5)
        hitmask = oleaf >= 0
        print(f"frame {frame}: oracle hits {hitmask.sum()} gpu hits {(leaf >= 0).sum()} cpu {t_cpu*1e3:.1f} ms gpu ms {ms}")
        print("  leaf equal:", np.array_equal(leaf, oleaf), " mismatches:", int((leaf != oleaf).sum()))
        same_pts = np.array_equal(pts[hitmask].view(np.uint32), ohit[hitmask].view(np.uint32))
        print("  hit points bit-equal:", same_pts, " pixels equal:", np.array_equal(pix, opix))
        if not same_pts:
            bad = np.argwhere((pts.view(np.uint32) != ohit.view(np.uint32)).any(axis=2) & hitmask)
            print("   first bad:", bad[:5], pts[tuple(bad[0])], ohit[tuple(bad[0])])
        # through the reference-facing API
        ctx.set_scene(T); ctx.set_camera(cam); ctx.set_quality(0)
        t0 = time.time(); ctx.render(); ctx.redraw(); t_api = time.time() - t0
        print("  vox_render+vox_redraw: %.3f ms wall, kernel %.3f ms, equal: %s, err=%r" %
              (t_api * 1e3, ctx.kernel_ms(), np.array_equal(ctx.pixels(), opix), vv.last_error()))
        for _ in range(3):
            t0 = time.time(); ctx.render(); ctx.redraw(); t_api = time.time() - t0
            print("    again: %.3f ms wall, kernel %.3f ms" % (t_api * 1e3, ctx.kernel_ms()))
        pix2, _, _, ms2 = render_device(T, cam, W, H, tex, ctx.pixfmt(), False, 20)
        print("  no-hits kernel ms: min %.4f median %.4f -> %.1f Mrays/s" %
              (min(ms2), sorted(ms2)[len(ms2) // 2], W * H / min(ms2) / 1e3))

    # rays
    n = 1 << 20
    o, d = ds.random_rays(n, side=256)
    t0 = time.time(); ol, oh, od = TP.rays(o, d, threads=os.cpu_count()); t_cpu = time.time() - t0
    t0 = time.time(); gl, gh, gv = T.rays(o, d); t_gpu = time.time() - t0
    m = ol >= 0
    print(f"rays: oracle hits {m.sum()} cpu {t_cpu:.3f}s gpu(host api) {t_gpu:.3f}s leaf eq {np.array_equal(ol, gl)} "
          f"pts eq {np.array_equal(oh[m].view(np.uint32), gh[m].view(np.uint32))} "
          f"vox eq {np.array_equal(od[m], gv[m].astype(np.int64))}")
    # single-call API
    for i in range(5):
        lid, p = T.ray(o[i], d[i])
        print("  single:", lid, p, ol[i], oh[i] if ol[i] >= 0 else None)
    # balls
    bmin, bmax = T.bounding_box()
    c, r = ds.random_balls(1 << 20, bmin, bmax)
    t0 = time.time(); ob = TP.balls(c, r, threads=os.cpu_count()); t_cpu = time.time() - t0
    t0 = time.time(); gb = T.balls(c, r); t_gpu = time.time() - t0
    print(f"balls: oracle {ob.sum()} gpu {gb.sum()} equal {np.array_equal(ob, gb)} cpu {t_cpu:.3f}s gpu(host api) {t_gpu:.3f}s")
    print("single ball:", T.ball(c[0], float(r[0])), ob[0])
    print("launches:", vv.launch_count())


if __name__ == "__main__":
    main()
