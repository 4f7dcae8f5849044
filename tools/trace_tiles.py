#!/usr/bin/env python
"""Per-tile timeline of the render kernel (development aid): how many warps are busy over the kernel's
lifetime, how long the tiles take, how long the tail is -- and how well the number of cut boxes that
project onto a tile (what k_mark counts, the heavy-tile criterion) predicts the tile's cost.

    python tools/trace_tiles.py [heavy threshold ...]
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxvision_b200 as vv
from voxvision_b200 import api
from tools import datasets as ds

W, H = 1920, 1080
g = api.libs.gpu
tree = vv.Tree.build(ds.volume_voxels("skull"))
dev = tree.device_handle()
flat = tree.export()
ctx = vv.Context(W, H)
d_tex = vv.DeviceBuffer.from_array(ctx.texture())
d_pix = vv.DeviceBuffer(W * H * 4)
TX, TY = W // 8, H // 4
ntiles = TX * TY
d_tr = vv.DeviceBuffer(ntiles * 24)
cam = vv.Camera()
fr = api.GpuFrame(W, H, 0, H, W, 0, ctx.pixfmt())


def box_counts(gc):
    """What k_mark counts, on the host (double instead of fp32: a tile at a rectangle's edge may differ)."""
    # the cut: deepest level with <= 131072 nodes + leaves above it (tree.inl)
    n = len(flat["flags"])
    level = np.zeros(n, np.int32)
    inner = (flat["flags"] & 3) == 0
    ids = np.flatnonzero(inner)
    nc = np.array([bin(int(m)).count("1") for m in flat["child_mask"][ids]])
    for i, k in zip(ids, nc):
        level[flat["first_child"][i]:flat["first_child"][i] + k] = level[i] + 1
    at = np.bincount(level)
    leaves_at = np.bincount(level[~inner], minlength=len(at))
    best, above = 0, 0
    for l in range(len(at)):
        if at[l] + above <= 131072:
            best = l
        above += leaves_at[l]
    cut = np.flatnonzero((level == best) | ((level < best) & ~inner))
    mn, mx = flat["bb_min"][cut].astype(np.float64), flat["bb_max"][cut].astype(np.float64)
    pos = np.array(gc.position[:], np.float64)
    qw, qi, qj, qk = (float(x) for x in gc.rotation[:])
    lo = np.full((len(cut), 2), 1e30); hi = -lo
    for k in range(8):
        p = np.where(np.array([k & 1, k & 2, k & 4], bool), mx, mn) - pos
        u = -np.array([qi, qj, qk])
        t = 2 * np.cross(u, p)
        v = p + qw * t + np.cross(u, t)
        sx = (v[:, 0] / v[:, 1] + gc.xsub) / gc.mul
        sy = (gc.ysub - v[:, 2] / v[:, 1]) / gc.mul
        lo = np.minimum(lo, np.stack([sx, sy], 1)); hi = np.maximum(hi, np.stack([sx, sy], 1))
    cnt = np.zeros((TY, TX), np.int32)
    x0 = np.clip(np.floor(lo[:, 0]) - 2, 0, W).astype(int) >> 3
    x1 = (np.clip(np.ceil(hi[:, 0]) + 3, 0, W).astype(int) + 7) >> 3
    y0 = np.clip(np.floor(lo[:, 1]) - 2, 0, H).astype(int) >> 2
    y1 = (np.clip(np.ceil(hi[:, 1]) + 3, 0, H).astype(int) + 3) >> 2
    for a, b, c, d in zip(x0, x1, y0, y1):
        cnt[c:d, a:b] += 1
    return cnt.reshape(-1), len(cut)


for thr in [int(a) for a in sys.argv[1:]] or [48]:
    g.voxgpu_set_tile_feedback(thr)
    for f in (0, 90):
        ds.setup_camera(cam, *ds.orbit_pose(f), 1.2, W, H)
        gc = cam.gpu_camera()
        for _ in range(3):
            g.voxgpu_render(dev, C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, None, None, None)
        g.voxgpu_memset(d_tr.ptr, 0, ntiles * 24, None)
        g.voxgpu_set_trace(d_tr.ptr)
        g.voxgpu_render(dev, C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, None, None, None)
        g.voxgpu_set_trace(None)
        tr = d_tr.download(np.uint64, (ntiles, 3))
        ran = tr[:, 1] != 0                         # tiles that were traversed (dark tiles leave no record)
        t0, t1 = tr[ran, 0].astype(np.int64), tr[ran, 1].astype(np.int64)
        base = t0.min()
        t0 -= base; t1 -= base
        dur = (t1 - t0) / 1e3
        total = t1.max() / 1e3
        it = (tr[ran, 2] & np.uint64(0xffffffff)).astype(np.int64)
        print(f"feedback {thr} frame {f}: traversal span {total:.1f} us, traversed tiles {int(ran.sum())} of {ntiles}, tile us: median {np.median(dur):.2f} "
              f"p90 {np.percentile(dur, 90):.1f} p99 {np.percentile(dur, 99):.1f} max {dur.max():.1f}; sum {dur.sum()/1e3:.1f} ms")
        edges = np.linspace(0, t1.max(), 41)
        busy = [np.clip(np.minimum(t1, b) - np.maximum(t0, a), 0, None).sum() / (b - a) for a, b in zip(edges[:-1], edges[1:])]
        print("   busy warps per 2.5% slice:", " ".join(f"{x:.0f}" for x in busy))
        print("   iterations per tile: median %d p90 %d p99 %d max %d" % (np.median(it), np.percentile(it, 90), np.percentile(it, 99), it.max()))
        late = np.argsort(t1)[-5:]
        idx = np.flatnonzero(ran)
        for i in late:
            print(f"   late tile (x {idx[i] % TX}, y {idx[i] // TX}): start {t0[i]/1e3:.1f} end {t1[i]/1e3:.1f} us dur {dur[i]:.1f} iters {it[i]}")
        if thr == 48:
            cnt, ncut = box_counts(gc)
            c = cnt[ran]
            print(f"   cut boxes per traversed tile ({ncut} boxes): median {int(np.median(c))} p90 {int(np.percentile(c, 90))} p99 {int(np.percentile(c, 99))} max {c.max()}; "
                  f"corr(count, iterations) {np.corrcoef(c, it)[0, 1]:.3f}, corr(count, us) {np.corrcoef(c, dur)[0, 1]:.3f}")
            heavy = it >= np.percentile(it, 97)
            for T in (8, 12, 16, 24, 32, 48, 64, 96):
                sel = c >= T
                if sel.any():
                    print(f"   T={T:3d}: {int(sel.sum()):6d} tiles ({100*sel.mean():4.1f}% of traversed), holds {100*(heavy & sel).sum()/heavy.sum():5.1f}% of the slowest 3% "
                          f"and {100*dur[sel].sum()/dur.sum():4.1f}% of the warp time")
