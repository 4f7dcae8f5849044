"""Inputs of BASELINE.json's five configs, made concrete as SURVEY.md 8(d) fixes them.

Shared by tests/ and bench.py.  Pure numpy; nothing here is timed or shipped
as product code.

  C1  bonsai (sample > 40, 1 591 663 voxels), 800x600, camera (128,128,400)
  C2  skull  (sample > 40, 1 073 426 voxels), 1920x1080, 360-frame xz orbit
  C3  ~10 M random voxels in 512^3 (splitmix64, p = 0.075), 16 Mi random rays
  C4  skull, 4 Mi random balls
  C5  ~100 M voxels in 1024^3 (union of balls with hash holes), 7680x4320
"""
import lzma
import math
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "data")

SEED = 0x9E3779B97F4A7C15
GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def occupancy(name):
    """256^3 bool array [x,y,z] of `sample > 40` for 'skull' or 'bonsai'."""
    with open(os.path.join(DATA, f"{name}_gt40.bits.xz"), "rb") as f:
        bits = np.frombuffer(lzma.decompress(f.read()), np.uint8)
    return np.unpackbits(bits).astype(bool).reshape(256, 256, 256)


def volume_voxels(name):
    """Voxels in the order vox_read_raw_data emits them (x outer, z inner;
    reference src/voxtrees/datareader.c:51-75), vox_voxel = (1,1,1)."""
    return np.argwhere(occupancy(name)).astype(np.float32)


def write_raw_volume(name, path):
    """A raw 256^3 x 1-byte file whose `sample > 40` set equals the fixture
    (255 where occupied, 0 elsewhere) -- exercises the raw-file loader."""
    (occupancy(name).astype(np.uint8) * 255).tofile(path)


# ---------------------------------------------------------------- splitmix64

def splitmix64(start_index, count, seed=SEED):
    """Draws number start_index .. start_index+count-1 (0-based) of the stream
    whose state starts at `seed`; each draw first adds the golden-ratio
    increment, then mixes."""
    with np.errstate(over="ignore"):
        i = np.arange(start_index + 1, start_index + count + 1, dtype=np.uint64)
        z = np.uint64(seed) + i * GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(start_index, count, seed=SEED):
    """float64 in [0,1) from the top 53 bits of the draws."""
    return (splitmix64(start_index, count, seed) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def random_voxels(side=512, p=0.075, seed=SEED):
    """C3 voxels: Bernoulli(p) per cell, x-outer / z-inner order.
    Returns (xyz float32 (n,3), draws_consumed)."""
    thr = np.uint64(int(p * 2.0 ** 64))
    out = []
    plane = side * side
    for x in range(side):
        d = splitmix64(x * plane, plane, seed)
        keep = np.flatnonzero(d < thr)
        v = np.empty((keep.size, 3), np.float32)
        v[:, 0] = x
        v[:, 1] = keep // side
        v[:, 2] = keep % side
        out.append(v)
    return np.concatenate(out), side ** 3


def random_rays(n, side=512, start_index=0, seed=SEED):
    """C3 rays: origin uniform in the cube of side 3*side centred on the volume,
    direction = (uniform point in the volume) - origin, un-normalised."""
    u = uniform01(start_index, 6 * n, seed).reshape(n, 6)
    o = (u[:, 0:3] * 3.0 - 1.0) * side
    t = u[:, 3:6] * side
    return o.astype(np.float32), (t - o).astype(np.float32)


def random_balls(n, bb_min, bb_max, seed=SEED + 1):
    """C4 balls: centres uniform in the tree bbox expanded by 16, radii in [0.5, 8)."""
    u = uniform01(0, 4 * n, seed).reshape(n, 4)
    lo = np.asarray(bb_min, np.float64) - 16.0
    hi = np.asarray(bb_max, np.float64) + 16.0
    c = lo + u[:, 0:3] * (hi - lo)
    r = 0.5 + u[:, 3] * 7.5
    return c.astype(np.float32), r.astype(np.float32)


def ball_union_voxels(side=1024, n_balls=4096, seed=SEED + 2, rmin=8.0, rmax=28.0):
    """C5 voxels: union of solid balls with 1-in-8 hash-selected holes, in
    x-outer / z-inner order.  ~100 M voxels at the default arguments."""
    u = uniform01(0, 4 * n_balls, seed).reshape(n_balls, 4)
    cx, cy, cz = (u[:, 0] * side), (u[:, 1] * side), (u[:, 2] * side)
    rad = rmin + u[:, 3] * (rmax - rmin)
    occ = np.zeros((side, side, side), bool)
    for b in range(n_balls):
        r = rad[b]
        x0, x1 = max(0, int(cx[b] - r)), min(side, int(cx[b] + r) + 2)
        y0, y1 = max(0, int(cy[b] - r)), min(side, int(cy[b] + r) + 2)
        z0, z1 = max(0, int(cz[b] - r)), min(side, int(cz[b] + r) + 2)
        xs = (np.arange(x0, x1) - cx[b])[:, None, None]
        ys = (np.arange(y0, y1) - cy[b])[None, :, None]
        zs = (np.arange(z0, z1) - cz[b])[None, None, :]
        occ[x0:x1, y0:y1, z0:z1] |= (xs * xs + ys * ys + zs * zs) < r * r
    out = []
    for x in range(side):
        yz = np.flatnonzero(occ[x])
        if yz.size == 0:
            continue
        y, z = yz // side, yz % side
        with np.errstate(over="ignore"):
            h = (np.uint64(x) * np.uint64(0x9E3779B1) ^ y.astype(np.uint64) * np.uint64(0x85EBCA77) ^
                 z.astype(np.uint64) * np.uint64(0xC2B2AE3D))
            h ^= h >> np.uint64(15)
            h *= np.uint64(0x2C1B3C6D)
            h ^= h >> np.uint64(12)
        keep = (h & np.uint64(7)) != np.uint64(0)
        v = np.empty((int(keep.sum()), 3), np.float32)
        v[:, 0] = x
        v[:, 1] = y[keep]
        v[:, 2] = z[keep]
        out.append(v)
    return np.concatenate(out)


# ------------------------------------------------------------------ cameras

def edit_script(n_edits=3000, seed=SEED + 3):
    """A deterministic online-edit workload (f-3): start from a radius-12 voxel ball plus a solid
    6x6x6 box and a sparse cloud, then insert / delete voxels -- deletions mostly of present voxels,
    insertions mostly next to the solid parts (the dense-leaf rules of tree.c:461-503, 641-756).
    Returns (voxels, insert flags, points); points are un-aligned on purpose (vox_align)."""
    g = np.arange(-12, 12)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    v = np.stack([x, y, z], -1).reshape(-1, 3)
    ball = v[(v * v).sum(1) < 144]
    g = np.arange(0, 6)
    box = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + np.array([14, -3, -3])
    u = uniform01(0, 3 * 400 + 5 * n_edits, seed)
    cloud = np.floor(u[:1200].reshape(400, 3) * 30 - np.array([15, 15, -14])).astype(np.int64)
    vox = np.unique(np.concatenate([ball, box, cloud]), axis=0).astype(np.float32)
    e = u[1200:].reshape(n_edits, 5)
    ins = e[:, 0] < 0.5
    pts = (e[:, 1:4] * np.array([40, 34, 46]) - np.array([16, 17, 17])).astype(np.float32)
    pick = (e[:, 4] * len(vox)).astype(np.int64)
    use_existing = ~ins & (e[:, 4] < 0.8)
    pts[use_existing] = vox[(pick[use_existing] * 5) % len(vox)] + np.float32(0.25)
    return vox, ins, pts


def light_script(n=120, lo=-30.0, hi=290.0, seed=SEED + 4):
    """A deterministic light workload (f-4) for the 256^3 bundled volumes: n shadowless lights
    (centre uniform in [lo,hi)^3, radius in [12,90), colour in [0,1]^3), every fourth one deleted
    again after 8 more insertions (exercises M-tree splits, leaf drops and re-parenting,
    mtree.c:113-188, 274-330), and a dim ambient colour.
    Returns (ambient, ops) with ops = list of ('+', centre, radius, colour) / ('-', centre, radius)."""
    u = uniform01(0, 7 * n + 3, seed)
    amb = (u[:3] * 0.35).astype(np.float32)
    e = u[3:].reshape(n, 7).astype(np.float32)
    ops, pending = [], []
    for i in range(n):
        c = (e[i, :3] * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)
        r = np.float32(e[i, 3] * np.float32(78.0) + np.float32(12.0))
        ops.append(("+", c, float(r), e[i, 4:7].copy()))
        if i % 4 == 1:
            pending.append((i + 8, c, float(r)))
        while pending and pending[0][0] <= i:
            _, pc, pr = pending.pop(0)
            ops.append(("-", pc, pr))
    return amb, ops


def apply_lights(lm, script):
    """Run a light_script() on anything with set_ambient / insert / delete (product or oracle)."""
    amb, ops = script
    res = [lm.set_ambient(amb)]
    for op in ops:
        res.append(lm.insert(op[1], op[2], op[3]) if op[0] == "+" else lm.delete(op[1], op[2]))
    return res


def orbit_pose(frame, frames=360, centre=(128.0, 128.0, 128.0), radius=328.0):
    """C2 camera path: position on the xz orbit, looking at the centre.
    Frame 0 is exactly examples/scene1.cfg's camera (128,128,-200)."""
    th = frame * 2.0 * math.pi / frames
    pos = (np.float32(centre[0] + radius * math.sin(th)), np.float32(centre[1]),
           np.float32(centre[2] - radius * math.cos(th)))
    return pos, tuple(np.float32(c) for c in centre)


def setup_camera(cam, pos, look, fov, w, h):
    """The demo's call order (src/demo/main.c:195-213): position, fov, look_at,
    then set_window_size.  Works for product and oracle camera wrappers alike."""
    cam.set_position(pos)
    cam.set_fov(fov)
    cam.look_at(look)
    cam.set_window_size(w, h)
    return cam


# Known answers from the scalar -ffp-contract=off reference build:
# hits out of 2 073 600 pixels on C2 frames.
C2_KNOWN_HITS = {0: 415817, 45: 464270, 90: 602862, 135: 526881,
                 180: 562899, 225: 473459, 270: 345483, 315: 351487}
