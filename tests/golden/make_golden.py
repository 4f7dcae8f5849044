#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the REFERENCE ITSELF.

Run in the build container (needs /root/reference and `make -C oracle ref`):
the outputs come from oracle/_ref/libvoxref_scalar.so, i.e. the reference's
own search.c / geom.c / tree.c / *-camera.c object code (scalar build,
gcc -O2 -ffp-contract=off -- the oracle of record) driven by the
restated vox_render block loop.  The fixtures pin (a) the CPU restatement
oracle/vox_oracle.c and (b) the CUDA path, on machines where /root/reference
does not exist.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402
from tools import datasets as ds  # noqa: E402
import voxvision_b200 as vv  # noqa: E402  (only for the default texture bytes, host code)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def tree_digest(flat):
    return {k: (sha(v) if isinstance(v, np.ndarray) else v) for k, v in flat.items()}


def frame_digest(res, leaf_only=False):
    pix, leaf, hit = res
    m = leaf >= 0
    d = {"hits": int(m.sum()), "leaf_sha": sha(leaf), "hit_sha": sha(hit[m])}
    if not leaf_only:
        d["pix_sha"] = sha(pix)
    return d


# the 21-voxel 3-D cross of tests/tests.c:600-659, in the reference's insertion order
CROSS = [(0, 0, 0), (0, 0, -1), (0, 0, -2), (0, 0, 1), (0, 0, 2),
         (-1, 0, 0), (-2, 0, 0), (-3, 0, 0), (-4, 0, 0), (1, 0, 0), (2, 0, 0), (3, 0, 0), (4, 0, 0),
         (0, -1, 0), (0, -2, 0), (0, -3, 0), (0, -4, 0), (0, 1, 0), (0, 2, 0), (0, 3, 0), (0, 4, 0)]


def ball_voxels(radius):
    """prepare_tree of tests/tests.c:214-237 (radius 50 there)."""
    g = np.arange(-radius, radius)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    keep = (x * x + y * y + z * z).reshape(-1) < radius * radius
    return np.stack([x, y, z], -1).reshape(-1, 3)[keep].astype(np.float32)


def main():
    R = orc.Oracle("ref")
    assert R.name == "reference-scalar"
    threads = os.cpu_count()
    meta, arrays = {"oracle": R.name}, {}
    tex = vv.Context(16, 4).texture()
    meta["texture_sha"] = sha(tex)

    # ---- reference test vector: the cross, built by insertion as the test does
    cross = np.array(CROSS, np.float32)
    t_ins = R.tree(cross, by_insertion=True)
    t_make = R.tree(cross)
    o, d = np.array([[-2, -2, 0]], np.float32), np.array([[1, 1, -1]], np.float32)
    l1, h1, _ = t_ins.rays(o, d)
    l2, h2, _ = t_make.rays(o, d)
    assert l1[0] >= 0 and l2[0] >= 0          # CU_ASSERT (leaf != NULL), tests.c:658
    arrays["cross_ins_hit"], arrays["cross_make_hit"] = h1[0], h2[0]
    meta["cross_ins_leaf"], meta["cross_make_leaf"] = int(l1[0]), int(l2[0])
    ins_flat = t_ins.flatten()
    for k, v in ins_flat.items():
        if isinstance(v, np.ndarray):
            arrays["cross_ins_flat_" + k] = v
    meta["cross_ins_depth"] = ins_flat["depth"]
    oo, dd = ds.random_rays(2048, side=10)
    oo -= 5
    l, h, _ = t_ins.rays(oo, dd)
    arrays["cross_ins_rays_leaf"], arrays["cross_ins_rays_hit"] = l, h

    # ---- trees
    skull = ds.volume_voxels("skull")
    bonsai = ds.volume_voxels("bonsai")
    T = {"skull": R.tree(skull), "bonsai": R.tree(bonsai), "ball50": R.tree(ball_voxels(50))}
    meta["trees"] = {k: tree_digest(t.flatten()) for k, t in T.items()}
    for k, t in T.items():
        ti = t.info()
        meta["trees"][k]["counts"] = {f: int(getattr(ti, f)) for f, _ in ti._fields_}
        meta["trees"][k]["reordered_sha"] = sha(t.reordered_set())

    # ---- C2 frames (skull 1080p orbit), all quality modes on frame 0
    W, H = 1920, 1080
    meta["c2"] = {}
    for f in sorted(ds.C2_KNOWN_HITS):
        pos, look = ds.orbit_pose(f)
        cam = ds.setup_camera(R.camera(orc.CAMERA_SIMPLE), pos, look, 1.2, W, H)
        dg = frame_digest(T["skull"].render(cam, W, H, orc.QUALITY_BEST, tex, threads=threads))
        assert dg["hits"] == ds.C2_KNOWN_HITS[f], (f, dg["hits"])
        dg["camera_state"] = [float(x) for x in cam.state()]
        meta["c2"][str(f)] = dg
    meta["c2_modes"] = {}
    pos, look = ds.orbit_pose(0)
    for q in (1, 2, 2 | 4, 2 | 8):
        cam = ds.setup_camera(R.camera(orc.CAMERA_SIMPLE), pos, look, 1.2, W, H)
        pix = T["skull"].render(cam, W, H, q, tex, threads=threads, want_hits=False)
        meta["c2_modes"][str(q)] = {"lit": int((pix != 0).sum()), "pix_sha": sha(pix)}

    # ---- C1 (bonsai 800x600, examples/scene2.cfg) with both cameras
    meta["c1"] = {}
    for kind, name, pos in ((orc.CAMERA_SIMPLE, "simple", (128, 128, 400)), (orc.CAMERA_DOOM, "doom", (100, -250, 180))):
        cam = ds.setup_camera(R.camera(kind), pos, (128, 128, 128), 1.2, 800, 600)
        meta["c1"][name] = frame_digest(T["bonsai"].render(cam, 800, 600, orc.QUALITY_BEST, tex, threads=threads))
    # the reference's other quality modes (renderer.c:240-316), pixels only
    cam = ds.setup_camera(R.camera(orc.CAMERA_SIMPLE), (128, 128, 400), (128, 128, 128), 1.2, 800, 600)
    meta["c1"]["modes"] = {}
    for qname, q in (("fast", 1), ("adaptive", 2), ("adaptive+merge", 2 | 4), ("adaptive+merge_accurate", 2 | 8)):
        pix = T["bonsai"].render(cam, 800, 600, q, tex, threads=threads, want_hits=False)
        meta["c1"]["modes"][qname] = sha(pix)

    # ---- raw vectors: 4096 rays and 4096 balls against skull
    o, d = ds.random_rays(4096, side=256)
    leaf, hit, dot = T["skull"].rays(o, d)
    arrays["skull_rays_leaf"], arrays["skull_rays_hit"], arrays["skull_rays_dot"] = leaf, hit, dot
    fl = T["skull"].flatten()
    c, r = ds.random_balls(4096, fl["bb_min"][0], fl["bb_max"][0])
    arrays["skull_balls"] = T["skull"].balls(c, r)
    meta["skull_rays_hits"] = int((leaf >= 0).sum())
    meta["skull_balls_hits"] = int(arrays["skull_balls"].sum())

    # ---- online edits (tree.c:437-866): the tree after a scripted burst of insertions / deletions,
    #      its rays, and the tree vox_rebuild_tree makes of it
    vox, ins, pts = ds.edit_script()
    te = R.tree(vox)
    arrays["edit_results"] = te.edit(ins, pts)
    ef = te.flatten()
    meta["edits"] = {"tree": tree_digest(ef), "applied": int(arrays["edit_results"].sum()),
                     "voxels": int(ef["dots_num"][0]), "rebuilt": tree_digest(te.rebuild().flatten())}
    oo, dd = ds.random_rays(4096, side=48)
    oo -= 24
    l, h, _ = te.rays(oo, dd)
    arrays["edit_rays_leaf"], arrays["edit_rays_hit"] = l, h
    meta["edits"]["ray_hits"] = int((l >= 0).sum())

    # ---- lights (lights.c over mtree.c): the M-tree after a scripted set of insertions / deletions,
    #      the light at sample points, and lit frames (renderer.c:41-49)
    script = ds.light_script()
    L = R.lights()
    arrays["light_results"] = np.array(ds.apply_lights(L, script), np.int32)
    lf = L.flatten()
    meta["lights"] = {"flat": {k: sha(v) for k, v in lf.items()}, "count": L.count(),
                      "nodes": int(lf["node_bounds"].shape[0])}
    pts = ds.uniform01(0, 3 * 4096, ds.SEED + 5).reshape(4096, 3).astype(np.float32) * np.float32(300) - np.float32(20)
    arrays["light_at_points"] = L.get_light(pts)
    cam = ds.setup_camera(R.camera(orc.CAMERA_SIMPLE), (128, 128, 400), (128, 128, 128), 1.2, 800, 600)
    meta["lights"]["c1_frames"] = {}
    for qname, q in (("best", 0), ("adaptive", 2), ("adaptive+merge", 2 | 4)):
        pix = T["bonsai"].render_lit(cam, 800, 600, L, q, tex, threads=threads)
        meta["lights"]["c1_frames"][qname] = {"pix_sha": sha(pix), "lit": int((pix != 0).sum())}

    # reference benchmark known answers (benchmarks/voxtrees/hit-box.c:16-27, ball-intersection.c:14-23)
    one = R.tree(np.array([[0, 10, 0]], np.float32), by_insertion=True)
    j = np.arange(10, dtype=np.float32)
    oo = np.stack([j - 5, 0 * j, 0 * j], 1)
    dd = np.stack([(j - 5) / np.float32(10.0), 0 * j + 1, 0 * j], 1).astype(np.float32)
    meta["hitbox_hits"] = int((one.rays(oo, dd)[0] >= 0).sum())
    one0 = R.tree(np.array([[0, 0, 0]], np.float32), by_insertion=True)
    jj = np.arange(-100, 100, dtype=np.float32)
    cc = np.stack([0 * jj, 0 * jj + 10, jj], 1)
    meta["ballbench_hits"] = int(one0.balls(cc, np.full(200, 1000, np.float32)).sum())

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote golden.json / golden.npz;", {k: v["hits"] for k, v in meta["c2"].items()},
          "hitbox", meta["hitbox_hits"], "ballbench", meta["ballbench_hits"])


if __name__ == "__main__":
    main()
