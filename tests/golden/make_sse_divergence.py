#!/usr/bin/env python
"""Scalar <-> SSE divergence of the reference itself on the C2 golden frames.

Parity of this repository is defined against the reference's SCALAR build (oracle/_ref/libvoxref_scalar.so,
-O2 -ffp-contract=off); the CPU timing baseline of bench.py is the reference's default SSE build
(oracle/_ref/libvoxref_sse.so), whose hit_box differs (geom-sse.c:84-116: t zeroed where d == 0, every
lane with t == max takes the clamped value) and whose look_at normalises with rsqrtps
(vect-ops-sse.c:73-79).  This script measures, with the reference's own object code only, how far the two
builds are apart on each of the eight golden C2 frames, so that the number can travel to the GPU box as
a fixture (tests/golden/sse_divergence.json) and be quoted in bench.py's line:

  same_rays   both builds search the SAME rays (the scalar build's camera): isolates geom-sse.c / search
  own_camera  each build renders with its own camera state: what a user of either build sees

Needs /root/reference (oracle/_ref built by `make -C oracle ref`); run from the repo root.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc                       # noqa: E402
from tools import datasets as ds             # noqa: E402

W, H, FOV, FRAMES = 1920, 1080, 1.2, 360


def main():
    S, V = orc.Oracle("ref"), orc.Oracle("ref_sse")
    vox = ds.volume_voxels("skull")
    ts, tv = S.tree(vox), V.tree(vox)
    fs, fv = ts.flatten(), tv.flatten()
    same_tree = all(np.array_equal(fs[k], fv[k]) for k in fs if k != "depth")
    threads = os.cpu_count() or 1
    out = {"scalar": S.name, "sse": V.name, "trees_identical": bool(same_tree), "frames": {}}
    for f in sorted(ds.C2_KNOWN_HITS):
        cs = ds.setup_camera(S.camera(0), *ds.orbit_pose(f, FRAMES), FOV, W, H)
        cv = ds.setup_camera(V.camera(0), *ds.orbit_pose(f, FRAMES), FOV, W, H)
        o, d = cs.rays(W, H)
        ls, ps, _ = ts.rays(o, d, threads=threads)
        lv, pv, _ = tv.rays(o, d, threads=threads)
        ov, dv = cv.rays(W, H)
        lo, po, _ = tv.rays(ov, dv, threads=threads)

        def diff(la, pa, lb, pb):
            both = (la >= 0) & (lb >= 0)
            return {"hits_scalar": int((la >= 0).sum()), "hits_sse": int((lb >= 0).sum()),
                    "hit_miss_differs": int(((la >= 0) != (lb >= 0)).sum()),
                    "leaf_differs": int((la != lb).sum()),
                    "point_bits_differ_where_both_hit": int((pa[both].view(np.uint32) != pb[both].view(np.uint32)).any(1).sum()),
                    "max_point_distance_where_same_leaf": float(np.abs(pa[both & (la == lb)] - pb[both & (la == lb)]).max(initial=0.0))}
        out["frames"][str(f)] = {"same_rays": diff(ls, ps, lv, pv), "own_camera": diff(ls, ps, lo, po),
                                 "camera_state_identical": bool(np.array_equal(cs.state(), cv.state())),
                                 "dirs_bits_differ": int((d.view(np.uint32) != dv.view(np.uint32)).any(1).sum())}
        assert out["frames"][str(f)]["same_rays"]["hits_scalar"] == ds.C2_KNOWN_HITS[f]
        print(f, json.dumps(out["frames"][str(f)]), flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "sse_divergence.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
