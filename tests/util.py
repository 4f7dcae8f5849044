"""Helpers shared by the test modules."""
import hashlib

import numpy as np


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(a):
    """float32 array -> its bit patterns, for bit-exact comparison (NaN-safe)."""
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def flat_equal(a, b):
    """Compare two canonical flat trees (dicts) bit-for-bit; returns list of differing keys."""
    bad = []
    for k in ("bb_min", "bb_max", "flags", "dots_num", "center", "first_child", "child_mask", "dots_off", "dots"):
        x, y = a[k], b[k]
        if x.shape != y.shape:
            bad.append(k + ":shape")
            continue
        if x.dtype == np.float32:
            x, y = bits(x), bits(y)
        if not np.array_equal(x, y):
            bad.append(k)
    if a["depth"] != b["depth"]:
        bad.append("depth")
    return bad


def ball_voxels(radius):
    """The reference's own test scene: prepare_tree, tests/tests.c:214-237."""
    g = np.arange(-radius, radius)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    keep = (x * x + y * y + z * z).reshape(-1) < radius * radius
    return np.stack([x, y, z], -1).reshape(-1, 3)[keep].astype(np.float32)


CROSS = np.array(
    [(0, 0, 0), (0, 0, -1), (0, 0, -2), (0, 0, 1), (0, 0, 2),
     (-1, 0, 0), (-2, 0, 0), (-3, 0, 0), (-4, 0, 0), (1, 0, 0), (2, 0, 0), (3, 0, 0), (4, 0, 0),
     (0, -1, 0), (0, -2, 0), (0, -3, 0), (0, -4, 0), (0, 1, 0), (0, 2, 0), (0, 3, 0), (0, 4, 0)], np.float32)


def check_tree_invariants(flat, voxel=(1.0, 1.0, 1.0)):
    """check_tree of the reference's test-suite (tests/tests.c:239-300) on flat arrays."""
    voxel = np.asarray(voxel, np.float32)
    n = flat["flags"].shape[0]
    fl = flat["flags"]
    size = flat["bb_max"] - flat["bb_min"]
    vol = (size[:, 0] * size[:, 1] * size[:, 2]) / (voxel[0] * voxel[1] * voxel[2])
    dense = (fl & 2) != 0
    # dense leaves hold exactly as many voxels as their box has room for
    assert np.array_equal(flat["dots_num"][dense].astype(np.uint64), vol[dense].astype(np.uint64))
    leaf = ((fl & 1) != 0) & ~dense
    inner = (fl & 3) == 0
    for i in np.flatnonzero(leaf):
        d = flat["dots"][int(flat["dots_off"][i]):int(flat["dots_off"][i]) + int(flat["dots_num"][i])]
        assert (d >= flat["bb_min"][i]).all() and (d + voxel <= flat["bb_max"][i]).all()
    half = voxel / 2
    for i in np.flatnonzero(inner):
        mask, first = int(flat["child_mask"][i]), int(flat["first_child"][i])
        total, c = 0, first
        for s in range(8):
            if not (mask >> s) & 1:
                continue
            total += int(flat["dots_num"][c])
            assert (flat["bb_min"][c] >= flat["bb_min"][i]).all() and (flat["bb_max"][c] <= flat["bb_max"][i]).all()
            for corner in (flat["bb_min"][c] + half, flat["bb_max"][c] - half):
                sub = sum(int(flat["center"][i][k] > corner[k]) << k for k in range(3))
                assert sub == s
            c += 1
        assert total == int(flat["dots_num"][i])
    return n
