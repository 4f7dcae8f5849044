"""Pins the oracle: the CPU restatement (oracle/vox_oracle.c) against
  (1) the reference's own test vectors and known answers,
  (2) golden fixtures generated from the reference's object code
      (tests/golden/make_golden.py), and
  (3) where oracle/_ref is present, the reference itself, bit for bit.
CPU only."""
import math
import os

import numpy as np
import pytest

from util import CROSS, ball_voxels, bits, check_tree_invariants, flat_equal, sha

THREADS = os.cpu_count() or 1


def test_cross_ray_reference_test_vector(port, golden):
    """tests/tests.c:600-659 test_tree_g676d50c: ray (-2,-2,0)+(1,1,-1) must hit the cross."""
    meta, arr = golden
    t = port.tree(CROSS)
    leaf, hit, _ = t.rays([[-2, -2, 0]], [[1, 1, -1]])
    assert leaf[0] >= 0
    assert leaf[0] == meta["cross_make_leaf"]
    assert np.array_equal(bits(hit[0]), bits(arr["cross_make_hit"]))


def test_benchmark_known_answers(port, golden):
    """benchmarks/voxtrees/hit-box.c:16-27 and ball-intersection.c:14-23."""
    meta, _ = golden
    one = port.tree(np.array([[0, 10, 0]], np.float32))
    j = np.arange(10, dtype=np.float32)
    o = np.stack([j - 5, 0 * j, 0 * j], 1)
    d = np.stack([(j - 5) / np.float32(10.0), 0 * j + 1, 0 * j], 1).astype(np.float32)
    assert int((one.rays(o, d)[0] >= 0).sum()) == meta["hitbox_hits"]
    one0 = port.tree(np.array([[0, 0, 0]], np.float32))
    jj = np.arange(-100, 100, dtype=np.float32)
    c = np.stack([0 * jj, 0 * jj + 10, jj], 1)
    assert int(one0.balls(c, np.full(200, 1000, np.float32)).sum()) == meta["ballbench_hits"] == 200


@pytest.mark.parametrize("kind", [0, 1])
def test_camera_known_answers(port, kind):
    """tests/tests.c:661-738 test_cameras / test_cameras_look_at for both camera classes."""
    cam = port.camera(kind)
    cam.set_position([0, 0, 0])
    assert np.allclose(cam.get_position(), 0)
    cam.set_window_size(100, 100)
    assert np.allclose(cam.screen2world(50, 50), [0, 1, 0], atol=1e-5)
    cam.set_rotation([0, 0, math.pi / 2])
    assert np.allclose(cam.screen2world(50, 50), [-1, 0, 0], atol=1e-5)
    cam.rotate([0, math.pi / 2, 0])
    assert np.allclose(cam.screen2world(50, 50), [-1, 0, 0], atol=1e-5)
    cam.move([0, 1, 0])
    assert np.allclose(cam.get_position(), [-1, 0, 0], atol=1e-5)
    # look_at: the centre ray passes through the target (tolerance 0.5, tests.c:717-731)
    cam = port.camera(kind)
    cam.set_window_size(100, 100)
    cam.set_position([-100, 20, -20])
    cam.look_at([0, 0, 0])
    d = cam.screen2world(50, 50).astype(np.float64)
    p = np.array([-100, 20, -20], np.float64)
    t = -p.dot(d) / d.dot(d)
    assert np.linalg.norm(p + t * d) < 0.5


def test_tree_golden(port, golden, skull, bonsai):
    """vox_make_tree restatement vs the reference's trees (hashes of every flattened array)."""
    meta, _ = golden
    for name, vox in (("skull", skull), ("bonsai", bonsai), ("ball50", ball_voxels(50))):
        t = port.tree(vox)
        flat = t.flatten()
        want = meta["trees"][name]
        for k, v in flat.items():
            got = sha(v) if isinstance(v, np.ndarray) else v
            assert got == want[k], (name, k)
        assert sha(t.reordered_set()) == want["reordered_sha"]
        ti = t.info()
        assert {f: int(getattr(ti, f)) for f, _ in ti._fields_} == want["counts"]


def test_check_tree_invariants_ball50(port):
    """tests/tests.c:360-365 test_tree_cons: structural invariants on the radius-50 ball."""
    flat = port.tree(ball_voxels(50)).flatten()
    assert check_tree_invariants(flat) > 1000


def test_skull_vectors_golden(port, golden, skull, ds):
    meta, arr = golden
    t = port.tree(skull)
    o, d = ds.random_rays(4096, side=256)
    leaf, hit, dot = t.rays(o, d)
    assert np.array_equal(leaf, arr["skull_rays_leaf"])
    m = leaf >= 0
    assert int(m.sum()) == meta["skull_rays_hits"]
    assert np.array_equal(bits(hit[m]), bits(arr["skull_rays_hit"][m]))
    # the port knows the winning dot; the reference glue derives "first dot whose box holds the point"
    fl = t.flatten()
    win = dot[m & (dot >= 0)]
    p = hit[m & (dot >= 0)]
    lo = fl["dots"][win]
    assert ((p >= lo) & (p <= lo + 1)).all()
    c, r = ds.random_balls(4096, fl["bb_min"][0], fl["bb_max"][0])
    assert np.array_equal(t.balls(c, r), arr["skull_balls"])


@pytest.mark.parametrize("frame", [0, 45, 90, 135, 180, 225, 270, 315])
def test_c2_frames_golden(port, golden, skull, ds, frame, vv):
    """Known answers of SURVEY.md 8d (C2) + digests of leaf ids, hit points and pixels."""
    meta, _ = golden
    tex = vv.Context(16, 4).texture()
    assert sha(tex) == meta["texture_sha"]
    t = test_c2_frames_golden.__dict__.setdefault("tree", port.tree(skull))
    pos, look = ds.orbit_pose(frame)
    cam = ds.setup_camera(port.camera(0), pos, look, 1.2, 1920, 1080)
    want = meta["c2"][str(frame)]
    assert np.array_equal(bits(cam.state()), bits(np.array(want["camera_state"], np.float32)))
    pix, leaf, hit = t.render(cam, 1920, 1080, 0, tex, threads=THREADS)
    m = leaf >= 0
    assert int(m.sum()) == want["hits"] == ds.C2_KNOWN_HITS[frame]
    assert sha(leaf) == want["leaf_sha"] and sha(hit[m]) == want["hit_sha"] and sha(pix) == want["pix_sha"]


def test_quality_modes_and_c1_golden(port, golden, skull, bonsai, ds, vv):
    meta, _ = golden
    tex = vv.Context(16, 4).texture()
    t = port.tree(skull)
    pos, look = ds.orbit_pose(0)
    for q, want in meta["c2_modes"].items():
        cam = ds.setup_camera(port.camera(0), pos, look, 1.2, 1920, 1080)
        pix = t.render(cam, 1920, 1080, int(q), tex, threads=THREADS, want_hits=False)
        assert int((pix != 0).sum()) == want["lit"] and sha(pix) == want["pix_sha"], q
    tb = port.tree(bonsai)
    for kind, name, p in ((0, "simple", (128, 128, 400)), (1, "doom", (100, -250, 180))):
        cam = ds.setup_camera(port.camera(kind), p, (128, 128, 128), 1.2, 800, 600)
        pix, leaf, hit = tb.render(cam, 800, 600, 0, tex, threads=THREADS)
        want = meta["c1"][name]
        m = leaf >= 0
        assert int(m.sum()) == want["hits"]
        assert sha(leaf) == want["leaf_sha"] and sha(hit[m]) == want["hit_sha"] and sha(pix) == want["pix_sha"]
    # invalid geometry / quality are rejected like renderer.c:97-100,179-195
    cam = ds.setup_camera(port.camera(0), pos, look, 1.2, 1920, 1080)
    assert t.render(cam, 1000, 1080, 0, tex) is None
    assert t.render(cam, 1920, 1080, 1 | 4, tex) is None
    assert t.render(cam, 1920, 1080, 3, tex) is None


# ------------------------------------------------------------------ port == reference, bit for bit

def test_port_equals_reference_trees(port, ref, skull):
    rng = np.random.default_rng(7)
    sparse = np.unique(rng.integers(0, 64, (20000, 3)), axis=0).astype(np.float32)
    rng.shuffle(sparse)
    for vox in (skull[::7], sparse, ball_voxels(12), CROSS, CROSS[:1], CROSS[:7], CROSS[:8]):
        a, b = port.tree(vox), ref.tree(vox)
        assert flat_equal(a.flatten(), b.flatten()) == []
        assert np.array_equal(a.reordered_set(), b.reordered_set())


def test_port_equals_reference_nonunit_voxel(port, ref):
    rng = np.random.default_rng(11)
    idx = np.unique(rng.integers(-40, 40, (30000, 3)), axis=0)
    rng.shuffle(idx)
    size = np.array([0.5, 2.0, 0.25], np.float32)
    vox = idx.astype(np.float32) * size
    try:
        for o in (port, ref):
            o.set_voxel(*size)
        a, b = port.tree(vox), ref.tree(vox)
        assert flat_equal(a.flatten(), b.flatten()) == []
        oo, dd = (rng.random((20000, 3)).astype(np.float32) - 0.5) * 120, rng.standard_normal((20000, 3)).astype(np.float32)
        ra, rb = a.rays(oo, dd), b.rays(oo, dd)
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(bits(ra[1]), bits(rb[1]))
        c, r = oo[:5000], rng.random(5000).astype(np.float32) * 6
        assert np.array_equal(a.balls(c, r), b.balls(c, r))
    finally:
        for o in (port, ref):
            o.set_voxel(1, 1, 1)


def test_port_equals_reference_queries(port, ref, skull, ds):
    a, b = port.tree(skull), ref.tree(skull)
    rng = np.random.default_rng(3)
    o, d = ds.random_rays(200000, side=256)
    # axis-aligned and degenerate directions, origins inside the volume
    d[:1000, 0] = 0
    d[1000:2000, 1:] = 0
    o[2000:4000] = rng.random((2000, 3)).astype(np.float32) * 256
    o[4000:5000] = np.floor(o[4000:5000])          # exactly on voxel faces
    ra, rb = a.rays(o, d, threads=THREADS), b.rays(o, d, threads=THREADS)
    assert np.array_equal(ra[0], rb[0])
    m = ra[0] >= 0
    assert np.array_equal(bits(ra[1][m]), bits(rb[1][m]))
    fl = a.flatten()
    c, r = ds.random_balls(200000, fl["bb_min"][0], fl["bb_max"][0])
    assert np.array_equal(a.balls(c, r, threads=THREADS), b.balls(c, r, threads=THREADS))
    # searches that start inside the tree (FAST mode does, renderer.c:297)
    leafs = np.flatnonzero((fl["flags"] & 3) == 1)[:64]
    for nid in leafs:
        la, lb = a.rays(o[:2000], d[:2000], start_node=int(nid)), b.rays(o[:2000], d[:2000], start_node=int(nid))
        assert np.array_equal(la[0], lb[0]) and np.array_equal(bits(la[1]), bits(lb[1]))


@pytest.mark.parametrize("kind", [0, 1])
def test_port_equals_reference_render(port, ref, bonsai, ds, kind, vv):
    tex = vv.Context(16, 4).texture()
    a, b = port.tree(bonsai), ref.tree(bonsai)
    for q in (0, 1, 2, 6, 10):
        pos = (128, 128, 400) if kind == 0 else (100, -250, 180)
        ca = ds.setup_camera(port.camera(kind), pos, (128, 128, 128), 1.2, 800, 600)
        cb = ds.setup_camera(ref.camera(kind), pos, (128, 128, 128), 1.2, 800, 600)
        for c in (ca, cb):
            c.rotate([0.05, -0.1, 0.2])
            c.move([3, -2, 1])
        assert np.array_equal(bits(ca.state()), bits(cb.state()))
        pa, pb = a.render(ca, 800, 600, q, tex, threads=THREADS), b.render(cb, 800, 600, q, tex, threads=THREADS)
        assert np.array_equal(pa[0], pb[0]) and np.array_equal(pa[1], pb[1])
        assert np.array_equal(bits(pa[2]), bits(pb[2]))


def test_reference_insertion_built_cross(ref, golden):
    """The reference's test builds the cross by vox_insert_voxel; the golden file keeps that tree."""
    meta, arr = golden
    t = ref.tree(CROSS, by_insertion=True)
    leaf, hit, _ = t.rays([[-2, -2, 0]], [[1, 1, -1]])
    assert leaf[0] == meta["cross_ins_leaf"] and np.array_equal(bits(hit[0]), bits(arr["cross_ins_hit"]))
    fl = t.flatten()
    for k in ("bb_min", "bb_max", "flags", "dots_num", "center", "first_child", "child_mask", "dots_off", "dots"):
        assert np.array_equal(fl[k], arr["cross_ins_flat_" + k]), k


def test_scalar_vs_sse_divergence_fixture(orc, ref, ds):
    """tests/golden/sse_divergence.json (make_sse_divergence.py) states how far the reference's own two
    builds are apart on the C2 golden frames; bench.py quotes it beside the SSE-timed CPU arm.  Re-measured
    here for frame 0 on the same rays (the part that does not depend on the host's rsqrtps)."""
    import json
    if not orc.available("ref_sse"):
        pytest.skip("oracle/_ref/libvoxref_sse.so not built")
    with open(os.path.join(os.path.dirname(__file__), "golden", "sse_divergence.json")) as f:
        fx = json.load(f)
    assert set(fx["frames"]) == {str(k) for k in ds.C2_KNOWN_HITS}
    for k, want in ds.C2_KNOWN_HITS.items():
        assert fx["frames"][str(k)]["same_rays"]["hits_scalar"] == want
    sse = orc.Oracle("ref_sse")
    vox = ds.volume_voxels("skull")
    ts, tv = ref.tree(vox), sse.tree(vox)
    cam = ds.setup_camera(ref.camera(0), *ds.orbit_pose(0, 360), 1.2, 1920, 1080)
    o, d = cam.rays(1920, 1080)
    ls, _, _ = ts.rays(o, d, threads=THREADS)
    lv, _, _ = tv.rays(o, d, threads=THREADS)
    got = fx["frames"]["0"]["same_rays"]
    assert int((lv >= 0).sum()) == got["hits_sse"]
    assert int((ls != lv).sum()) == got["leaf_differs"]
    assert int(((ls >= 0) != (lv >= 0)).sum()) == got["hit_miss_differs"]
