"""Parity tests proper: the CUDA path, called through the C-ABI, against the
oracle (the reference's own object code when oracle/_ref is present, else the
CPU restatement) and against the committed golden fixtures.  Bit-exact
everywhere: leaf ids, hit points (fp32 bit patterns), voxel indices, collision
booleans, pixels."""
import ctypes as C
import os

import numpy as np
import pytest

from util import CROSS, ball_voxels, bits, flat_equal, sha

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


def render_device(vv, tree, cam, w, h, tex, fmt, y0=0, y1=None, want_hits=True, gcam=None, quality=0):
    """voxgpu_render into device buffers (the entry multi-GPU drivers use)."""
    api = vv.api
    y1 = h if y1 is None else y1
    rows = y1 - y0
    d_tex = vv.DeviceBuffer.from_array(tex)
    d_pix = vv.DeviceBuffer(w * rows * 4)
    d_leaf = vv.DeviceBuffer(w * rows * 4)
    d_pts = vv.DeviceBuffer(w * rows * 16)
    api.libs.gpu.voxgpu_memset(d_pix.ptr, 0xff, w * rows * 4, None)
    gc = gcam or cam.gpu_camera()
    fr = api.GpuFrame(w, h, y0, y1, w, quality, fmt)
    rc = api.libs.gpu.voxgpu_render(tree.device_handle(), C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr,
                                    d_leaf.ptr if want_hits else None, d_pts.ptr if want_hits else None, None)
    assert rc == 0, api.libs.gpu.voxgpu_last_error()
    pix = d_pix.download(np.uint32, (rows, w))
    leaf = d_leaf.download(np.int32, (rows, w)) if want_hits else None
    pts = d_pts.download(np.float32, (rows, w, 4))[:, :, :3] if want_hits else None
    return pix, leaf, pts


def assert_rays_equal(got, want, check_voxel=True):
    leaf, pts, vox = got
    oleaf, opts, odot = want
    assert np.array_equal(leaf, oleaf), f"{int((leaf != oleaf).sum())} leaf ids differ"
    m = oleaf >= 0
    assert np.array_equal(bits(pts[m]), bits(opts[m])), "hit points differ"
    if check_voxel:
        assert np.array_equal(vox[m].astype(np.int64), odot[m]), "hit voxels differ"
    assert (vox[~m] == -1).all()


# ------------------------------------------------------------------ K2 rays

def test_rays_skull_golden(gpu, golden, skull, ds):
    """Against vectors produced by the reference's object code (no oracle library at run time)."""
    meta, arr = golden
    t = gpu.Tree.build(skull)
    o, d = ds.random_rays(4096, side=256)
    leaf, pts, vox = t.rays(o, d)
    assert np.array_equal(leaf, arr["skull_rays_leaf"])
    m = leaf >= 0
    assert int(m.sum()) == meta["skull_rays_hits"]
    assert np.array_equal(bits(pts[m]), bits(arr["skull_rays_hit"][m]))
    c, r = ds.random_balls(4096, *t.bounding_box())
    assert np.array_equal(t.balls(c, r), arr["skull_balls"])


def test_cross_reference_test_vector(gpu, golden):
    """tests/tests.c:600-659: on the make_tree-built cross and on the reference's insertion-built one."""
    meta, arr = golden
    t = gpu.Tree.build(CROSS)
    leaf, p = t.ray([-2, -2, 0], [1, 1, -1])
    assert leaf == meta["cross_make_leaf"] and np.array_equal(bits(p), bits(arr["cross_make_hit"]))
    flat = {k[len("cross_ins_flat_"):]: v for k, v in arr.items() if k.startswith("cross_ins_flat_")}
    flat["depth"] = meta["cross_ins_depth"]
    ti = gpu.Tree.from_arrays(flat)
    leaf, p = ti.ray([-2, -2, 0], [1, 1, -1])
    assert leaf == meta["cross_ins_leaf"] and np.array_equal(bits(p), bits(arr["cross_ins_hit"]))
    oo, dd = __import__("tools.datasets", fromlist=["x"]).random_rays(2048, side=10)
    oo -= 5
    l, h, _ = ti.rays(oo, dd)
    assert np.array_equal(l, arr["cross_ins_rays_leaf"])
    assert np.array_equal(bits(h[l >= 0]), bits(arr["cross_ins_rays_hit"][l >= 0]))


def test_benchmark_known_answers(gpu, golden):
    meta, _ = golden
    one = gpu.Tree.build(np.array([[0, 10, 0]], np.float32))
    j = np.arange(10, dtype=np.float32)
    o = np.stack([j - 5, 0 * j, 0 * j], 1)
    d = np.stack([(j - 5) / np.float32(10.0), 0 * j + 1, 0 * j], 1).astype(np.float32)
    assert int((one.rays(o, d)[0] >= 0).sum()) == meta["hitbox_hits"]
    one0 = gpu.Tree.build(np.array([[0, 0, 0]], np.float32))
    jj = np.arange(-100, 100, dtype=np.float32)
    c = np.stack([0 * jj, 0 * jj + 10, jj], 1)
    assert int(one0.balls(c, np.full(200, 1000, np.float32)).sum()) == 200


@pytest.mark.parametrize("scene", ["skull", "bonsai", "ball50", "sparse", "cross", "dup"])
def test_rays_and_balls_vs_oracle(gpu, checker, ds, skull, bonsai, scene):
    rng = np.random.default_rng(17)
    if scene == "skull":
        vox, side, off = skull, 256, 0
    elif scene == "bonsai":
        vox, side, off = bonsai, 256, 0
    elif scene == "ball50":
        vox, side, off = ball_voxels(50), 100, -50
    elif scene == "sparse":
        vox = np.unique(rng.integers(0, 128, (150000, 3)), axis=0).astype(np.float32)
        rng.shuffle(vox)
        side, off = 128, 0
    elif scene == "cross":
        vox, side, off = CROSS, 10, -5
    else:   # duplicates -> an overflowed leaf with > 7 dots next to ordinary nodes
        vox = np.concatenate([np.tile(np.array([[3, 4, 5]], np.float32), (20, 1)), ball_voxels(4) + 20])
        side, off = 32, 0
    t, o = gpu.Tree.build(vox), checker.tree(vox)
    assert flat_equal(t.export(), o.flatten()) == []
    n = 300000
    org, d = ds.random_rays(n, side=side)
    org += off
    d[:2000, 0] = 0                      # division by zero inside hit_box (geom.c:79-85)
    d[2000:4000, 1:] = 0                 # axis-aligned
    org[4000:8000] = rng.random((4000, 3)).astype(np.float32) * side + off   # origins inside the volume
    org[8000:10000] = np.floor(org[8000:10000])                               # exactly on voxel faces
    d[10000:11000] = np.round(d[10000:11000] / side * 4)                      # small integer directions
    assert_rays_equal(t.rays(org, d), o.rays(org, d, threads=THREADS), check_voxel=(checker.which == "port"))
    lo, hi = t.bounding_box()
    c, r = ds.random_balls(n, lo, hi)
    r[:1000] = 0                                                                # zero radius never collides (strict <)
    assert np.array_equal(t.balls(c, r), o.balls(c, r, threads=THREADS))


def test_rays_from_inner_and_leaf_nodes(gpu, checker, ds, skull):
    """The returned leaf must be usable as a tree (renderer.c:297); so must any node."""
    t, o = gpu.Tree.build(skull), checker.tree(skull)
    fl = t.export()
    org, d = ds.random_rays(20000, side=256)
    kinds = fl["flags"] & 3
    picks = list(np.flatnonzero(kinds == 1)[:20]) + list(np.flatnonzero(kinds == 0)[1:20]) + \
        list(np.flatnonzero(kinds == 2)[:10])
    for nid in picks:
        got = t.rays(org, d, start_node=int(nid))
        want = o.rays(org, d, start_node=int(nid), threads=THREADS)
        assert_rays_equal(got, want, check_voxel=False)


def test_nonunit_voxel(gpu, checker, ds):
    rng = np.random.default_rng(12)
    idx = np.unique(rng.integers(-30, 30, (20000, 3)), axis=0)
    rng.shuffle(idx)
    size = np.array([0.5, 2.0, 0.25], np.float32)
    vox = idx.astype(np.float32) * size
    try:
        gpu.set_voxel_size(*size)
        checker.set_voxel(*size)
        t, o = gpu.Tree.build(vox), checker.tree(vox)
        org = (rng.random((100000, 3)).astype(np.float32) - 0.5) * np.array([60, 240, 30], np.float32)
        d = rng.standard_normal((100000, 3)).astype(np.float32)
        assert_rays_equal(t.rays(org, d), o.rays(org, d, threads=THREADS), check_voxel=False)
        r = rng.random(100000).astype(np.float32) * 4
        assert np.array_equal(t.balls(org, r), o.balls(org, r, threads=THREADS))
        # the frame too: get_color divides by the voxel size (renderer.c:31-33)
        w, h = 320, 160
        cam = ds.setup_camera(gpu.Camera(), (5, -90, 3), (0, 0, 0), 1.0, w, h)
        ocam = ds.setup_camera(checker.camera(0), (5, -90, 3), (0, 0, 0), 1.0, w, h)
        ctx = gpu.Context(w, h)
        pix, leaf, pts = render_device(gpu, t, cam, w, h, ctx.texture(), ctx.pixfmt())
        opix, oleaf, ohit = o.render(ocam, w, h, 0, ctx.texture(), threads=THREADS)
        assert (oleaf >= 0).sum() > 1000
        assert np.array_equal(pix, opix) and np.array_equal(leaf, oleaf)
    finally:
        gpu.set_voxel_size(1, 1, 1)
        checker.set_voxel(1, 1, 1)


def test_single_query_api_and_edge_cases(gpu, checker, ds, skull):
    t, o = gpu.Tree.build(skull), checker.tree(skull)
    org, d = ds.random_rays(64, side=256)
    oleaf, ohit, _ = o.rays(org, d)
    for i in range(64):
        leaf, p = t.ray(org[i], d[i])
        if oleaf[i] < 0:
            assert leaf is None
        else:
            assert leaf == oleaf[i] and np.array_equal(bits(p), bits(ohit[i]))
    # the leaf handle returned by one search is accepted by the next
    i = int(np.flatnonzero(oleaf >= 0)[0])
    leaf, p = t.ray(org[i], d[i])
    leaf2, p2 = t.ray(org[i], d[i], start=t.node(leaf))
    assert leaf2 == leaf
    c, r = ds.random_balls(64, *t.bounding_box())
    ob = o.balls(c, r)
    assert [t.ball(c[k], float(r[k])) for k in range(64)] == list(ob)
    # empty batches, empty tree
    assert t.rays(np.zeros((0, 3)), np.zeros((0, 3)))[0].shape == (0,)
    assert t.balls(np.zeros((0, 3)), np.zeros(0)).shape == (0,)
    e = gpu.Tree.build(np.zeros((0, 3), np.float32))
    assert (e.rays(org, d)[0] == -1).all() and not e.balls(c, r).any()
    # NaN / inf inputs do not hang or crash and agree with the oracle's booleans
    bad_o, bad_d = org.copy(), d.copy()
    bad_d[:8] = np.nan; bad_o[8:16] = np.inf; bad_d[16:24] = 0
    got, want = t.rays(bad_o, bad_d), o.rays(bad_o, bad_d)
    assert np.array_equal(got[0], want[0])


def test_traversal_counters_match_oracle(gpu, port, ds, skull):
    """The instrumented kernel visits exactly the nodes the reference's recursion visits
    (this is what makes the oracle's visit counts the roofline's algorithmic bytes)."""
    t, o = gpu.Tree.build(skull), port.tree(skull)
    org, d = ds.random_rays(200000, side=256)
    *_, st = o.rays(org, d, threads=THREADS, want_stats=True)
    g = gpu.api.libs.gpu
    d_o, d_d = gpu.DeviceBuffer.from_array(gpu.api.dots4(org)), gpu.DeviceBuffer.from_array(gpu.api.dots4(d))
    gs = gpu.api.GpuStats()
    assert g.voxgpu_rays_stats(t.device_handle(), 0, d_o.ptr, d_d.ptr, len(org), None, None, None, C.byref(gs)) == 0
    gs = gs.as_dict()
    assert gs["rays"] == len(org)
    assert gs["calls"] == st["calls"] - st["null_calls"]
    assert (gs["inner"], gs["leaf"], gs["dense"]) == (st["inner"], st["leaf"], st["dense"])
    assert gs["dots_tested"] == st["dots_tested"]
    assert gs["expansions"] * 3 == st["plane_tests"]
    assert gs["max_stack"] <= 2 * t.info().depth


# ----------------------------------------------------------------- K1 render

@pytest.mark.parametrize("frame", [0, 45, 90, 135, 180, 225, 270, 315])
def test_c2_frames_golden(gpu, golden, skull, ds, frame):
    """Skull 1080p at full size against the reference-generated digests (SURVEY.md 8d known answers)."""
    meta, _ = golden
    cache = test_c2_frames_golden.__dict__
    t = cache.setdefault("tree", gpu.Tree.build(skull))
    ctx = cache.setdefault("ctx", gpu.Context(1920, 1080))
    assert sha(ctx.texture()) == meta["texture_sha"]
    pos, look = ds.orbit_pose(frame)
    cam = ds.setup_camera(gpu.Camera("simple-camera"), pos, look, 1.2, 1920, 1080)
    want = meta["c2"][str(frame)]
    pix, leaf, pts = render_device(gpu, t, cam, 1920, 1080, ctx.texture(), ctx.pixfmt())
    m = leaf >= 0
    assert int(m.sum()) == want["hits"] == ds.C2_KNOWN_HITS[frame]
    assert sha(leaf) == want["leaf_sha"]
    assert sha(np.ascontiguousarray(pts[m])) == want["hit_sha"]
    assert sha(pix) == want["pix_sha"]
    # the reference-facing path: vox_render + vox_redraw into the (pinned) surface
    ctx.set_scene(t); ctx.set_camera(cam)
    assert ctx.set_quality(gpu.QUALITY_BEST)
    ctx.pixels()[:] = 0xdeadbeef
    ctx.render(); ctx.redraw()
    assert gpu.last_error() == ""
    assert sha(ctx.pixels()) == want["pix_sha"]
    assert ctx.kernel_ms() > 0


@pytest.mark.parametrize("name,kind,pos", [("simple-camera", 0, (128, 128, 400)), ("doom-camera", 1, (100, -250, 180))])
def test_c1_bonsai_both_cameras(gpu, checker, golden, bonsai, ds, name, kind, pos):
    meta, _ = golden
    t = gpu.Tree.build(bonsai)
    ctx = gpu.Context(800, 600)
    cam = ds.setup_camera(gpu.Camera(name), pos, (128, 128, 128), 1.2, 800, 600)
    pix, leaf, pts = render_device(gpu, t, cam, 800, 600, ctx.texture(), ctx.pixfmt())
    want = meta["c1"]["simple" if kind == 0 else "doom"]
    m = leaf >= 0
    assert int(m.sum()) == want["hits"] and sha(leaf) == want["leaf_sha"]
    assert sha(np.ascontiguousarray(pts[m])) == want["hit_sha"] and sha(pix) == want["pix_sha"]
    # a moved / rotated camera, against the live oracle
    ocam = ds.setup_camera(checker.camera(kind), pos, (128, 128, 128), 1.2, 800, 600)
    for c in (cam, ocam):
        c.rotate([0.05, -0.1, 0.2]); c.move([3, -2, 1])
    pix, leaf, pts = render_device(gpu, t, cam, 800, 600, ctx.texture(), ctx.pixfmt())
    opix, oleaf, ohit = checker.tree(bonsai).render(ocam, 800, 600, 0, ctx.texture(), threads=THREADS)
    assert np.array_equal(leaf, oleaf) and np.array_equal(pix, opix)
    assert np.array_equal(bits(pts[oleaf >= 0]), bits(ohit[oleaf >= 0]))


QUALITIES = {"fast": 1, "adaptive": 2, "adaptive+merge": 2 | 4, "adaptive+merge_accurate": 2 | 8}


@pytest.mark.parametrize("qname", list(QUALITIES))
@pytest.mark.parametrize("scene", ["bonsai", "skull_far"])
def test_quality_modes_match_the_reference_block_loop(gpu, checker, golden, bonsai, skull, ds, qname, scene):
    """FAST / ADAPTIVE / ray merge (renderer.c:240-316): the reference's default modes, in which the
    pixels of a 4x4 block depend on each other.  Pixels, leaf ids and hit points of every written
    pixel must equal the reference's block loop; unwritten pixels stay 0 / -1."""
    q = QUALITIES[qname]
    if scene == "bonsai":
        vox, w, h, pos, look = bonsai, 800, 600, (128, 128, 400), (128, 128, 128)
    else:       # far enough for the MAX_DIST = 150 ray-merge criterion to fire (renderer.c:198,274)
        vox, w, h, pos, look = skull, 640, 360, (128, 128, -330), (128, 128, 128)
    t = gpu.Tree.build(vox)
    ctx = gpu.Context(w, h)
    cam = ds.setup_camera(gpu.Camera(), pos, look, 1.2, w, h)
    ocam = ds.setup_camera(checker.camera(0), pos, look, 1.2, w, h)
    pix, leaf, pts = render_device(gpu, t, cam, w, h, ctx.texture(), ctx.pixfmt(), quality=q)
    opix, oleaf, ohit = checker.tree(vox).render(ocam, w, h, q, ctx.texture(), threads=THREADS)
    assert (opix != 0).sum() > 10000
    assert np.array_equal(leaf, oleaf), f"{int((leaf != oleaf).sum())} leaf ids differ"
    assert np.array_equal(pix, opix), f"{int((pix != opix).sum())} pixels differ"
    m = oleaf >= 0
    assert np.array_equal(bits(pts[m]), bits(ohit[m]))
    if scene == "bonsai" and qname in meta_modes(golden):
        assert sha(pix) == meta_modes(golden)[qname]
    # row bands (the presenter, multi-GPU) give the same pixels
    band = render_device(gpu, t, cam, w, h, ctx.texture(), ctx.pixfmt(), 120, 240, quality=q)
    assert np.array_equal(band[0], pix[120:240]) and np.array_equal(band[1], leaf[120:240])
    # and so does the reference-facing API, whose contexts default to ADAPTIVE (renderer.c:78)
    ctx.set_scene(t); ctx.set_camera(cam)
    assert ctx.set_quality(q)
    ctx.render(); ctx.redraw()
    assert gpu.last_error() == "" and np.array_equal(ctx.pixels(), opix)


def meta_modes(golden):
    meta, _ = golden
    return meta.get("c1", {}).get("modes", {})


def test_default_quality_is_adaptive_like_the_reference(gpu, checker, skull, ds):
    t = gpu.Tree.build(skull)
    w, h = 320, 160
    pos, look = ds.orbit_pose(77)
    cam = ds.setup_camera(gpu.Camera(), pos, look, 1.2, w, h)
    ocam = ds.setup_camera(checker.camera(0), pos, look, 1.2, w, h)
    ctx = gpu.Context(w, h)                      # quality untouched
    ctx.set_scene(t); ctx.set_camera(cam)
    ctx.render(); ctx.redraw()
    opix = checker.tree(skull).render(ocam, w, h, 2, ctx.texture(), threads=THREADS, want_hits=False)
    assert np.array_equal(ctx.pixels(), opix)
    assert not ctx.set_quality(1 | 4) and not ctx.set_quality(3)     # renderer.c:179-195


def test_row_bands_tile_the_frame(gpu, skull, ds):
    """Multi-GPU partition: any 4-aligned row band equals the same rows of the full frame."""
    t = gpu.Tree.build(skull)
    w, h = 640, 360
    ctx = gpu.Context(w, h)
    pos, look = ds.orbit_pose(30)
    cam = ds.setup_camera(gpu.Camera(), pos, look, 1.2, w, h)
    full = render_device(gpu, t, cam, w, h, ctx.texture(), ctx.pixfmt())
    for y0, y1 in ((0, 92), (92, 180), (180, 360), (356, 360)):
        band = render_device(gpu, t, cam, w, h, ctx.texture(), ctx.pixfmt(), y0, y1)
        assert np.array_equal(band[0], full[0][y0:y1]) and np.array_equal(band[1], full[1][y0:y1])
    api = gpu.api
    fr = api.GpuFrame(w, h, 2, 10, w, 0, ctx.pixfmt())
    gc = cam.gpu_camera()
    d = gpu.DeviceBuffer(w * h * 4)
    assert api.libs.gpu.voxgpu_render(t.device_handle(), C.byref(gc), C.byref(fr), d.ptr, d.ptr, None, None, None) != 0
    assert b"row band" in api.libs.gpu.voxgpu_last_error()


def test_interleaved_tile_rows_tile_the_frame(gpu, skull, ds):
    """voxgpu_frame.interleave_*: N launches (one per GPU in the multi-GPU drivers), each drawing every
    N-th 4-row tile row into the SAME buffer, give the frame one launch gives -- and a launch touches
    nothing but its own rows."""
    t = gpu.Tree.build(skull)
    w, h = 640, 360
    ctx = gpu.Context(w, h)
    cam = ds.setup_camera(gpu.Camera(), *ds.orbit_pose(200), 1.2, w, h)
    gc = cam.gpu_camera()
    api = gpu.api
    g = api.libs.gpu
    d_tex = gpu.DeviceBuffer.from_array(ctx.texture())
    full = render_device(gpu, t, cam, w, h, ctx.texture(), ctx.pixfmt(), want_hits=False)[0]
    for world in (2, 3, 8):
        d_pix = gpu.DeviceBuffer(w * h * 4)
        g.voxgpu_memset(d_pix.ptr, 0x5a, w * h * 4, None)
        for rank in range(world):
            fr = api.GpuFrame(w, h, 0, h, w, 0, ctx.pixfmt(), None, world, rank)
            assert g.voxgpu_render(t.device_handle(), C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, None, None, None) == 0
            if rank == 0:
                part = d_pix.download(np.uint32, (h, w))
                rows = (np.arange(h) // 4) % world
                assert np.array_equal(part[rows == 0], full[rows == 0])
                assert (part[rows != 0] == 0x5a5a5a5a).all()
        assert np.array_equal(d_pix.download(np.uint32, (h, w)), full)
    # a band [y0, y1) interleaved: rows are counted from y0
    d_pix = gpu.DeviceBuffer(w * 120 * 4)
    for rank in range(2):
        fr = api.GpuFrame(w, h, 120, 240, w, 0, ctx.pixfmt(), None, 2, rank)
        assert g.voxgpu_render(t.device_handle(), C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, None, None, None) == 0
    assert np.array_equal(d_pix.download(np.uint32, (120, w)), full[120:240])
    fr = api.GpuFrame(w, h, 0, h, w, 2, ctx.pixfmt(), None, 2, 0)       # not with the block-sequential modes
    assert g.voxgpu_render(t.device_handle(), C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, None, None, None) != 0


@pytest.mark.parametrize("camera", ["simple-camera", "doom-camera"])
def test_root_box_culling_changes_nothing(gpu, checker, skull, ds, camera):
    """Tiles outside the screen bounds of the root box are cleared without casting their rays
    (screen_bounds in voxgpu.cu).  Random poses -- far, near, inside the volume, rolled, looking away,
    narrow and wide fov: the frame, the leaf ids and the hit points must not depend on the switch, and
    must equal the oracle's."""
    t = gpu.Tree.build(skull)
    otree = checker.tree(skull)
    w, h = 320, 160
    ctx = gpu.Context(w, h)
    g = gpu.api.libs.gpu
    g.voxgpu_set_cull.argtypes = [C.c_int]
    rng = np.random.default_rng(11)
    kind = 0 if camera == "simple-camera" else 1
    culled_something = 0
    try:
        for trial in range(24):
            dist = [600.0, 330.0, 200.0, 140.0, 60.0, 0.0][trial % 6]
            v = rng.normal(size=3)
            pos = 128 + dist * v / np.linalg.norm(v) + rng.normal(size=3) * 3
            look = 128 + rng.normal(size=3) * (40 if trial % 4 else 400)
            fov = [0.3, 1.2, 2.5][trial % 3]
            cam, ocam = gpu.Camera(camera), checker.camera(kind)
            rot = [float(x) for x in rng.normal(size=3) * (0.5 if trial % 2 else 0.0)]
            for c_ in (cam, ocam):
                ds.setup_camera(c_, pos, look, fov, w, h)
                c_.rotate(rot)
            g.voxgpu_set_cull(2)                         # root-box bounds only
            mid = render_device(gpu, t, cam, w, h, ctx.texture(), ctx.pixfmt())
            g.voxgpu_set_cull(1)                         # + the screen mask from the tree's cut
            on = render_device(gpu, t, cam, w, h, ctx.texture(), ctx.pixfmt())
            assert np.array_equal(on[0], mid[0]) and np.array_equal(on[1], mid[1]), trial
            g.voxgpu_set_cull(0)
            off = render_device(gpu, t, cam, w, h, ctx.texture(), ctx.pixfmt())
            assert np.array_equal(on[0], off[0]) and np.array_equal(on[1], off[1]), trial
            m = off[1] >= 0
            assert np.array_equal(bits(on[2][m]), bits(off[2][m])), trial
            opix, oleaf, _ = otree.render(ocam, w, h, 0, ctx.texture(), threads=THREADS)
            assert np.array_equal(on[0], opix) and np.array_equal(on[1], oleaf), trial
            culled_something += int((~m).any())
    finally:
        g.voxgpu_set_cull(1)
    assert culled_something > 5


def test_many_launches_in_flight_on_three_streams(gpu, skull, ds):
    """More launches queued than there are render slots (16 pairs of tile counters + screen masks), on a
    number of streams that does not divide 64: a slot that comes round again while its previous borrower
    is still running on another stream must wait for it (slot_acquire / slot_release)."""
    t = gpu.Tree.build(skull)
    w, h = 640, 360
    ctx = gpu.Context(w, h)
    api = gpu.api
    g = api.libs.gpu
    d_tex = gpu.DeviceBuffer.from_array(ctx.texture())
    cams, want = [], []
    cam = gpu.Camera()
    for f in (10, 100, 190, 280):
        ds.setup_camera(cam, *ds.orbit_pose(f), 1.2, w, h)
        cams.append(cam.gpu_camera())
        want.append(render_device(gpu, t, cam, w, h, ctx.texture(), ctx.pixfmt(), want_hits=False)[0])
    n = 200
    bufs = [gpu.DeviceBuffer(w * h * 4) for _ in range(n)]
    streams = [C.c_void_p(g.voxgpu_stream_create()) for _ in range(3)]
    fr = api.GpuFrame(w, h, 0, h, w, 0, ctx.pixfmt(), None)
    for i in range(n):
        assert g.voxgpu_render(t.device_handle(), C.byref(cams[i % 4]), C.byref(fr), d_tex.ptr, bufs[i].ptr, None, None,
                               streams[i % 3]) == 0
    g.voxgpu_synchronize(None)
    for st in streams:
        g.voxgpu_synchronize(st)
    bad = [i for i in range(n) if not np.array_equal(bufs[i].download(np.uint32, (h, w)), want[i % 4])]
    assert bad == []
    for st in streams:
        g.voxgpu_stream_destroy(st)


def test_c_example_matches_the_oracle(gpu, checker, ds, tmp_path):
    """examples/orbit.c -- a C program written against the reference's API only -- renders the frames the
    oracle renders for the same scene and poses (FNV-1a of the surface bytes, lit pixel count)."""
    import math
    import subprocess
    from test_host import build_c_example
    exe = build_c_example(tmp_path)
    out = subprocess.run([exe, "3"], capture_output=True, text=True, check=True).stdout.split("\n")
    otree = checker.tree(ball_voxels(20))
    tex = gpu.Context(16, 4).texture()
    for f in range(3):
        th = f * 2.0 * math.pi / 16
        pos = (np.float32(70.0 * math.sin(th)), np.float32(-70.0 * math.cos(th)), np.float32(10.0))
        ocam = ds.setup_camera(checker.camera(0), pos, (0, 0, 0), 1.0, 256, 128)
        pix = otree.render(ocam, 256, 128, 0, tex, want_hits=False)
        hsh = 1469598103934665603
        for b in pix.astype("<u4").tobytes():
            hsh = ((hsh ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        assert out[f] == f"frame {f} lit {int((pix != 0).sum())} hash {hsh:016x}", out[f]


def test_foreign_camera_without_describe(gpu, checker, skull, ds):
    """A camera written against the reference's 10-slot vtable (no `describe`) still renders:
    its host screen2world is tabulated."""
    t = gpu.Tree.build(skull)
    w, h = 320, 160
    pos, look = ds.orbit_pose(200)
    cam = ds.setup_camera(gpu.Camera(), pos, look, 1.2, w, h)
    ctx = gpu.Context(w, h)
    ctx.set_scene(t); ctx.set_camera(cam); ctx.set_quality(0)
    ctx.render(); ctx.redraw()
    want = ctx.pixels().copy()
    assert (want != 0).sum() > 1000
    null_describe = type(cam.iface.describe)()          # NULL function pointer
    cam.obj.contents.iface.contents.describe = null_describe
    ctx.pixels()[:] = 1
    ctx.render(); ctx.redraw()
    assert gpu.last_error() == ""
    assert np.array_equal(ctx.pixels(), want)


def test_pixel_formats_and_texture(gpu, checker, skull, ds, orc):
    t, o = gpu.Tree.build(skull), checker.tree(skull)
    w, h = 320, 160
    pos, look = ds.orbit_pose(100)
    cam = ds.setup_camera(gpu.Camera(), pos, look, 1.2, w, h)
    ocam = ds.setup_camera(checker.camera(0), pos, look, 1.2, w, h)
    tex = np.random.default_rng(1).integers(0, 256, 32768).astype(np.uint8)
    for amask, loss, shift in ((0, (0, 0, 0), (16, 8, 0)), (0xff000000, (0, 0, 0), (0, 8, 16)), (0, (3, 2, 3), (11, 5, 0))):
        fmt = gpu.api.GpuPixFmt(amask, *loss, 0, *shift, 0)
        ofmt = orc.PixFmt(amask, *loss, 0, *shift, 0)
        pix, _, _ = render_device(gpu, t, cam, w, h, tex, fmt)
        opix = o.render(ocam, w, h, 0, tex, fmt=ofmt, want_hits=False)
        assert np.array_equal(pix, opix)


def test_collision_detector(gpu, skull):
    """vox_cd_collide (cd.c:38-55): a move into the skull is undone, a move in free space is kept."""
    t = gpu.Tree.build(skull)
    ctx = gpu.Context(64, 16)
    cam = gpu.Camera()
    cam.set_position([128, 128, -200])
    ctx.set_scene(t); ctx.set_camera(cam)
    r = gpu.libs.rnd
    cd = r.vox_make_cd()
    r.vox_cd_attach_camera(cd, cam.h, 3.0)
    r.vox_cd_attach_context(cd, ctx.h)
    cam.set_position([128, 128, -150]); r.vox_cd_collide(cd)
    assert np.array_equal(cam.get_position(), [128, 128, -150])
    inside = t.export()["dots"][1000]
    assert t.ball(inside, 3.0) == 1
    cam.set_position(inside); r.vox_cd_collide(cd)
    assert np.array_equal(cam.get_position(), [128, 128, -150])
    r.vox_destroy_cd(cd)


# ------------------------------------------------- online edits (f-3): the device copy follows

def test_edited_tree_on_device(gpu, checker, golden, ds):
    """vox_insert_voxel / vox_delete_voxel between queries: every burst of edits is followed by
    ray and ball batches that must see the edited tree (re-flattened and re-uploaded lazily)."""
    meta, arr = golden
    vox, ins, pts = ds.edit_script()
    t = gpu.Tree.build(vox)
    o, d = ds.random_rays(4096, side=48)
    o -= 24
    assert_rays_equal(t.rays(o, d), checker.tree(vox).rays(o, d))     # device copy of the unedited tree exists now
    for lo in range(0, len(ins), 750):
        t.edit(ins[lo:lo + 750], pts[lo:lo + 750])
        ot = checker.tree_from_flat(t.export())
        assert_rays_equal(t.rays(o, d), ot.rays(o, d))
        c, r = ds.random_balls(2048, *t.bounding_box())
        assert np.array_equal(t.balls(c, r), ot.balls(c, r))
    leaf, hit, _ = t.rays(o, d)
    assert np.array_equal(leaf, arr["edit_rays_leaf"])                 # the reference's own edited tree, via golden
    m = leaf >= 0
    assert int(m.sum()) == meta["edits"]["ray_hits"]
    assert np.array_equal(bits(hit[m]), bits(arr["edit_rays_hit"][m]))
    # single-query API on the edited tree, including a search restarted inside a returned leaf
    i = int(np.flatnonzero(m)[0])
    lid, p = t.ray(o[i], d[i])
    assert lid == leaf[i] and np.array_equal(bits(p), bits(hit[i]))
    lid2, p2 = t.ray(o[i], d[i], start=t.node(lid))
    assert lid2 == lid


def test_render_follows_edits(gpu, checker, ds):
    """A context keeps its scene pointer across edits (the root handle is stable): the next
    vox_render shows the edited tree; then vox_rebuild_tree's tree as a new scene."""
    vox = ball_voxels(20)
    t = gpu.Tree.build(vox)
    w, h = 256, 128
    cam = ds.setup_camera(gpu.Camera("simple-camera"), (5, -70, 10), (0, 0, 0), 1.0, w, h)
    ctx = gpu.Context(w, h)
    ctx.set_scene(t)
    ctx.set_camera(cam)
    ctx.set_quality(gpu.QUALITY_BEST)

    def frame():
        ctx.render()
        ctx.redraw()
        assert gpu.last_error() == ""
        return ctx.pixels().copy()

    def oracle_frame(tree):
        ocam = ds.setup_camera(checker.camera(0), (5, -70, 10), (0, 0, 0), 1.0, w, h)
        return checker.tree_from_flat(tree.export()).render(ocam, w, h, 0, ctx.texture(), want_hits=False)

    p0 = frame()
    assert np.array_equal(p0, oracle_frame(t))
    # drill a tunnel along the view axis and float a slab in front of the ball
    g = np.arange(-2, 3)
    tunnel = np.stack(np.meshgrid(g, np.arange(-20, 20), g + 2, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    slab = np.stack(np.meshgrid(np.arange(-12, -4), [-30.0], np.arange(-3, 9), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    assert t.delete(tunnel).sum() > 500
    assert t.insert(slab).sum() == len(slab)
    p1 = frame()
    assert not np.array_equal(p0, p1)
    assert np.array_equal(p1, oracle_frame(t))
    rb = t.rebuild()
    ctx.set_scene(rb)
    p2 = frame()
    assert np.array_equal(p2, oracle_frame(rb))
    assert rb.voxels() == t.voxels()                   # same voxels, differently shaped tree


# ------------------------------------------------- lights (f-4): shading through the flattened M-tree

@pytest.mark.parametrize("qname,q", [("best", 0), ("adaptive", 2), ("adaptive+merge", 2 | 4)])
def test_lit_frames(gpu, checker, golden, bonsai, ds, qname, q):
    """vox_render with a light manager attached (renderer.c:41-49, lights.c:91-116): per-channel
    min(1, ambient + lights) * texture, the lights summed in M-tree order.  Pixels equal the reference's
    (golden, made with its lights.c / mtree.c object code) and the live oracle's evaluation of the same
    flat light set; through vox_render/vox_redraw and through voxgpu_render with row bands."""
    meta, _ = golden
    t = gpu.Tree.build(bonsai)
    w, h = 800, 600
    lm = gpu.LightManager()
    ds.apply_lights(lm, ds.light_script())
    cam = ds.setup_camera(gpu.Camera(), (128, 128, 400), (128, 128, 128), 1.2, w, h)
    ctx = gpu.Context(w, h)
    ctx.set_scene(t); ctx.set_camera(cam); ctx.set_light_manager(lm)
    assert ctx.set_quality(q)
    ctx.render(); ctx.redraw()
    assert gpu.last_error() == ""
    pix = ctx.pixels().copy()
    want = meta["lights"]["c1_frames"][qname]
    assert int((pix != 0).sum()) == want["lit"]
    assert sha(pix) == want["pix_sha"]
    ocam = ds.setup_camera(checker.camera(0), (128, 128, 400), (128, 128, 128), 1.2, w, h)
    olights = checker.lights(lm.flatten())
    opix = checker.tree(bonsai).render_lit(ocam, w, h, olights, q, ctx.texture(), threads=THREADS)
    assert np.array_equal(pix, opix)
    # coloured, not grey: the three channels differ somewhere
    assert ((pix & 0xff) != ((pix >> 8) & 0xff)).any()
    # the device entry point with an explicit light handle, one row band
    api = gpu.api
    dl = gpu.upload_lights(lm.flatten())
    d_tex = gpu.DeviceBuffer.from_array(ctx.texture())
    d_pix = gpu.DeviceBuffer(w * 120 * 4)
    fr = api.GpuFrame(w, h, 240, 360, w, q, ctx.pixfmt(), dl)
    gc = cam.gpu_camera()
    assert api.libs.gpu.voxgpu_render(t.device_handle(), C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, None, None, None) == 0
    assert np.array_equal(d_pix.download(np.uint32, (120, w)), pix[240:360])
    api.libs.gpu.voxgpu_synchronize(None)
    api.libs.gpu.voxgpu_lights_free(dl)


def test_lights_follow_changes(gpu, checker, skull, ds):
    """The device copy of the lights is refreshed after every change of the manager; detaching the
    manager gives the unlit frame again; a manager without lights applies the ambient colour only."""
    t = gpu.Tree.build(skull)
    w, h = 320, 160
    pos, look = ds.orbit_pose(30)
    cam = ds.setup_camera(gpu.Camera(), pos, look, 1.2, w, h)
    ocam = ds.setup_camera(checker.camera(0), pos, look, 1.2, w, h)
    otree = checker.tree(skull)
    ctx = gpu.Context(w, h)
    ctx.set_scene(t); ctx.set_camera(cam); ctx.set_quality(0)

    def frame():
        ctx.render(); ctx.redraw()
        assert gpu.last_error() == ""
        return ctx.pixels().copy()

    def oracle(lm):
        return otree.render_lit(ocam, w, h, checker.lights(lm.flatten()) if lm else None, 0, ctx.texture(), threads=THREADS)

    unlit = frame()
    assert np.array_equal(unlit, oracle(None))
    lm = gpu.LightManager()
    ctx.set_light_manager(lm)
    assert np.array_equal(frame(), unlit)                       # ambient (1,1,1), no lights: same as unlit
    lm.set_ambient([0.2, 0.1, 0.05])
    dim = frame()
    assert np.array_equal(dim, oracle(lm)) and not np.array_equal(dim, unlit)
    lm.insert([128, 60, 128], 120.0, [1.0, 0.6, 0.2])
    lm.insert([60, 128, 200], 90.0, [0.1, 0.9, 1.0])
    two = frame()
    assert np.array_equal(two, oracle(lm)) and not np.array_equal(two, dim)
    lm.delete([128, 60, 128], 120.0)
    one = frame()
    assert np.array_equal(one, oracle(lm)) and not np.array_equal(one, two)
    ctx.set_light_manager(None)
    assert np.array_equal(frame(), unlit)


# ------------------------------------------------- full-size configs, by properties

def test_c3_full_size_properties(gpu, port, ds):
    """C3: 16 Mi rays against the ~10 M-voxel random tree: known voxel count and tree shape, size-independent
    properties on the full batch (hit point on the hit voxel; leaf id is a leaf; determinism; the order the
    batch is searched in changes nothing) and EVERY ray against the oracle, bit for bit (the CPU search does
    ~10 Mrays/s on this tree with all host threads: seconds, not minutes)."""
    vox, _ = ds.random_voxels()
    assert len(vox) == 10063815
    t = gpu.Tree.build(vox)
    ti = t.info()
    assert (ti.n_inner, ti.n_leaf, ti.n_dense) == (479723, 2096194, 883177)
    n = 1 << 24
    o, d = ds.random_rays(n, side=512, start_index=512 ** 3)
    leaf, pts, voxel = t.rays(o, d)
    fl = t.export()
    m = leaf >= 0
    assert 0.2 < m.mean() <= 1.0
    kinds = fl["flags"][leaf[m]] & 3
    assert (kinds != 0).all()
    lm = m & (voxel >= 0)
    lo = fl["dots"][voxel[lm]]
    p = pts[lm]
    assert ((p >= lo) & (p <= lo + 1)).all()        # the hit point lies on the hit voxel
    dm = m & (voxel < 0)
    assert ((pts[dm] >= fl["bb_min"][leaf[dm]]) & (pts[dm] <= fl["bb_max"][leaf[dm]])).all()
    g = gpu.api.libs.gpu
    try:
        g.voxgpu_set_rays_sort(0)                   # as given (the full batch above went in entry order)
        leaf2, pts2, _ = t.rays(o[: 1 << 21], d[: 1 << 21])
    finally:
        g.voxgpu_set_rays_sort(-1)
    assert np.array_equal(leaf2, leaf[: 1 << 21]) and np.array_equal(bits(pts2), bits(pts[: 1 << 21]))
    # the whole batch against the oracle (the port builds the same tree in a few seconds)
    ot = port.tree(vox)
    assert flat_equal(fl, ot.flatten()) == []
    assert_rays_equal((leaf, pts, voxel), ot.rays(o, d, threads=THREADS))


def test_c4_full_size(gpu, checker, ds, skull):
    """C4: 4 Mi balls vs the skull tree, whole batch against the oracle (booleans bit-exact)."""
    t, o = gpu.Tree.build(skull), checker.tree(skull)
    c, r = ds.random_balls(1 << 22, *t.bounding_box())
    got = t.balls(c, r)
    assert np.array_equal(got, o.balls(c, r, threads=THREADS))
    # monotone in the radius: growing a colliding ball keeps it colliding
    assert (t.balls(c[:100000], r[:100000] * 2) >= got[:100000]).all()
