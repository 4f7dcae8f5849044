"""Host-side logic of the product, on CPU: the C-ABI libraries load and export
everything include/*.h declares, the flattened tree builder equals the
reference's tree, cameras/contexts behave like the reference's, and -- with
no GPU -- every search fails loudly instead of falling back."""
import ctypes as C
import math
import os
import re
import subprocess

import numpy as np
import pytest

from util import CROSS, ball_voxels, bits, check_tree_invariants, flat_equal, sha

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(vv):
    """Every VOX_EXPORT / VOXGPU_EXPORT declaration in include/."""
    names = set()
    for dirpath, _, files in os.walk(vv.INCLUDE_DIR):
        for f in files:
            if not f.endswith(".h"):
                continue
            text = open(os.path.join(dirpath, f)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            text = re.sub(r"^\s*#.*$", "", text, flags=re.M)      # preprocessor lines
            for m in re.finditer(r"VOX(?:GPU)?_EXPORT\s+([^;{]*?);", text, flags=re.S):
                decl = " ".join(m.group(1).split())
                fn = re.search(r"(\w+)\s*\(", decl)
                names.add(fn.group(1) if fn else re.search(r"(\w+)\s*$", decl).group(1))
    return names


def test_libraries_export_every_declared_symbol(vv):
    names = declared_symbols(vv)
    assert {"vox_make_tree", "vox_ray_tree_intersection", "vox_tree_ball_collidep", "vox_render",
            "vox_redraw", "vox_voxel", "voxgpu_render", "voxgpu_rays_device", "voxgpu_balls_device"} <= names
    exported = set()
    for lib in ("libvoxgpu.so", "libvoxtrees.so", "libvoxrnd.so"):
        out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(vv.LIB_DIR, lib)],
                             capture_output=True, text=True, check=True).stdout
        exported |= {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = sorted(names - exported)
    assert not missing, f"declared in include/ but not exported: {missing}"
    # and they resolve through dlopen as well
    for n in ("vox_make_tree", "vox_ray_tree_intersection_batch", "vox_tree_ball_collidep_batch"):
        assert getattr(vv.libs.trees, n)
    for n in ("vox_render", "vox_redraw", "vox_camera_methods", "vox_cd_collide"):
        assert getattr(vv.libs.rnd, n)
    for lib in ("simple-camera.so", "doom-camera.so"):
        out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(vv.LIB_DIR, lib)],
                             capture_output=True, text=True, check=True).stdout
        assert " module_init" in out


def test_gpu_code_is_sm100a(vv):
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(vv.LIB_DIR, "libvoxgpu.so")],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_builder_equals_oracle(vv, port, skull, bonsai):
    rng = np.random.default_rng(5)
    sparse = np.unique(rng.integers(0, 100, (60000, 3)), axis=0).astype(np.float32)
    rng.shuffle(sparse)
    cases = {"skull": skull, "bonsai": bonsai, "ball50": ball_voxels(50), "sparse": sparse, "cross": CROSS,
             "one": CROSS[:1], "seven": CROSS[:7], "eight": CROSS[:8], "line": CROSS[5:13]}
    for name, vox in cases.items():
        t, o = vv.Tree.build(vox), port.tree(vox)
        assert flat_equal(t.export(), o.flatten()) == [], name
        assert np.array_equal(t.reordered[:, :3], o.reordered_set()), name
        assert t.voxels() == len(vox)
        fl = t.export()
        lo, hi = t.bounding_box()
        assert np.array_equal(lo, fl["bb_min"][0]) and np.array_equal(hi, fl["bb_max"][0])


def test_builder_golden_hashes(vv, golden, skull):
    """Product builder vs the hashes of the reference's own tree (no oracle library involved)."""
    meta, _ = golden
    t = vv.Tree.build(skull)
    want = meta["trees"]["skull"]
    for k, v in t.export().items():
        assert (sha(v) if isinstance(v, np.ndarray) else v) == want[k], k
    assert sha(np.ascontiguousarray(t.reordered[:, :3])) == want["reordered_sha"]
    assert check_tree_invariants(vv.Tree.build(ball_voxels(20)).export()) > 100


def test_builder_overflow_leaf_and_empty(vv, port):
    # identical points cannot be separated: one overflowed leaf (tree.c:281-305)
    dup = np.tile(np.array([[3, 4, 5]], np.float32), (20, 1))
    t, o = vv.Tree.build(dup), port.tree(dup)
    fl = t.export()
    assert flat_equal(fl, o.flatten()) == []
    assert fl["flags"][0] == 5 and fl["dots_num"][0] == 20
    # a cluster of duplicates next to ordinary voxels
    mix = np.concatenate([dup, ball_voxels(4) + 20]).astype(np.float32)
    assert flat_equal(vv.Tree.build(mix).export(), port.tree(mix).flatten()) == []
    # n == 0 is the NULL tree (tree.c:251-254)
    e = vv.Tree.build(np.zeros((0, 3), np.float32))
    assert e.h is None and e.voxels() == 0


def test_builder_nonunit_voxel(vv, port):
    rng = np.random.default_rng(12)
    idx = np.unique(rng.integers(-30, 30, (20000, 3)), axis=0)
    rng.shuffle(idx)
    size = np.array([0.5, 2.0, 0.25], np.float32)
    vox = idx.astype(np.float32) * size
    try:
        vv.set_voxel_size(*size)
        port.set_voxel(*size)
        assert np.array_equal(vv.get_voxel_size(), size)
        assert flat_equal(vv.Tree.build(vox).export(), port.tree(vox).flatten()) == []
    finally:
        vv.set_voxel_size(1, 1, 1)
        port.set_voxel(1, 1, 1)


def test_import_export_roundtrip_and_validation(vv, port):
    fl = port.tree(ball_voxels(10)).flatten()
    t = vv.Tree.from_arrays(fl)
    assert flat_equal(t.export(), fl) == []
    bad = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in fl.items()}
    bad["first_child"][0] += 1
    with pytest.raises(vv.VoxError):
        vv.Tree.from_arrays(bad)


def test_dense_leaf_tree(vv):
    box = vv.api._Box()
    for i, (a, b) in enumerate(((1, 5), (2, 4), (0, 8))):
        box.min[i], box.max[i] = a, b
    h = vv.libs.trees.vox_make_dense_leaf(C.byref(box))
    t = vv.Tree(h)
    assert t.voxels() == 4 * 2 * 8 and t.export()["flags"][0] == 2


@pytest.mark.parametrize("name,kind", [("simple-camera", 0), ("doom-camera", 1)])
def test_cameras_reference_known_answers_and_oracle(vv, port, name, kind):
    """tests/tests.c:661-777 on the product cameras, then state parity with the oracle."""
    cam = vv.Camera(name)
    cam.set_position([0, 0, 0])
    cam.set_window_size(100, 100)
    assert np.allclose(cam.screen2world(50, 50), [0, 1, 0], atol=1e-5)
    cam.set_rotation([0, 0, math.pi / 2])
    assert np.allclose(cam.screen2world(50, 50), [-1, 0, 0], atol=1e-5)
    cam.rotate([0, math.pi / 2, 0])
    assert np.allclose(cam.screen2world(50, 50), [-1, 0, 0], atol=1e-5)
    cam.move([0, 1, 0])
    assert np.allclose(cam.get_position(), [-1, 0, 0], atol=1e-5)

    a, b = vv.Camera(name), port.camera(kind)
    for c in (a, b):
        c.set_position([100, -250, 180]); c.set_fov(1.2); c.look_at([128, 128, 128]); c.set_window_size(1920, 1080)
        c.rotate([0.1, 0.2, 0.3]); c.move([1, 2, 3])
    sa, sb = a.state(), b.state()
    sb[10] = 0                      # fov is not part of the device description
    if kind == 1:
        sb[6] = 0                   # nor is phi (sin/cos are)
    assert np.array_equal(bits(sa), bits(sb))
    for sx, sy in ((0, 0), (100, 200), (1919, 1079), (960, 540)):
        assert np.array_equal(bits(a.screen2world(sx, sy)), bits(b.screen2world(sx, sy)))
    # construct_camera(copy) keeps the state (tests.c:765-777)
    h2 = a.iface.construct_camera(a.h)
    pos = vv.api.dots4([[0, 0, 0]])
    C.cast(h2, C.POINTER(vv.api._CameraObj)).contents.iface.contents.get_position(h2, pos.ctypes.data)
    assert np.array_equal(pos[0, :3], a.get_position())


def test_quaternion_known_answer(vv):
    """tests/tests.c:302-358 quat_mul: (4,1,2,3)*(2,0,2,0) = (4,-4,12,8); rotation keeps the norm."""
    q1, q2, res = vv.api.dots4([[0, 0, 0]]), vv.api.dots4([[0, 0, 0]]), vv.api.dots4([[0, 0, 0]])
    q1[0] = (4, 1, 2, 3); q2[0] = (2, 0, 2, 0)
    vv.libs.rnd.vox_quat_mul(q1.ctypes.data, q2.ctypes.data, res.ctypes.data)
    assert np.array_equal(res[0], [4, -4, 12, 8])
    base, v = vv.api.dots4([[0, 0, 0]]), vv.api.dots4([[1, 5, 2]])
    base[0] = (0.5, 0.5, 0.5, 0.5)
    vv.libs.rnd.vox_rotate_vector(base.ctypes.data, v.ctypes.data, res.ctypes.data)
    assert abs(float(res[0, :3] @ res[0, :3]) - 30.0) < 1e-5


def test_context_contracts(vv):
    with pytest.raises(vv.VoxError):
        vv.Context(1000, 600)          # width % 16 (renderer.c:97-100)
    with pytest.raises(vv.VoxError):
        vv.Context(1024, 601)          # height % 4
    ctx = vv.Context(64, 16)
    assert ctx.set_quality(vv.QUALITY_BEST) and ctx.set_quality(vv.QUALITY_ADAPTIVE | vv.QUALITY_RAY_MERGE)
    assert not ctx.set_quality(vv.QUALITY_FAST | vv.QUALITY_RAY_MERGE)   # merge needs ADAPTIVE
    assert not ctx.set_quality(3) and not ctx.set_quality(12 | 2)
    f = ctx.pixfmt()
    s = ctx.h.contents.surface.contents
    assert (s.w, s.h, s.pitch) == (64, 16, 256)
    assert vv.libs.rnd.vox_map_rgb(s.format, 0x12, 0x34, 0x56) == 0x123456
    # a NULL scene renders an all-zero frame without touching the GPU (search.c:22)
    cam = vv.Camera()
    ctx.set_camera(cam)
    ctx.pixels()[:] = 7
    ctx.render(); ctx.redraw()
    assert not ctx.pixels().any()
    tex = ctx.texture()
    assert tex.shape == (32768,) and tex.min() >= 64
    ctx.set_texture(np.arange(32768, dtype=np.uint32).astype(np.uint8))
    assert ctx.texture()[300] == 300 % 256


def test_metrics(vv):
    a, b = vv.api.dots4([[1, 2, 3]]), vv.api.dots4([[4, 0, 3.5]])
    t = vv.libs.trees
    assert t.vox_abs_metric(a.ctypes.data, b.ctypes.data) == 5.5
    assert t.vox_sqr_metric(a.ctypes.data, b.ctypes.data) == 13.25
    assert t.vox_sqr_norm(a.ctypes.data) == 14.0


def test_raw_volume_loader(vv, port, ds, tmp_path, skull):
    path = str(tmp_path / "skull_like.dat")
    ds.write_raw_volume("skull", path)
    t = vv.Tree.from_raw_file(path, (256, 256, 256), 1, 40)
    assert t.voxels() == len(skull)
    assert flat_equal(t.export(), port.tree(skull).flatten()) == []
    with pytest.raises(vv.VoxError, match="Wrong size"):
        vv.Tree.from_raw_file(path, (256, 256, 255), 1, 40)
    with pytest.raises(vv.VoxError):
        vv.Tree.from_raw_file(str(tmp_path / "missing.dat"), (2, 2, 2), 1, 0)


def test_raw_volume_range_loader(vv, tmp_path, monkeypatch):
    """vox_read_raw_data_range: min <= sample < max (modules/voxtrees.c:306-342), 2-byte little-endian samples,
    x outer / z inner (datareader.c:51-76).  The host loop, whatever the machine has."""
    monkeypatch.setenv("VOXTREES_LOADER", "host")
    rng = np.random.default_rng(5)
    dim = (20, 17, 23)
    samples = rng.integers(0, 1 << 16, dim[0] * dim[1] * dim[2], dtype=np.uint32)
    path = str(tmp_path / "v16.dat")
    samples.astype("<u2").tofile(path)
    idx = np.flatnonzero((samples >= 20000) & (samples < 30000))
    want = np.stack([idx // (dim[1] * dim[2]), (idx // dim[2]) % dim[1], idx % dim[2]], 1).astype(np.float32)
    t = vv.Tree.from_raw_file(path, dim, 2, sample_range=(20000, 30000))
    assert t.voxels() == len(idx)
    assert flat_equal(t.export(), vv.Tree.build(want).export()) == []
    monkeypatch.setenv("VOXTREES_LOADER", "nonsense")
    with pytest.raises(vv.VoxError, match="host or gpu"):
        vv.Tree.from_raw_file(path, dim, 2, sample_range=(0, 5))


def test_set_devices_contract_without_a_device(vv):
    """vox_context_set_devices answers how many devices will draw: 0 when there is none; the context stays usable."""
    if vv.gpu_available():
        pytest.skip("a CUDA device is present")
    ctx = vv.Context(64, 16)
    for n in (0, 1, 8, -3):
        assert ctx.set_devices(n) == 0
    ctx.set_scene(vv.Tree.build(CROSS)); ctx.set_camera(vv.Camera())
    ctx.render()
    assert "no CUDA device" in vv.last_error()


def test_no_gpu_means_loud_failure(vv):
    """The product has no CPU search path: without a device every query raises."""
    if vv.gpu_available():
        pytest.skip("a CUDA device is present")
    t = vv.Tree.build(CROSS)
    with pytest.raises(vv.VoxError, match="no CUDA device"):
        t.rays([[0, 0, 9]], [[0, 0, -1]])
    with pytest.raises(vv.VoxError, match="no CUDA device"):
        t.balls([[0, 0, 0]], [1.0])
    with pytest.raises(vv.VoxError, match="no CUDA device"):
        t.ray([0, 0, 9], [0, 0, -1])
    with pytest.raises(vv.VoxError, match="no CUDA device"):
        t.ball([0, 0, 0], 1.0)
    ctx = vv.Context(64, 16)
    ctx.set_scene(t); ctx.set_camera(vv.Camera())
    ctx.render()
    assert "no CUDA device" in vv.last_error()


def test_product_never_touches_the_oracle(vv):
    """No source file of the product mentions the oracle directory or its API."""
    root = os.path.dirname(vv.LIB_DIR)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".c", ".h", ".cu", ".cuh", ".py")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "orc_" not in text and "libvoxoracle" not in text and "libvoxref" not in text, f
                if f.endswith(".py"):
                    assert "import orc" not in text and "from oracle" not in text, f


@pytest.mark.parametrize("kind,name", [(0, "simple-camera"), (1, "doom-camera")])
def test_screen_bounds_contain_every_hit_pixel(vv, port, skull, ds, kind, name):
    """voxgpu_screen_bounds (the render kernel clears tiles outside it without casting their rays):
    over random poses, no pixel whose ray the ORACLE lets into the root box -- let alone one that hits
    a voxel -- may lie outside the bounds.  Host arithmetic only."""
    import ctypes as C
    g = vv.libs.gpu
    g.voxgpu_screen_bounds.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(vv.api.GpuCamera), C.c_int, C.c_int, C.c_void_p]
    # the root box alone as a tree: every ray that enters the box hits it (a dense leaf)
    bb_min, bb_max = skull.min(0), skull.max(0) + 1
    box = port.tree_from_flat({"bb_min": bb_min[None].astype(np.float32), "bb_max": bb_max[None].astype(np.float32),
                               "flags": np.array([2], np.uint32), "dots_num": np.array([1], np.uint32),
                               "center": np.zeros((1, 3), np.float32), "first_child": np.zeros(1, np.uint32),
                               "child_mask": np.zeros(1, np.uint32), "dots_off": np.zeros(1, np.uint64),
                               "dots": np.zeros((0, 3), np.float32)})
    w, h = 160, 96
    rng = np.random.default_rng(5 + kind)
    bounded = 0
    for trial in range(int(os.environ.get("VOX_BOUNDS_TRIALS", "60"))):
        dist = [900.0, 400.0, 250.0, 170.0, 90.0, 0.0][trial % 6]
        v = rng.normal(size=3)
        pos = 128 + dist * v / np.linalg.norm(v) + rng.normal(size=3) * 3
        look = 128 + rng.normal(size=3) * (60 if trial % 4 else 500)
        fov = [0.2, 1.2, 3.0][trial % 3]
        cam, ocam = vv.Camera(name), port.camera(kind)
        rot = [float(x) for x in rng.normal(size=3) * (0.7 if trial % 2 else 0.0)]
        for c_ in (cam, ocam):
            ds.setup_camera(c_, pos, look, fov, w, h)
            c_.rotate(rot)
        gc = cam.gpu_camera()
        out = (C.c_int * 4)()
        mn, mx = np.ascontiguousarray(bb_min, np.float32), np.ascontiguousarray(bb_max, np.float32)
        g.voxgpu_screen_bounds(mn.ctypes.data, mx.ctypes.data, C.byref(gc), w, h, out)
        x0, x1, y0, y1 = list(out)
        o, d = ocam.rays(w, h)
        leaf, _, _ = box.rays(o, d)
        enters = (leaf >= 0).reshape(h, w)
        inside = np.zeros((h, w), bool)
        inside[y0:y1, x0:x1] = True
        assert not (enters & ~inside).any(), (trial, list(out))
        bounded += int((x1 - x0) * (y1 - y0) < w * h)
    assert bounded > 15          # the bounds do bite in many poses (VOX_BOUNDS_TRIALS=1500 was run once: no escape)


def host_cut_boxes(a, limit=512):
    """The coarse host cut of tree.inl (deepest level whose cut has <= limit nodes, plus the leaves above it) of an
    exported tree, as the six arrays voxgpu_visible_spans takes."""
    n = len(a["flags"])
    leafp = (a["flags"] & 3) != 0
    level = np.zeros(n, np.int32)
    for i in np.flatnonzero(~leafp):
        nc, fc = bin(int(a["child_mask"][i])).count("1"), int(a["first_child"][i])
        level[fc:fc + nc] = level[i] + 1
    at, leaves = np.bincount(level), np.bincount(level[leafp], minlength=level.max() + 1)
    best = above = 0
    for lv in range(len(at)):
        if at[lv] + above <= limit:
            best = lv
        above += leaves[lv]
    cut = np.flatnonzero((level == best) | ((level < best) & leafp))
    return np.ascontiguousarray(np.concatenate([a["bb_min"][cut].T, a["bb_max"][cut].T]), np.float32), len(cut)


def test_visible_spans_contain_every_hit_pixel(vv, port, skull, ds):
    """voxgpu_visible_spans (the presenter copies only these columns of every 4-row unit and zeroes the rest on the
    host): over random poses of both cameras, every pixel the ORACLE finds a voxel for lies inside its row's span;
    a pose with a box beside or behind the camera gets no claim; and the spans do bite (they hold clearly fewer
    pixels than the root box's rectangle).  Host arithmetic only."""
    g = vv.libs.gpu
    g.voxgpu_visible_spans.argtypes = [C.c_void_p, C.c_uint, C.POINTER(vv.api.GpuCamera), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    g.voxgpu_visible_spans.restype = C.c_int
    g.voxgpu_screen_bounds.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(vv.api.GpuCamera), C.c_int, C.c_int, C.c_void_p]
    boxes, n = host_cut_boxes(vv.Tree.build(skull).export())
    assert 64 < n <= 512
    otree = port.tree(skull)
    w, h = 320, 160
    rng = np.random.default_rng(77)
    claims = kept = rect = 0
    for trial in range(int(os.environ.get("VOX_BOUNDS_TRIALS", "60"))):
        kind = trial % 2
        dist = [700.0, 330.0, 220.0, 150.0, 60.0][trial % 5]
        v = rng.normal(size=3)
        pos = 128 + dist * v / np.linalg.norm(v) + rng.normal(size=3) * 3
        look = 128 + rng.normal(size=3) * (40 if trial % 4 else 400)
        cam, ocam = vv.Camera("doom-camera" if kind else "simple-camera"), port.camera(kind)
        rot = [float(x) for x in rng.normal(size=3) * (0.5 if trial % 3 == 0 else 0.0)]
        for c_ in (cam, ocam):
            ds.setup_camera(c_, pos, look, [0.4, 1.2, 2.5][trial % 3], w, h)
            c_.rotate(rot)
        gc = cam.gpu_camera()
        lo, hi = np.zeros(h // 4, np.int32), np.zeros(h // 4, np.int32)
        if not g.voxgpu_visible_spans(boxes.ctypes.data, n, C.byref(gc), w, h, lo.ctypes.data, hi.ctypes.data):
            continue
        claims += 1
        assert (lo >= 0).all() and (hi <= w).all()
        hit = (otree.rays(*ocam.rays(w, h))[0] >= 0).reshape(h, w)
        x = np.arange(w)[None, :]
        inside = (x >= np.repeat(lo, 4)[:, None]) & (x < np.repeat(hi, 4)[:, None])
        assert not (hit & ~inside).any(), trial
        out = (C.c_int * 4)()
        mn, mx = np.ascontiguousarray(boxes[:3].min(1)), np.ascontiguousarray(boxes[3:].max(1))
        g.voxgpu_screen_bounds(mn.ctypes.data, mx.ctypes.data, C.byref(gc), w, h, out)
        kept += int(inside.sum())
        rect += (out[1] - out[0]) * (out[3] - out[2])
    assert claims > 15 and kept < 0.9 * rect, (claims, kept, rect)
    # a camera in the middle of the volume has boxes behind it: no claim
    cam = ds.setup_camera(vv.Camera("simple-camera"), [128, 128, 128], [128, 300, 128], 1.2, w, h)
    gc = cam.gpu_camera()
    assert g.voxgpu_visible_spans(boxes.ctypes.data, n, C.byref(gc), w, h, lo.ctypes.data, hi.ctypes.data) == 0


def test_present_plan_keeps_every_hit_pixel_and_half_the_bytes(vv, port, skull, ds):
    """voxgpu_present_plan = what the presenter copies (bands, visible spans, strips chosen by cost), host arithmetic
    only.  Over random poses of both cameras, 1 and 4 bands, the default knobs and a strip per group of rows: the
    kept columns hold every pixel the ORACLE finds a voxel for, lie on 64-byte lines, add up to the bytes returned,
    and never exceed what the root box's rectangle alone keeps.  On the C2 orbit at 1080p the default plan copies
    less than half of the frames' bytes (46 %; the rectangle: 55.5 %)."""
    g = vv.libs.gpu
    g.voxgpu_present_plan.argtypes = [C.c_void_p, C.c_uint, C.POINTER(vv.api.GpuCamera), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    g.voxgpu_present_plan.restype = C.c_longlong
    boxes, n = host_cut_boxes(vv.Tree.build(skull).export())
    otree = port.tree(skull)
    w, h = 320, 160
    lo, hi = np.zeros(h // 4, np.int32), np.zeros(h // 4, np.int32)

    def plan(gc, ww, hh, nbands, lo_, hi_):
        return int(g.voxgpu_present_plan(boxes.ctypes.data, n, C.byref(gc), ww, hh, nbands, lo_.ctypes.data, hi_.ctypes.data))

    rng = np.random.default_rng(99)
    tighter = 0
    try:
        for trial in range(40):
            kind = trial % 2
            dist = [700.0, 330.0, 220.0, 150.0, 60.0][trial % 5]
            v = rng.normal(size=3)
            pos = 128 + dist * v / np.linalg.norm(v) + rng.normal(size=3) * 3
            look = 128 + rng.normal(size=3) * (40 if trial % 4 else 400)
            cam, ocam = vv.Camera("doom-camera" if kind else "simple-camera"), port.camera(kind)
            rot = [float(x) for x in rng.normal(size=3) * (0.5 if trial % 3 == 0 else 0.0)]
            for c_ in (cam, ocam):
                ds.setup_camera(c_, pos, look, [0.4, 1.2, 2.5][trial % 3], w, h)
                c_.rotate(rot)
            gc = cam.gpu_camera()
            hit = (otree.rays(*ocam.rays(w, h))[0] >= 0).reshape(h, w)
            x = np.arange(w)[None, :]
            sizes = []
            for knobs in ((0, 16, 98304), (1, 16, 98304), (1, 32, 0)):
                g.voxgpu_set_present_spans(*knobs)
                for nbands in (1, 4):
                    got = plan(gc, w, h, nbands, lo, hi)
                    assert got >= -1
                    if got == -1:
                        assert not hit.any() and not hi.any()
                        continue
                    assert not (lo % 16).any() and not (hi % 16).any() and (lo >= 0).all() and (hi <= w).all()
                    assert got == int(np.maximum(hi - lo, 0).sum()) * 16
                    inside = (x >= np.repeat(lo, 4)[:, None]) & (x < np.repeat(hi, 4)[:, None])
                    assert not (hit & ~inside).any(), (trial, knobs, nbands)
                    if nbands == 4:
                        sizes.append(got)
            if len(sizes) == 3:
                assert w * h * 4 >= sizes[0] >= sizes[1] >= sizes[2]
                tighter += sizes[2] < sizes[0]
        assert tighter > 8
        # the C2 orbit at full size, as bench.py's one-context leg draws it (4 bands)
        g.voxgpu_set_present_spans(1, 16, 98304)
        W, H = 1920, 1080
        lo2, hi2 = np.zeros(H // 4, np.int32), np.zeros(H // 4, np.int32)
        cam = vv.Camera("simple-camera")
        total = rect = 0
        for f in range(0, 360, 6):
            ds.setup_camera(cam, *ds.orbit_pose(f, 360), 1.2, W, H)
            gc = cam.gpu_camera()
            total += plan(gc, W, H, 4, lo2, hi2)
            g.voxgpu_set_present_spans(0, 16, 98304)
            rect += plan(gc, W, H, 4, lo2, hi2)
            g.voxgpu_set_present_spans(1, 16, 98304)
        frames = len(range(0, 360, 6))
        assert 0.40 < total / (frames * W * H * 4) < 0.50 and 0.53 < rect / (frames * W * H * 4) < 0.58, (total, rect)
    finally:
        g.voxgpu_set_present_spans(1, 16, 98304)


def test_screen_mask_argument_on_the_host(vv, port, skull, ds):
    """The argument behind the render kernel's screen mask (k_mark in voxgpu.cu), replayed in numpy: take
    the same cut through the tree (deepest level with <= 131072 nodes plus the leaves above it), project
    the eight corners of every cut box with the inverse of screen2world in fp32, widen the bounding
    rectangle by 2 pixels, mark 8x4 tiles.  Every pixel the ORACLE finds a voxel for must lie in a marked
    tile -- over random poses of both kinds of camera."""
    t = vv.Tree.build(skull)
    a = t.export()
    n = len(a["flags"])
    leafp = (a["flags"] & 3) != 0
    level = np.zeros(n, np.int32)
    for i in np.flatnonzero(~leafp):
        nc, fc = bin(int(a["child_mask"][i])).count("1"), int(a["first_child"][i])
        level[fc:fc + nc] = level[i] + 1
    at, leaves = np.bincount(level), np.bincount(level[leafp], minlength=level.max() + 1)
    best = above = 0
    for lv in range(len(at)):
        if at[lv] + above <= 131072:
            best = lv
        above += leaves[lv]
    cut = np.flatnonzero((level == best) | ((level < best) & leafp))
    mn, mx = a["bb_min"][cut], a["bb_max"][cut]
    otree = port.tree(skull)
    w, h = 320, 160
    rng = np.random.default_rng(21)
    f32 = np.float32
    useful = 0
    for trial in range(40):
        kind = trial % 2
        dist = [700.0, 330.0, 220.0, 150.0][trial % 4]
        v = rng.normal(size=3)
        pos = 128 + dist * v / np.linalg.norm(v)
        look = 128 + rng.normal(size=3) * 30
        cam, ocam = vv.Camera("doom-camera" if kind else "simple-camera"), port.camera(kind)
        rot = [float(x) for x in rng.normal(size=3) * (0.4 if trial % 3 == 0 else 0.0)]
        for c_ in (cam, ocam):
            ds.setup_camera(c_, pos, look, [0.6, 1.2][trial % 2], w, h)
            c_.rotate(rot)
        gc = cam.gpu_camera()
        p0 = np.array(gc.position, f32)
        lo = np.full((len(cut), 2), np.inf, f32)
        hi = -lo
        front = np.ones(len(cut), bool)
        for k in range(8):
            c = np.where(np.array([k & 1, k & 2, k & 4], bool)[None, :], mx, mn).astype(f32) - p0
            if kind == 0:
                qw, u = f32(gc.rotation[0]), -np.array(gc.rotation[1:], f32)
                tt = f32(2) * np.cross(u, c).astype(f32)
                vv_ = (c + qw * tt + np.cross(u, tt)).astype(f32)
            else:
                cs, sn = f32(gc.cosphi), f32(gc.sinphi)
                vv_ = np.stack([c[:, 0] * cs + c[:, 1] * sn, c[:, 1] * cs - c[:, 0] * sn, c[:, 2]], 1).astype(f32)
            front &= vv_[:, 1] > f32(1e-3) * np.abs(c).sum(1)
            with np.errstate(divide="ignore", invalid="ignore"):
                rz = vv_[:, 2] / vv_[:, 1] - (f32(gc.k) if kind else f32(0))
                s = np.stack([(vv_[:, 0] / vv_[:, 1] + f32(gc.xsub)) / f32(gc.mul), (f32(gc.ysub) - rz) / f32(gc.mul)], 1)
            lo, hi = np.minimum(lo, s), np.maximum(hi, s)
        leaf = otree.rays(*ocam.rays(w, h))[0].reshape(h, w)
        if not front.all():
            continue                    # the kernel raises its "no bounds" flag for such a frame
        mask = np.zeros((h // 4, w // 8), bool)
        x0 = np.clip(np.floor(lo[:, 0]) - 2, 0, w).astype(int) >> 3
        x1 = (np.clip(np.ceil(hi[:, 0]) + 3, 0, w).astype(int) + 7) >> 3
        y0 = np.clip(np.floor(lo[:, 1]) - 2, 0, h).astype(int) >> 2
        y1 = (np.clip(np.ceil(hi[:, 1]) + 3, 0, h).astype(int) + 3) >> 2
        for i in range(len(cut)):
            mask[y0[i]:y1[i], x0[i]:x1[i]] = True
        hit_tiles = (leaf >= 0).reshape(h // 4, 4, w // 8, 8).any((1, 3))
        assert not (hit_tiles & ~mask).any(), trial
        useful += int(mask.mean() < 0.9)
    assert useful > 10


def build_c_example(tmp_path):
    import subprocess
    exe = str(tmp_path / "orbit")
    lib = os.path.join(ROOT, "voxvision_b200", "lib")
    cmd = ["gcc", "-std=gnu11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "orbit.c"), "-L" + lib, "-lvoxrnd", "-lvoxtrees", "-lm",
           "-Wl,-rpath," + lib, "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def test_c_example_compiles_against_the_headers(vv, tmp_path):
    """examples/orbit.c: the reference's scene setup in plain C against include/, warnings as errors.
    Without a CUDA device the program must fail loudly (exit code 2, the library's message), never render."""
    import subprocess
    exe = build_c_example(tmp_path)
    if not vv.gpu_available():
        r = subprocess.run([exe, "1"], capture_output=True, text=True)
        assert r.returncode == 2 and "no CUDA device" in r.stderr


def test_screenshot_is_a_readable_bmp(vv, tmp_path):
    """vox_screenshot (screenshot.c:7-36): numbered 24-bit BMPs of the surface, host only."""
    R = vv.libs.rnd
    R.vox_screenshot.argtypes = [C.c_void_p, C.c_char_p]
    R.vox_screenshot.restype = C.c_int
    ctx = vv.Context(32, 8)
    pix = ctx.pixels()
    rng = np.random.default_rng(3)
    pix[:] = rng.integers(0, 1 << 24, pix.shape, dtype=np.uint32)           # XRGB8888
    want = pix.copy()
    assert R.vox_screenshot(ctx.h, str(tmp_path).encode()) == 1
    assert R.vox_screenshot(ctx.h, str(tmp_path).encode()) == 1
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 2 and all(re.fullmatch(r"screenshot\d+\.bmp", f) for f in files)
    raw = open(tmp_path / files[0], "rb").read()
    assert raw[:2] == b"BM" and int.from_bytes(raw[18:22], "little") == 32 and int.from_bytes(raw[22:26], "little") == 8
    assert int.from_bytes(raw[28:30], "little") == 24 and len(raw) == 54 + 32 * 3 * 8
    bgr = np.frombuffer(raw[54:], np.uint8).reshape(8, 32, 3)[::-1]
    assert np.array_equal(bgr[..., 2], (want >> 16) & 0xff) and np.array_equal(bgr[..., 1], (want >> 8) & 0xff)
    assert np.array_equal(bgr[..., 0], want & 0xff)
    assert R.vox_screenshot(ctx.h, b"/nonexistent-dir/xyz") == 0 and "vox_screenshot" in vv.last_error()
