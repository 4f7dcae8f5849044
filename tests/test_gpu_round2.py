"""Parity tests for what round 2 added to the CUDA path: ray batches searched in entry order, the screen
mask's ticket phases with split frames (C5), frames split over the devices of one process behind
vox_render, rendering from several host threads, subtree scenes, the raw-volume loader on the device.
Bit-exact against the oracle / the golden fixtures, through the C-ABI, like tests/test_gpu_parity.py."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

from test_gpu_parity import assert_rays_equal, render_device
from util import bits, flat_equal, sha

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


# ------------------------------------------------------------ K2 in entry order

@pytest.mark.parametrize("scene", ["skull", "sparse"])
def test_sorted_ray_batches_equal_unsorted_and_oracle(gpu, checker, ds, skull, scene):
    """voxgpu_set_rays_sort: the order a batch is searched in changes nothing, ray by ray -- checked for sizes
    around the sort's internal units (32-key steps, 4096-key warp runs), with rays that miss the root box,
    rays starting inside it and duplicates."""
    g = gpu.api.libs.gpu
    if scene == "skull":
        vox, side = skull, 256
    else:
        rng = np.random.default_rng(7)
        vox, side = np.unique(rng.integers(0, 96, (60000, 3)), axis=0).astype(np.float32), 96
    t, ot = gpu.Tree.build(vox), checker.tree(vox)
    o, d = ds.random_rays(70001, side=side)
    o[100:200] = o[0]; d[100:200] = d[0]                         # duplicates
    o[400:500] = np.float32(side / 2)                            # from inside the box
    d[600:700] = -d[600:700]                                     # mostly away from the volume
    want = ot.rays(o, d, threads=THREADS)
    try:
        for n in (1, 31, 32, 33, 4095, 4096, 4097, 70001):
            g.voxgpu_set_rays_sort(0)
            plain = t.rays(o[:n], d[:n])
            g.voxgpu_set_rays_sort(1)
            srt = t.rays(o[:n], d[:n])
            for a, b in zip(plain, srt):
                assert np.array_equal(bits(a) if a.dtype == np.float32 else a, bits(b) if b.dtype == np.float32 else b), n
            assert_rays_equal(srt, tuple(x[:n] for x in want))
        # from an inner node and from a leaf: the keys come from that node's box
        flat = t.export()
        inner = int(np.flatnonzero((flat["flags"] & 3) == 0)[3])
        leaf = int(np.flatnonzero((flat["flags"] & 3) == 1)[5])
        for nid in (inner, leaf):
            got = t.rays(o[:5000], d[:5000], start_node=nid)
            assert_rays_equal(got, ot.rays(o[:5000], d[:5000], start_node=nid, threads=THREADS))
    finally:
        g.voxgpu_set_rays_sort(-1)


# ----------------------------------------------------------------- C5, split frames

def c5_small(ds):
    """A C5-shaped scene at a size the oracle walks in seconds: 256^3 grid, union of 96 balls with hash holes."""
    return ds.ball_union_voxels(256, 96, rmin=6.0, rmax=22.0)


def test_c5_split_frame_equals_one_launch_and_oracle(gpu, checker, ds):
    """Config C5 in small: a frame of a ball-union volume drawn (a) by one launch, (b) by the tile rows of 2, 3
    and 8 virtual ranks into ONE buffer, each rank with its own screen mask of only its rows -- the N-GPU
    partition of bench.py's C5 leg -- and (c) searched ray by ray by the oracle: 1 327 104 rays, all compared."""
    vox = c5_small(ds)
    assert 1_000_000 < len(vox) < 4_000_000
    t, ot = gpu.Tree.build(vox), checker.tree(vox)
    assert flat_equal(t.export(), ot.flatten()) == []
    w, h = 1536, 864
    ctx = gpu.Context(w, h)
    tex, fmt = ctx.texture(), ctx.pixfmt()
    pos, look = ds.orbit_pose(37, 360, (128.0, 128.0, 128.0), 328.0)
    cam = ds.setup_camera(gpu.Camera(), pos, look, 1.2, w, h)
    ocam = ds.setup_camera(checker.camera(0), pos, look, 1.2, w, h)
    pix, leaf, pts = render_device(gpu, t, cam, w, h, tex, fmt)
    o, d = ocam.rays(w, h)
    oleaf, ohit, _ = ot.rays(o, d, threads=THREADS)
    assert w * h >= 1_000_000 and int((oleaf >= 0).sum()) > 150_000
    assert np.array_equal(leaf.reshape(-1), oleaf)
    m = oleaf >= 0
    assert np.array_equal(bits(pts.reshape(-1, 3)[m]), bits(ohit[m]))
    opix = ot.render(ocam, w, h, 0, tex, threads=THREADS, want_hits=False)
    assert np.array_equal(pix, opix)

    api, g = gpu.api, gpu.api.libs.gpu
    gc = cam.gpu_camera()
    d_tex = gpu.DeviceBuffer.from_array(tex)
    for cull in (1, 2, 0):                       # mask + bounds, bounds only, nothing
        g.voxgpu_set_cull(cull)
        try:
            for world in (2, 3, 8):
                d_pix, d_leaf = gpu.DeviceBuffer(w * h * 4), gpu.DeviceBuffer(w * h * 4)
                g.voxgpu_memset(d_pix.ptr, 0x5a, w * h * 4, None)
                for rank in range(world):
                    fr = api.GpuFrame(w, h, 0, h, w, 0, fmt, None, world, rank)
                    assert g.voxgpu_render(t.device_handle(), C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, d_leaf.ptr, None, None) == 0
                assert np.array_equal(d_pix.download(np.uint32, (h, w)), opix), (cull, world)
                assert np.array_equal(d_leaf.download(np.int32, (h * w,)), oleaf), (cull, world)
        finally:
            g.voxgpu_set_cull(1)


def test_masked_ticket_phases_leave_no_pixel_behind(gpu, skull, ds):
    """Every pixel of every band is written exactly as without culling, whatever the frame showed before:
    poisoned buffers, odd band cuts, a frame whose mask marks nothing (camera looking away) and one with a
    box behind the camera (no bounds: everything is traversed)."""
    t = gpu.Tree.build(skull)
    w, h = 1280, 720
    ctx = gpu.Context(w, h)
    tex, fmt = ctx.texture(), ctx.pixfmt()
    g = gpu.api.libs.gpu
    poses = [ds.orbit_pose(123), ((128, 128, -200), (128, 128, -900)), ((128, 128, 128), (300, 128, 128)),
             ((40, 200, 90), (128, 100, 128))]
    for pos, look in poses:
        cam = ds.setup_camera(gpu.Camera(), pos, look, 1.2, w, h)
        g.voxgpu_set_cull(0)
        want = render_device(gpu, t, cam, w, h, tex, fmt)
        g.voxgpu_set_cull(1)
        got = render_device(gpu, t, cam, w, h, tex, fmt)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        for y0, y1 in ((0, 4), (4, 360), (360, 716), (716, 720)):
            band = render_device(gpu, t, cam, w, h, tex, fmt, y0, y1)
            assert np.array_equal(band[0], want[0][y0:y1]) and np.array_equal(band[1], want[1][y0:y1])


@pytest.mark.parametrize("camera", ["simple-camera", "doom-camera"])
def test_visible_rectangle_frames_equal_the_oracle(gpu, checker, skull, ds, camera):
    """vox_render draws and copies only the part of the frame inside the root box's screen bounds -- and of that,
    per strip of rows, only the columns the tree's coarse cut can show (the visible spans) -- and zeroes the rest
    on the host (present.inl).  Random poses -- far, near, inside the volume, rolled, looking away, the box
    partly off screen, narrow and wide fov -- through vox_render / vox_redraw into a POISONED surface, in BEST and in
    the reference's default quality, on one context reused for all poses (so every frame meets the leftovers of a
    frame with a different rectangle): every pixel must equal the oracle's frame, with the spans off, with the
    default knobs and with every band cut into as many strips as it has (copy cost 0)."""
    t, otree = gpu.Tree.build(skull), checker.tree(skull)
    w, h = 320, 160
    ctx = gpu.Context(w, h)
    ctx.set_scene(t)
    rng = np.random.default_rng(23)
    kind = 0 if camera == "simple-camera" else 1
    g = gpu.api.libs.gpu
    shrunk = tighter = 0
    for trial in range(30):
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
        q = gpu.QUALITY_BEST if trial % 3 else gpu.QUALITY_ADAPTIVE
        assert ctx.set_quality(q)
        ctx.set_camera(cam)
        opix = otree.render(ocam, w, h, q, ctx.texture(), threads=THREADS, want_hits=False)
        sizes = []
        try:
            for knobs in ((0, 16, 98304), (1, 16, 98304), (1, 32, 0)):
                g.voxgpu_set_present_spans(*knobs)
                ctx.pixels()[:] = 0xdeadbeef
                before = int(g.voxgpu_present_bytes_copied())
                ctx.render()
                ctx.redraw()
                assert gpu.last_error() == ""
                sizes.append(int(g.voxgpu_present_bytes_copied()) - before)
                assert np.array_equal(ctx.pixels(), opix), (trial, knobs, sizes)
        finally:
            g.voxgpu_set_present_spans(1, 16, 98304)
        assert w * h * 4 >= sizes[0] >= sizes[1] >= sizes[2]
        shrunk += sizes[0] < w * h * 4
        tighter += sizes[2] < sizes[0]
    assert shrunk > 8           # the rectangle did shrink the copy in a good part of the poses
    assert tighter > 4          # and the spans shrank it further


def test_tile_cost_feedback_changes_nothing(gpu, golden, skull, ds):
    """voxgpu_set_tile_feedback: the order in which a frame's tiles are handed out follows the cost the previous
    frame recorded; pixels, leaf ids and hit points do not -- for a sequence of frames, a change of geometry in
    between (the map is re-made), every tile heavy (threshold 1), the default, and feedback off."""
    meta, _ = golden
    t = gpu.Tree.build(skull)
    ctx = gpu.Context(1920, 1080)
    tex, fmt = ctx.texture(), ctx.pixfmt()
    g = gpu.api.libs.gpu
    frames = sorted(int(k) for k in meta["c2"])
    try:
        for thr in (1, 48, 0):
            g.voxgpu_set_tile_feedback(thr)
            for rep in range(2):
                for f in frames[:4] if rep else frames:
                    cam = ds.setup_camera(gpu.Camera(), *ds.orbit_pose(f), 1.2, 1920, 1080)
                    pix, leaf, pts = render_device(gpu, t, cam, 1920, 1080, tex, fmt)
                    want = meta["c2"][str(f)]
                    assert sha(pix) == want["pix_sha"] and sha(leaf) == want["leaf_sha"] and sha(pts[leaf >= 0]) == want["hit_sha"], (thr, f)
                cam = ds.setup_camera(gpu.Camera(), *ds.orbit_pose(77), 1.2, 640, 360)
                small = render_device(gpu, t, cam, 640, 360, tex, fmt, want_hits=False)[0]
                g.voxgpu_set_cull(0)
                plain = render_device(gpu, t, cam, 640, 360, tex, fmt, want_hits=False)[0]
                g.voxgpu_set_cull(1)
                assert np.array_equal(small, plain)
    finally:
        g.voxgpu_set_tile_feedback(48)
        g.voxgpu_set_cull(1)


def test_render_errors_leave_the_device_usable(gpu, skull, ds):
    """Arguments are validated before anything is queued; a rejected call must not disturb later ones."""
    t = gpu.Tree.build(skull)
    w, h = 640, 360
    ctx = gpu.Context(w, h)
    cam = ds.setup_camera(gpu.Camera(), *ds.orbit_pose(10), 1.2, w, h)
    want = render_device(gpu, t, cam, w, h, ctx.texture(), ctx.pixfmt(), want_hits=False)[0]
    api, g = gpu.api, gpu.api.libs.gpu
    gc = cam.gpu_camera()
    d_tex, d_pix = gpu.DeviceBuffer.from_array(ctx.texture()), gpu.DeviceBuffer(w * h * 4)
    for _ in range(70):                          # more rejected calls than there are render slots
        for bad in (api.GpuFrame(w, h, 0, h, w, 3, ctx.pixfmt()), api.GpuFrame(w, h, 0, h, w, 1 | 4, ctx.pixfmt()),
                    api.GpuFrame(w, h, 0, h, w, 0, ctx.pixfmt(), None, 4, 4),
                    api.GpuFrame(w, h, 0, h, w, 0, ctx.pixfmt(), None, 0, 0, 1 << 30)):
            assert g.voxgpu_render(t.device_handle(), C.byref(gc), C.byref(bad), d_tex.ptr, d_pix.ptr, None, None, None) != 0
    for q in (0, 2):
        got = render_device(gpu, t, cam, w, h, ctx.texture(), ctx.pixfmt(), want_hits=False, quality=q)[0]
        if q == 0:
            assert np.array_equal(got, want)


# ---------------------------------------------------------- vox_render from threads

def test_vox_render_from_four_threads(gpu, golden, skull, ds):
    """vox_render may be called from any thread (Manual.md:171-175): four host threads, each with its own
    context and camera on the shared tree, render the golden C2 frames at the same time, in all quality
    paths that share device state (BEST: slots + masks; ADAPTIVE: the slots' scratch)."""
    meta, _ = golden
    t = gpu.Tree.build(skull)
    t.device_handle()
    frames = sorted(int(k) for k in meta["c2"])
    w, h = 1920, 1080
    errors, results = [], {}

    def worker(idx):
        try:
            gpu.set_device(0)                    # the current device is per thread
            ctx = gpu.Context(w, h)
            ctx.set_scene(t)
            cam = gpu.Camera()
            for rep in range(3):
                for f in frames[idx::4] + frames[(idx + 1) % 4::4]:
                    ctx.set_quality(gpu.QUALITY_BEST if (rep + idx) % 2 == 0 else gpu.QUALITY_ADAPTIVE)
                    ds.setup_camera(cam, *ds.orbit_pose(f), 1.2, w, h)
                    ctx.set_camera(cam)
                    ctx.render()
                    ctx.redraw()
                    if (rep + idx) % 2 == 0:
                        results.setdefault(f, set()).add(sha(ctx.pixels()))
                    else:
                        results.setdefault(("adaptive", f), set()).add(sha(ctx.pixels()))
            err = gpu.last_error()
            if err:
                errors.append(err)
        except Exception as e:                   # noqa: BLE001 - reported below
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for f in frames:
        assert results[f] == {meta["c2"][str(f)]["pix_sha"]}, f
    for key, digests in results.items():
        assert len(digests) == 1, key            # ADAPTIVE: the same frame from whichever thread


# ------------------------------------------------------------------ subtree scenes

def test_scene_may_be_any_node(gpu, checker, skull, ds):
    """vox_context_set_scene accepts any node (the reference searches inside returned leaves, renderer.c:297):
    the frame of an inner node's handle shows that subtree only."""
    t, ot = gpu.Tree.build(skull), checker.tree(skull)
    flat = t.export()
    inner = np.flatnonzero((flat["flags"] & 3) == 0)
    w, h = 640, 360
    ctx = gpu.Context(w, h)
    ctx.set_quality(gpu.QUALITY_BEST)
    cam = ds.setup_camera(gpu.Camera(), *ds.orbit_pose(60), 1.2, w, h)
    ocam = ds.setup_camera(checker.camera(0), *ds.orbit_pose(60), 1.2, w, h)
    o, d = ocam.rays(w, h)
    ctx.set_camera(cam)
    api, g = gpu.api, gpu.api.libs.gpu
    for nid in (int(inner[1]), int(inner[40]), int(inner[len(inner) // 2])):
        oleaf, ohit, _ = ot.rays(o, d, start_node=nid, threads=THREADS)
        # the C-ABI entry, with per-pixel results
        d_tex, d_pix = gpu.DeviceBuffer.from_array(ctx.texture()), gpu.DeviceBuffer(w * h * 4)
        d_leaf, d_pts = gpu.DeviceBuffer(w * h * 4), gpu.DeviceBuffer(w * h * 16)
        fr = api.GpuFrame(w, h, 0, h, w, 0, ctx.pixfmt(), None, 0, 0, nid)
        gc = cam.gpu_camera()
        assert g.voxgpu_render(t.device_handle(), C.byref(gc), C.byref(fr), d_tex.ptr, d_pix.ptr, d_leaf.ptr, d_pts.ptr, None) == 0
        leaf = d_leaf.download(np.int32, (h * w,))
        assert np.array_equal(leaf, oleaf)
        m = oleaf >= 0
        assert np.array_equal(bits(d_pts.download(np.float32, (h * w, 4))[:, :3][m]), bits(ohit[m]))
        pix = d_pix.download(np.uint32, (h, w))
        # and the reference-facing one
        sub = gpu.Tree(t.node(nid), owner=False)
        ctx.set_scene(sub)
        ctx.render()
        ctx.redraw()
        assert gpu.last_error() == ""
        assert np.array_equal(ctx.pixels(), pix)
        assert int((pix != 0).sum()) == int(m.sum())
    whole = int((render_device(gpu, t, cam, w, h, ctx.texture(), ctx.pixfmt(), want_hits=False)[0] != 0).sum())
    assert whole > int(m.sum())


# ------------------------------------------------- one frame on several devices

def test_frames_split_over_the_devices_of_one_process(gpu, golden, checker, skull, bonsai, ds):
    """vox_context_set_devices: vox_render deals the tile rows of every frame to all visible devices, each
    copies its rows into the surface itself; frames equal the golden ones, lit frames equal the oracle's,
    the quality modes that cannot be split still come out right, edits reach every replica."""
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs at least two CUDA devices")
    meta, _ = golden
    t = gpu.Tree.build(skull)
    w, h = 1920, 1080
    ctx = gpu.Context(w, h)
    assert ctx.set_devices(0) == n
    ctx.set_scene(t)
    ctx.set_quality(gpu.QUALITY_BEST)
    cam = gpu.Camera()
    for rep in range(2):
        for f in sorted(int(k) for k in meta["c2"]):
            ds.setup_camera(cam, *ds.orbit_pose(f), 1.2, w, h)
            ctx.set_camera(cam)
            ctx.pixels()[:] = 0xdeadbeef
            ctx.render()
            ctx.redraw()
            assert gpu.last_error() == ""
            assert sha(ctx.pixels()) == meta["c2"][str(f)]["pix_sha"], f
    assert ctx.kernel_ms() > 0
    # the default (ADAPTIVE) cannot be split: one device draws it, same pixels as a one-device context
    one = gpu.Context(w, h)
    one.set_scene(t); one.set_camera(cam)
    ctx.set_quality(gpu.QUALITY_ADAPTIVE); one.set_quality(gpu.QUALITY_ADAPTIVE)
    for c in (ctx, one):
        c.render(); c.redraw()
    assert np.array_equal(ctx.pixels(), one.pixels())
    # back to BEST with two devices only, after an edit of the tree (every replica is refreshed)
    assert ctx.set_devices(2) == 2
    ctx.set_quality(gpu.QUALITY_BEST); one.set_quality(gpu.QUALITY_BEST)
    t.delete(skull[1000:1400])
    for c in (ctx, one):
        c.render(); c.redraw()
    assert gpu.last_error() == ""
    assert np.array_equal(ctx.pixels(), one.pixels())
    # lights are replicated too
    lm = gpu.LightManager()
    ds.apply_lights(lm, ds.light_script())
    ctx.set_light_manager(lm); one.set_light_manager(lm)
    for c in (ctx, one):
        c.render(); c.redraw()
    assert gpu.last_error() == ""
    assert np.array_equal(ctx.pixels(), one.pixels())
    assert int((ctx.pixels() != 0).sum()) > 100000


# ----------------------------------------------------------- loader on the device

@pytest.mark.parametrize("name", ["skull", "bonsai"])
def test_raw_volume_extraction_on_the_device(gpu, ds, tmp_path, monkeypatch, name):
    """f-2: threshold + stable compaction on the GPU give the voxel list of datareader.c:51-76 (x outer, z
    inner) and therefore the same tree as the host loader."""
    path = str(tmp_path / f"{name}.dat")
    ds.write_raw_volume(name, path)
    vox = ds.volume_voxels(name)
    raw = np.fromfile(path, np.uint8)
    g, api = gpu.api.libs.gpu, gpu.api
    dots, count = api._vp(), C.c_size_t()
    dim = (C.c_uint * 3)(256, 256, 256)
    voxel = (C.c_float * 3)(1.0, 1.0, 1.0)
    assert g.voxgpu_volume_extract(raw.ctypes.data, C.byref(dim), 1, 41, 1 << 32, C.byref(voxel), C.byref(dots), C.byref(count)) == 0
    assert count.value == len(vox)
    got = np.ctypeslib.as_array(C.cast(dots.value, C.POINTER(C.c_float)), (count.value, 4)).copy()
    g.voxgpu_host_free(dots)
    assert np.array_equal(got[:, :3], vox) and (got[:, 3] == 0).all()
    monkeypatch.setenv("VOXTREES_LOADER", "gpu")
    on_device = gpu.Tree.from_raw_file(path, (256, 256, 256), 1, 40)
    monkeypatch.setenv("VOXTREES_LOADER", "host")
    on_host = gpu.Tree.from_raw_file(path, (256, 256, 256), 1, 40)
    assert on_device.voxels() == len(vox)
    assert flat_equal(on_device.export(), on_host.export()) == []


def test_raw_volume_range_predicate_and_sample_sizes(gpu, tmp_path, monkeypatch):
    """min <= sample < max (modules/voxtrees.c:306-342) on 1-, 2- and 3-byte little-endian samples, ragged
    dimensions, non-unit voxels; device list == numpy == host loader."""
    rng = np.random.default_rng(11)
    dim = (37, 50, 29)
    nsamp = dim[0] * dim[1] * dim[2]
    g, api = gpu.api.libs.gpu, gpu.api
    gpu.set_voxel_size(0.5, 2.0, 1.25)
    try:
        for size, lo, hi in ((1, 100, 180), (2, 9000, 40000), (3, 1 << 20, 1 << 23), (2, 0, 1), (1, 0, 256), (2, 70000, 70001)):
            samples = rng.integers(0, 1 << (8 * size), nsamp, dtype=np.uint32)
            raw = np.zeros((nsamp, size), np.uint8)
            for b in range(size):
                raw[:, b] = (samples >> (8 * b)) & 0xff
            path = str(tmp_path / f"v{size}_{lo}.dat")
            raw.tofile(path)
            keep = (samples >= lo) & (samples < hi)
            idx = np.flatnonzero(keep)
            want = np.stack([(idx // (dim[1] * dim[2])).astype(np.float32) * np.float32(0.5),
                             ((idx // dim[2]) % dim[1]).astype(np.float32) * np.float32(2.0),
                             (idx % dim[2]).astype(np.float32) * np.float32(1.25)], 1)
            dots, count = api._vp(), C.c_size_t()
            cdim, voxel = (C.c_uint * 3)(*dim), (C.c_float * 3)(0.5, 2.0, 1.25)
            assert g.voxgpu_volume_extract(raw.ctypes.data, C.byref(cdim), size, lo, hi, C.byref(voxel), C.byref(dots), C.byref(count)) == 0
            assert count.value == len(idx)
            if len(idx):
                got = np.ctypeslib.as_array(C.cast(dots.value, C.POINTER(C.c_float)), (count.value, 4)).copy()
                g.voxgpu_host_free(dots)
                assert np.array_equal(bits(got[:, :3]), bits(want))
            trees = {}
            for how in ("gpu", "host"):
                monkeypatch.setenv("VOXTREES_LOADER", how)
                if len(idx):
                    trees[how] = gpu.Tree.from_raw_file(path, dim, size, sample_range=(lo, hi)).export()
                else:
                    assert gpu.api.libs.trees.vox_read_raw_data_range(path.encode(), C.byref(cdim), size, lo, hi, None) is None
            if len(idx):
                assert flat_equal(trees["gpu"], trees["host"]) == []
    finally:
        gpu.set_voxel_size(1, 1, 1)
