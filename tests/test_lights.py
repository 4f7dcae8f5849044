"""Light manager + M-tree, host side: the M-tree behind vox_insert_shadowless_light /
vox_delete_shadowless_light has, node for node and bit for bit, the shape the reference's mtree.c gives
it (compared through the flat snapshot the renderer uploads), and vox_get_light returns the reference's
fp32 sums.  Against oracle/_ref where present, and against golden digests made from it everywhere."""
import numpy as np
import pytest

from util import bits, sha

KEYS = ("node_bounds", "node_link", "sphere_geom", "sphere_light", "ambient")


def flat_diff(a, b):
    bad = []
    for k in KEYS:
        x, y = np.ascontiguousarray(a[k]), np.ascontiguousarray(b[k])
        if x.shape != y.shape:
            bad.append(k + ":shape")
        elif not np.array_equal(bits(x) if x.dtype == np.float32 else x, bits(y) if y.dtype == np.float32 else y):
            bad.append(k)
    return bad


def sample_points(ds):
    return ds.uniform01(0, 3 * 4096, ds.SEED + 5).reshape(4096, 3).astype(np.float32) * np.float32(300) - np.float32(20)


def test_light_script_golden(vv, port, golden, ds):
    meta, arr = golden
    lm = vv.LightManager()
    res = ds.apply_lights(lm, ds.light_script())
    assert np.array_equal(np.array(res, np.int32), arr["light_results"])
    assert lm.count() == meta["lights"]["count"]
    fl = lm.flatten()
    assert fl["node_bounds"].shape[0] == meta["lights"]["nodes"]
    assert {k: sha(v) for k, v in fl.items()} == meta["lights"]["flat"]
    pts = sample_points(ds)
    want = arr["light_at_points"]
    assert np.array_equal(bits(lm.get_light(pts)), bits(want))
    # the restated evaluator over the flat snapshot (what the GPU tests check the kernel against)
    assert np.array_equal(bits(port.lights(fl).get_light(pts)), bits(want))
    assert (want != fl["ambient"]).any(1).sum() > 1000


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_lights_equal_reference(vv, ref, port, seed):
    rng = np.random.default_rng(seed)
    lm, ol = vv.LightManager(), ref.lights()
    n = int(rng.integers(50, 400))
    centers = (rng.random((n, 3)) * 200 - 50).astype(np.float32)
    radii = (rng.random(n) * 40 + 1).astype(np.float32)
    colors = rng.random((n, 3)).astype(np.float32)
    amb = (rng.random(3) * 0.3).astype(np.float32)
    assert lm.set_ambient(amb) == ol.set_ambient(amb) == 1
    assert lm.set_ambient([1.5, 0, 0]) == ol.set_ambient([1.5, 0, 0]) == 0        # lights.c:81-89
    live = []
    for i in range(n):
        a, b = lm.insert(centers[i], radii[i], colors[i]), ol.insert(centers[i], radii[i], colors[i])
        assert a == b
        if a:
            live.append(i)
        if rng.random() < 0.25 and live:
            j = live.pop(int(rng.integers(len(live))))
            assert lm.delete(centers[j], radii[j]) == ol.delete(centers[j], radii[j]) == 1
        if i % 17 == 0:
            assert flat_diff(lm.flatten(), ol.flatten()) == [], i
    k = live[0]
    assert lm.insert(centers[k], radii[k], colors[0]) == ol.insert(centers[k], radii[k], colors[0]) == 0   # duplicate
    assert lm.insert([0, 0, 0], -1.0, [0, 0, 0]) == ol.insert([0, 0, 0], -1.0, [0, 0, 0]) == 0            # radius <= 0
    assert lm.insert([0, 0, 0], 1.0, [0, 2, 0]) == ol.insert([0, 0, 0], 1.0, [0, 2, 0]) == 0              # bad colour
    assert lm.delete([1e3, 1e3, 1e3], 1.0) == ol.delete([1e3, 1e3, 1e3], 1.0) == 0
    assert lm.count() == ol.count() == len(live)
    fl = lm.flatten()
    assert flat_diff(fl, ol.flatten()) == []
    pts = (rng.random((5000, 3)) * 220 - 60).astype(np.float32)
    want = ol.get_light(pts)
    assert np.array_equal(bits(lm.get_light(pts)), bits(want))
    for O in (ref, port):
        assert np.array_equal(bits(O.lights(fl).get_light(pts)), bits(want)), O.name
    for j in live:
        assert lm.delete(centers[j], radii[j]) == ol.delete(centers[j], radii[j]) == 1
    assert lm.count() == ol.count() == 0 and flat_diff(lm.flatten(), ol.flatten()) == []
    lm.insert([1, 2, 3], 4.0, [1, 1, 1])
    lm.clear()
    assert lm.count() == 0


def test_mtree_api(vv):
    """The exported M-tree functions directly (include/voxtrees/mtree.h)."""
    C = vv.api.C
    T = vv.libs.trees

    class Sphere(C.Structure):
        _fields_ = [("center", C.c_float * 4), ("color", C.c_float * 4), ("radius", C.c_float), ("sqr_radius", C.c_float)]

    T.vox_mtree_add_sphere.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Sphere)]
    T.vox_mtree_remove_sphere.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Sphere)]
    T.vox_mtree_items.argtypes = [C.c_void_p]
    T.vox_mtree_items.restype = C.c_uint
    T.vox_mtree_contains_sphere.argtypes = [C.c_void_p, C.POINTER(Sphere)]
    T.vox_mtree_contains_sphere.restype = C.c_void_p
    T.vox_mtree_destroy.argtypes = [C.c_void_p]
    CB = C.CFUNCTYPE(None, C.POINTER(Sphere), C.c_void_p)
    T.vox_mtree_spheres_containing_f.argtypes = [C.c_void_p, C.c_void_p, CB, C.c_void_p]
    root = C.c_void_p(None)
    spheres = [Sphere((C.c_float * 4)(10.0 * i, 0, 0, 0), (C.c_float * 4)(1, 1, 1, 0), 6.0, 0) for i in range(40)]
    for s in spheres:
        assert T.vox_mtree_add_sphere(C.byref(root), C.byref(s)) == 1
    assert T.vox_mtree_add_sphere(C.byref(root), C.byref(spheres[3])) == 0
    assert T.vox_mtree_items(root) == 40
    assert T.vox_mtree_contains_sphere(root, C.byref(spheres[7]))
    seen = []
    cb = CB(lambda sp, arg: seen.append(sp.contents.center[0]))
    dot = (C.c_float * 4)(55.0, 0, 0, 0)
    T.vox_mtree_spheres_containing_f(root, dot, cb, None)
    assert sorted(seen) == [50.0, 60.0]            # |55-50| = |55-60| = 5 < 6
    for s in spheres:
        assert T.vox_mtree_remove_sphere(C.byref(root), C.byref(s)) == 1
    assert root.value is None and T.vox_mtree_items(root) == 0
    T.vox_mtree_destroy(root)
