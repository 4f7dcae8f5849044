"""ctypes wrapper around the oracle API (oracle/oracle_api.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(voxvision_b200) never imports this module.

Two interchangeable implementations sit behind the same API:
  Oracle("port")  -> oracle/_build/libvoxoracle.so  (CPU restatement, vox_oracle.c)
  Oracle("ref")   -> oracle/_ref/libvoxref_scalar.so (reference sources, scalar,
                     -O2 -ffp-contract=off: the oracle of record)
  Oracle("ref_fast") -> oracle/_ref/libvoxref_fast.so (-O2 -march=native, timing only)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

LIBS = {
    "port": os.path.join(HERE, "_build", "libvoxoracle.so"),
    "ref": os.path.join(HERE, "_ref", "libvoxref_scalar.so"),
    "ref_fast": os.path.join(HERE, "_ref", "libvoxref_fast.so"),
    "ref_sse": os.path.join(HERE, "_ref", "libvoxref_sse.so"),
}

QUALITY_BEST, QUALITY_FAST, QUALITY_ADAPTIVE = 0, 1, 2
QUALITY_RAY_MERGE, QUALITY_RAY_MERGE_ACCURATE = 4, 8
CAMERA_SIMPLE, CAMERA_DOOM = 0, 1
LEAF, DENSE_LEAF, OVERFLOW = 1, 2, 4


class TreeInfo(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("n_nodes", "n_inner", "n_leaf", "n_dense", "n_overflow", "n_null", "n_dots")] + \
               [("depth", C.c_uint32)]


class Flat(C.Structure):
    _fields_ = [("bb_min", C.c_void_p), ("bb_max", C.c_void_p), ("flags", C.c_void_p),
                ("dots_num", C.c_void_p), ("center", C.c_void_p), ("first_child", C.c_void_p),
                ("child_mask", C.c_void_p), ("dots_off", C.c_void_p), ("dots", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("calls", "null_calls", "inner", "leaf", "dense", "dots_tested", "plane_tests")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class PixFmt(C.Structure):
    _fields_ = [("amask", C.c_uint32),
                ("rloss", C.c_uint8), ("gloss", C.c_uint8), ("bloss", C.c_uint8), ("pad0", C.c_uint8),
                ("rshift", C.c_uint8), ("gshift", C.c_uint8), ("bshift", C.c_uint8), ("pad1", C.c_uint8)]


class LightsFlat(C.Structure):
    _fields_ = [("ambient", C.c_float * 3), ("n_nodes", C.c_uint32), ("n_spheres", C.c_uint32),
                ("node_bounds", C.c_void_p), ("node_link", C.c_void_p), ("sphere_geom", C.c_void_p),
                ("sphere_light", C.c_void_p)]


def pixfmt_xrgb8888():
    """SDL_PIXELFORMAT_RGB888 (the usual window-surface format): R<<16 | G<<8 | B."""
    return PixFmt(0, 0, 0, 0, 0, 16, 8, 0, 0)


def build(which="port"):
    """Build the oracle libraries (make port / make ref)."""
    target = "port" if which == "port" else "ref"
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


def available(which):
    return os.path.exists(LIBS[which])


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Oracle:
    def __init__(self, which="port"):
        path = LIBS[which]
        if not os.path.exists(path):
            raise FileNotFoundError(f"oracle library {path} missing: run `make -C oracle {which}`")
        self.which = which
        L = self.lib = C.CDLL(path)
        vp, f, i, u, sz = C.c_void_p, C.c_float, C.c_int, C.c_uint, C.c_size_t
        L.orc_impl_name.restype = C.c_char_p
        L.orc_set_voxel.argtypes = [f, f, f]
        L.orc_get_voxel.argtypes = [vp]
        L.orc_tree_build.restype = vp
        L.orc_tree_build.argtypes = [vp, sz]
        L.orc_tree_build_by_insertion.restype = vp
        L.orc_tree_build_by_insertion.argtypes = [vp, sz]
        L.orc_tree_free.argtypes = [vp]
        L.orc_tree_from_flat.restype = vp
        L.orc_tree_from_flat.argtypes = [C.POINTER(Flat), C.c_uint64]
        L.orc_tree_edit.restype = i
        L.orc_tree_edit.argtypes = [vp, vp, vp, sz, vp]
        L.orc_tree_rebuild.restype = vp
        L.orc_tree_rebuild.argtypes = [vp]
        L.orc_tree_info_get.argtypes = [vp, C.POINTER(TreeInfo)]
        L.orc_tree_flatten.argtypes = [vp, C.POINTER(Flat)]
        L.orc_tree_reordered_set.argtypes = [vp, vp]
        L.orc_ray_batch.argtypes = [vp, C.c_uint32, vp, vp, sz, vp, vp, vp, C.POINTER(Stats), i]
        L.orc_ball_batch.argtypes = [vp, vp, vp, sz, vp, i]
        L.orc_camera_new.restype = vp
        L.orc_camera_new.argtypes = [i]
        L.orc_camera_free.argtypes = [vp]
        for name in ("set_position", "set_rotation", "look_at", "rotate", "move", "get_position"):
            getattr(L, "orc_camera_" + name).argtypes = [vp, vp]
        L.orc_camera_set_fov.argtypes = [vp, f]
        L.orc_camera_set_window_size.argtypes = [vp, i, i]
        L.orc_camera_screen2world.argtypes = [vp, vp, i, i]
        L.orc_camera_get_state.argtypes = [vp, vp]
        L.orc_camera_rays.argtypes = [vp, i, i, vp, vp]
        L.orc_render.restype = i
        L.orc_lights_new.restype = vp
        L.orc_lights_from_flat.restype = vp
        L.orc_lights_from_flat.argtypes = [C.POINTER(LightsFlat)]
        L.orc_lights_free.argtypes = [vp]
        L.orc_lights_insert.restype = i; L.orc_lights_insert.argtypes = [vp, vp, f, vp]
        L.orc_lights_delete.restype = i; L.orc_lights_delete.argtypes = [vp, vp, f]
        L.orc_lights_clear.argtypes = [vp]
        L.orc_lights_count.restype = i; L.orc_lights_count.argtypes = [vp]
        L.orc_lights_set_ambient.restype = i; L.orc_lights_set_ambient.argtypes = [vp, vp]
        L.orc_lights_flatten.argtypes = [vp, C.POINTER(LightsFlat)]
        L.orc_lights_get.argtypes = [vp, vp, sz, vp]
        L.orc_render_lit.restype = i
        L.orc_render_lit.argtypes = [vp, vp, i, i, u, vp, C.POINTER(PixFmt), vp, vp, i]
        L.orc_render.argtypes = [vp, vp, i, i, u, vp, C.POINTER(PixFmt), vp, vp, vp, i]

    @property
    def name(self):
        return self.lib.orc_impl_name().decode()

    def set_voxel(self, x, y, z):
        self.lib.orc_set_voxel(x, y, z)

    def tree(self, xyz, by_insertion=False):
        return OracleTree(self, xyz, by_insertion)

    def tree_from_flat(self, a):
        """A tree from canonical flat arrays (dict as OracleTree.flatten() / vv.Tree.export())."""
        keep = {k: np.ascontiguousarray(a[k]) for k, _ in Flat._fields_}
        fl = Flat(*[keep[k].ctypes.data for k, _ in Flat._fields_])
        return OracleTree(self, None, _handle=self.lib.orc_tree_from_flat(C.byref(fl), keep["flags"].shape[0]))

    def camera(self, kind=CAMERA_SIMPLE):
        return OracleCamera(self, kind)

    def lights(self, flat=None):
        """A light manager driven through the reference's API (reference glue only), or -- with
        `flat` (dict as vv.LightManager.flatten()) -- a given flat light set."""
        return OracleLights(self, flat)


class OracleTree:
    def __init__(self, orc, xyz, by_insertion=False, _handle=None):
        self.orc = orc
        if _handle is not None:
            self.h, self.n = _handle, 0
            return
        xyz = _f32(xyz).reshape(-1, 3)
        self.n = xyz.shape[0]
        fn = orc.lib.orc_tree_build_by_insertion if by_insertion else orc.lib.orc_tree_build
        self.h = fn(xyz.ctypes.data, self.n)
        if not self.h:
            raise RuntimeError("oracle could not build the tree (insertion is reference-only)")

    def __del__(self):
        if getattr(self, "h", None):
            self.orc.lib.orc_tree_free(self.h)
            self.h = None

    def edit(self, insert, xyz):
        """vox_insert_voxel (insert[i] true) / vox_delete_voxel per row of xyz; their return values."""
        xyz = _f32(xyz).reshape(-1, 3)
        ins = np.ascontiguousarray(np.broadcast_to(np.asarray(insert, np.uint8), (xyz.shape[0],)))
        res = np.zeros(xyz.shape[0], np.uint8)
        if not self.orc.lib.orc_tree_edit(self.h, ins.ctypes.data, xyz.ctypes.data, xyz.shape[0], res.ctypes.data):
            raise RuntimeError("this oracle cannot edit trees (reference glue only)")
        return res

    def rebuild(self):
        h = self.orc.lib.orc_tree_rebuild(self.h)
        if not h:
            raise RuntimeError("this oracle cannot rebuild trees (reference glue only)")
        return OracleTree(self.orc, None, _handle=h)

    def info(self):
        ti = TreeInfo()
        self.orc.lib.orc_tree_info_get(self.h, C.byref(ti))
        return ti

    def reordered_set(self):
        out = np.empty((self.n, 3), np.float32)
        self.orc.lib.orc_tree_reordered_set(self.h, out.ctypes.data)
        return out

    def flatten(self):
        """dict of numpy arrays in the canonical breadth-first layout."""
        ti = self.info()
        n, nd = int(ti.n_nodes), int(ti.n_dots)
        a = {
            "bb_min": np.zeros((n, 3), np.float32), "bb_max": np.zeros((n, 3), np.float32),
            "flags": np.zeros(n, np.uint32), "dots_num": np.zeros(n, np.uint32),
            "center": np.zeros((n, 3), np.float32), "first_child": np.zeros(n, np.uint32),
            "child_mask": np.zeros(n, np.uint32), "dots_off": np.zeros(n, np.uint64),
            "dots": np.zeros((nd, 3), np.float32),
        }
        fl = Flat(*[a[k].ctypes.data for k, _ in Flat._fields_])
        self.orc.lib.orc_tree_flatten(self.h, C.byref(fl))
        a["depth"] = int(ti.depth)
        return a

    def rays(self, origins, dirs, start_node=0, threads=1, want_stats=False):
        o, d = _f32(origins).reshape(-1, 3), _f32(dirs).reshape(-1, 3)
        n = o.shape[0]
        leaf = np.full(n, -1, np.int32)
        hit = np.zeros((n, 3), np.float32)
        dot = np.full(n, -1, np.int64)
        st = Stats()
        self.orc.lib.orc_ray_batch(self.h, start_node, o.ctypes.data, d.ctypes.data, n,
                                   leaf.ctypes.data, hit.ctypes.data, dot.ctypes.data,
                                   C.byref(st), threads)
        if want_stats:
            return leaf, hit, dot, st.as_dict()
        return leaf, hit, dot

    def balls(self, centers, radii, threads=1):
        c = _f32(centers).reshape(-1, 3)
        r = _f32(radii).reshape(-1)
        out = np.zeros(c.shape[0], np.uint8)
        self.orc.lib.orc_ball_batch(self.h, c.ctypes.data, r.ctypes.data, c.shape[0],
                                    out.ctypes.data, threads)
        return out

    def render_lit(self, cam, w, h, lights, quality=QUALITY_BEST, texture=None, fmt=None, threads=1):
        """vox_render with a light manager attached (renderer.c:41-49) -> pixels."""
        texture = np.ascontiguousarray(texture, np.uint8)
        fmt = fmt or pixfmt_xrgb8888()
        pixels = np.zeros((h, w), np.uint32)
        ok = self.orc.lib.orc_render_lit(self.h, cam.h, w, h, quality, texture.ctypes.data, C.byref(fmt),
                                         lights.h if lights is not None else None, pixels.ctypes.data, threads)
        return pixels if ok else None

    def render(self, cam, w, h, quality=QUALITY_BEST, texture=None, fmt=None, threads=1,
               want_hits=True):
        if texture is None:
            texture = np.arange(32 * 32 * 32, dtype=np.uint32).astype(np.uint8)
        texture = np.ascontiguousarray(texture, np.uint8)
        fmt = fmt or pixfmt_xrgb8888()
        pixels = np.zeros((h, w), np.uint32)
        leaf = np.full((h, w), -1, np.int32) if want_hits else None
        hit = np.zeros((h, w, 3), np.float32) if want_hits else None
        ok = self.orc.lib.orc_render(self.h, cam.h, w, h, quality, texture.ctypes.data,
                                     C.byref(fmt), pixels.ctypes.data,
                                     leaf.ctypes.data if want_hits else None,
                                     hit.ctypes.data if want_hits else None, threads)
        if not ok:
            return None
        return (pixels, leaf, hit) if want_hits else pixels


class OracleLights:
    def __init__(self, orc, flat=None):
        self.orc = orc
        if flat is None:
            self.h = orc.lib.orc_lights_new()
            if not self.h:
                raise RuntimeError("this oracle cannot build light managers (reference glue only)")
        else:
            keep = {k: np.ascontiguousarray(flat[k]) for k in ("node_bounds", "node_link", "sphere_geom", "sphere_light")}
            f = LightsFlat((C.c_float * 3)(*[float(x) for x in flat["ambient"]]), keep["node_bounds"].shape[0],
                           keep["sphere_geom"].shape[0], keep["node_bounds"].ctypes.data, keep["node_link"].ctypes.data,
                           keep["sphere_geom"].ctypes.data, keep["sphere_light"].ctypes.data)
            self.h = orc.lib.orc_lights_from_flat(C.byref(f))

    def __del__(self):
        if getattr(self, "h", None):
            self.orc.lib.orc_lights_free(self.h)
            self.h = None

    def insert(self, center, radius, color):
        c, col = _f32(center), _f32(color)
        return int(self.orc.lib.orc_lights_insert(self.h, c.ctypes.data, radius, col.ctypes.data))

    def delete(self, center, radius):
        c = _f32(center)
        return int(self.orc.lib.orc_lights_delete(self.h, c.ctypes.data, radius))

    def clear(self): self.orc.lib.orc_lights_clear(self.h)
    def count(self): return int(self.orc.lib.orc_lights_count(self.h))
    def set_ambient(self, color):
        c = _f32(color)
        return int(self.orc.lib.orc_lights_set_ambient(self.h, c.ctypes.data))

    def flatten(self):
        f = LightsFlat()
        self.orc.lib.orc_lights_flatten(self.h, C.byref(f))
        nn, ns = f.n_nodes, f.n_spheres
        a = {"node_bounds": np.zeros((nn, 4), np.float32), "node_link": np.zeros((nn, 4), np.uint32),
             "sphere_geom": np.zeros((ns, 4), np.float32), "sphere_light": np.zeros((ns, 4), np.float32)}
        f = LightsFlat((C.c_float * 3)(), 0, 0, a["node_bounds"].ctypes.data, a["node_link"].ctypes.data,
                       a["sphere_geom"].ctypes.data, a["sphere_light"].ctypes.data)
        self.orc.lib.orc_lights_flatten(self.h, C.byref(f))
        a["ambient"] = np.array(list(f.ambient), np.float32)
        return a

    def get_light(self, points):
        p = _f32(points).reshape(-1, 3)
        out = np.zeros((p.shape[0], 3), np.float32)
        self.orc.lib.orc_lights_get(self.h, p.ctypes.data, p.shape[0], out.ctypes.data)
        return out


class OracleCamera:
    def __init__(self, orc, kind=CAMERA_SIMPLE):
        self.orc, self.kind = orc, kind
        self.h = orc.lib.orc_camera_new(kind)

    def __del__(self):
        if getattr(self, "h", None):
            self.orc.lib.orc_camera_free(self.h)
            self.h = None

    def _v(self, fn, v):
        a = _f32(v).reshape(3)
        getattr(self.orc.lib, "orc_camera_" + fn)(self.h, a.ctypes.data)

    def set_position(self, p): self._v("set_position", p)
    def set_rotation(self, a): self._v("set_rotation", a)
    def look_at(self, p): self._v("look_at", p)
    def rotate(self, d): self._v("rotate", d)
    def move(self, d): self._v("move", d)
    def set_fov(self, fov): self.orc.lib.orc_camera_set_fov(self.h, fov)
    def set_window_size(self, w, h): self.orc.lib.orc_camera_set_window_size(self.h, w, h)

    def get_position(self):
        a = np.zeros(3, np.float32)
        self.orc.lib.orc_camera_get_position(self.h, a.ctypes.data)
        return a

    def screen2world(self, sx, sy):
        a = np.zeros(3, np.float32)
        self.orc.lib.orc_camera_screen2world(self.h, a.ctypes.data, sx, sy)
        return a

    def rays(self, w, h):
        """All primary rays of a w x h frame (origins, dirs), row-major."""
        o, d = np.empty((h * w, 3), np.float32), np.empty((h * w, 3), np.float32)
        self.orc.lib.orc_camera_rays(self.h, w, h, o.ctypes.data, d.ctypes.data)
        return o, d

    def state(self):
        a = np.zeros(11, np.float32)
        self.orc.lib.orc_camera_get_state(self.h, a.ctypes.data)
        return a
