"""ctypes binding over the C API declared in include/ (no compute here)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(HERE, "lib")
INCLUDE_DIR = os.path.join(os.path.dirname(HERE), "include")

QUALITY_BEST, QUALITY_FAST, QUALITY_ADAPTIVE = 0, 1, 2
QUALITY_RAY_MERGE, QUALITY_RAY_MERGE_ACCURATE = 4, 8


class VoxError(RuntimeError):
    pass


class _TreeInfo(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_nodes", "n_inner", "n_leaf", "n_dense", "n_overflow", "n_dots")] + \
               [("depth", C.c_uint32), ("node_id", C.c_uint32)]


class _TreeArrays(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("bb_min", "bb_max", "flags", "dots_num", "center",
                                          "first_child", "child_mask", "dots_off", "dots")]


class _Box(C.Structure):
    _fields_ = [("min", C.c_float * 4), ("max", C.c_float * 4)]


class _PixelFormat(C.Structure):
    _fields_ = [("format", C.c_uint32), ("palette", C.c_void_p),
                ("BitsPerPixel", C.c_uint8), ("BytesPerPixel", C.c_uint8), ("padding", C.c_uint8 * 2),
                ("Rmask", C.c_uint32), ("Gmask", C.c_uint32), ("Bmask", C.c_uint32), ("Amask", C.c_uint32),
                ("Rloss", C.c_uint8), ("Gloss", C.c_uint8), ("Bloss", C.c_uint8), ("Aloss", C.c_uint8),
                ("Rshift", C.c_uint8), ("Gshift", C.c_uint8), ("Bshift", C.c_uint8), ("Ashift", C.c_uint8),
                ("refcount", C.c_int), ("next", C.c_void_p)]


class _Rect(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("w", C.c_int), ("h", C.c_int)]


class _Surface(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("format", C.POINTER(_PixelFormat)),
                ("w", C.c_int), ("h", C.c_int), ("pitch", C.c_int), ("pixels", C.c_void_p),
                ("userdata", C.c_void_p), ("locked", C.c_int), ("list_blitmap", C.c_void_p),
                ("clip_rect", _Rect), ("map", C.c_void_p), ("refcount", C.c_int)]


class _Ctx(C.Structure):
    _fields_ = [("window", C.c_void_p), ("surface", C.POINTER(_Surface)), ("scene", C.c_void_p),
                ("camera", C.c_void_p), ("light_manager", C.c_void_p)]


class _CameraDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("position", C.c_float * 3), ("mul", C.c_float), ("xsub", C.c_float),
                ("ysub", C.c_float), ("rotation", C.c_float * 4), ("sinphi", C.c_float),
                ("cosphi", C.c_float), ("k", C.c_float)]


_FN = C.CFUNCTYPE
_vp, _f, _i = C.c_void_p, C.c_float, C.c_int


class _CameraIface(C.Structure):
    _fields_ = [
        ("set_property_number", _FN(None, _vp, C.c_char_p, _f)),
        ("set_property_dot", _FN(None, _vp, C.c_char_p, _vp)),
        ("screen2world", _FN(None, _vp, _vp, _i, _i)),
        ("rotate_camera", _FN(None, _vp, _vp)),
        ("move_camera", _FN(None, _vp, _vp)),
        ("look_at", _FN(None, _vp, _vp)),
        ("get_position", _FN(None, _vp, _vp)),
        ("set_window_size", _FN(None, _vp, _i, _i)),
        ("construct_camera", _FN(_vp, _vp)),
        ("destroy_camera", _FN(None, _vp)),
        ("describe", _FN(_i, _vp, C.POINTER(_CameraDesc))),
    ]


class _CameraObj(C.Structure):
    _fields_ = [("iface", C.POINTER(_CameraIface))]


class GpuCamera(C.Structure):
    """voxgpu_camera (include/voxgpu.h)."""
    _fields_ = [("kind", C.c_int), ("position", C.c_float * 3), ("mul", C.c_float), ("xsub", C.c_float),
                ("ysub", C.c_float), ("rotation", C.c_float * 4), ("sinphi", C.c_float),
                ("cosphi", C.c_float), ("k", C.c_float), ("dir_table", C.c_void_p)]


class GpuPixFmt(C.Structure):
    _fields_ = [("amask", C.c_uint32), ("rloss", C.c_uint8), ("gloss", C.c_uint8), ("bloss", C.c_uint8),
                ("pad0", C.c_uint8), ("rshift", C.c_uint8), ("gshift", C.c_uint8), ("bshift", C.c_uint8),
                ("pad1", C.c_uint8)]


class GpuFrame(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("y0", C.c_int), ("y1", C.c_int), ("pitch_pixels", C.c_int),
                ("quality", C.c_uint), ("fmt", GpuPixFmt), ("lights", C.c_void_p),
                ("interleave_world", C.c_int), ("interleave_rank", C.c_int), ("start_node", C.c_uint32)]


class GpuLightsFlat(C.Structure):
    """voxgpu_lights_flat (include/voxgpu.h)."""
    _fields_ = [("ambient", C.c_float * 3), ("n_nodes", C.c_uint32), ("n_spheres", C.c_uint32),
                ("node_bounds", C.c_void_p), ("node_link", C.c_void_p), ("sphere_geom", C.c_void_p),
                ("sphere_light", C.c_void_p)]


class GpuStats(C.Structure):
    _fields_ = [(n, C.c_ulonglong) for n in ("rays", "hits", "calls", "inner", "leaf", "dense",
                                             "dots_tested", "expansions", "max_stack")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class _Libs:
    """Loads the shared libraries once; raises VoxError if they are not built."""

    def __init__(self):
        self._loaded = False

    def _load(self):
        if self._loaded:
            return
        paths = {n: os.path.join(LIB_DIR, n) for n in ("libvoxgpu.so", "libvoxtrees.so", "libvoxrnd.so")}
        for p in paths.values():
            if not os.path.exists(p):
                raise VoxError(f"{p} is missing: build with `make -C voxvision_b200` "
                               f"(or __graft_entry__.build()); there is no fallback implementation")
        os.environ.setdefault("VOXVISION_MODULES", LIB_DIR)
        # RTLD_LOCAL on purpose: the libraries export the reference's own symbol
        # names (that is what drop-in means); they must not interpose on any
        # other library in the process.  They find each other through DT_NEEDED.
        self.gpu = C.CDLL(paths["libvoxgpu.so"])
        self.trees = C.CDLL(paths["libvoxtrees.so"])
        self.rnd = C.CDLL(paths["libvoxrnd.so"])
        self._proto()
        self._loaded = True

    def __getattr__(self, name):
        if name in ("gpu", "trees", "rnd"):
            self._load()
            return self.__dict__[name]
        raise AttributeError(name)

    def _proto(self):
        g, t, r = self.__dict__["gpu"], self.__dict__["trees"], self.__dict__["rnd"]
        sz, u32, u64, l = C.c_size_t, C.c_uint32, C.c_uint64, C.c_long
        g.voxgpu_last_error.restype = C.c_char_p
        g.voxgpu_malloc.restype = _vp; g.voxgpu_malloc.argtypes = [sz]
        g.voxgpu_free.argtypes = [_vp]
        g.voxgpu_host_alloc.restype = _vp; g.voxgpu_host_alloc.argtypes = [sz]
        g.voxgpu_host_free.argtypes = [_vp]
        g.voxgpu_host_device_pointer.restype = _vp; g.voxgpu_host_device_pointer.argtypes = [_vp]
        g.voxgpu_memcpy_h2d.argtypes = [_vp, _vp, sz, _vp]
        g.voxgpu_memcpy_d2h.argtypes = [_vp, _vp, sz, _vp]
        g.voxgpu_memset.argtypes = [_vp, _i, sz, _vp]
        g.voxgpu_synchronize.argtypes = [_vp]
        g.voxgpu_stream_create.restype = _vp
        g.voxgpu_stream_destroy.argtypes = [_vp]
        g.voxgpu_stream_join.argtypes = [_vp, _vp]
        g.voxgpu_ipc_export.argtypes = [_vp, _vp]
        g.voxgpu_ipc_open.restype = _vp; g.voxgpu_ipc_open.argtypes = [_vp]
        g.voxgpu_ipc_close.argtypes = [_vp]
        g.voxgpu_tree_device_bytes.restype = sz; g.voxgpu_tree_device_bytes.argtypes = [_vp]
        g.voxgpu_lights_upload.restype = _vp; g.voxgpu_lights_upload.argtypes = [C.POINTER(GpuLightsFlat)]
        g.voxgpu_lights_free.argtypes = [_vp]
        g.voxgpu_rays_device.argtypes = [_vp, u32, _vp, _vp, sz, _vp, _vp, _vp, _vp]
        g.voxgpu_rays_stats.argtypes = [_vp, u32, _vp, _vp, sz, _vp, _vp, _vp, C.POINTER(GpuStats)]
        g.voxgpu_balls_device.argtypes = [_vp, _vp, _vp, sz, _vp, _vp]
        g.voxgpu_render.argtypes = [_vp, C.POINTER(GpuCamera), C.POINTER(GpuFrame), _vp, _vp, _vp, _vp, _vp]
        g.voxgpu_set_cull.argtypes = [_i]
        g.voxgpu_set_tile_feedback.argtypes = [_i]
        g.voxgpu_set_variant.argtypes = [_i]
        g.voxgpu_set_rays_sort.argtypes = [_i]
        g.voxgpu_set_rays_refill.argtypes = [_i]
        g.voxgpu_volume_extract.argtypes = [_vp, C.POINTER(C.c_uint * 3), C.c_uint, C.c_uint, C.c_ulonglong,
                                            C.POINTER(C.c_float * 3), C.POINTER(_vp), C.POINTER(sz)]
        g.voxgpu_set_trace.argtypes = [_vp]
        g.voxgpu_present_create.restype = _vp; g.voxgpu_present_create.argtypes = [_i, _i, _i]
        g.voxgpu_present_destroy.argtypes = [_vp]
        g.voxgpu_present_render.argtypes = [_vp, _vp, C.POINTER(GpuCamera), C.POINTER(GpuFrame), _vp, _vp, sz]
        g.voxgpu_present_wait.argtypes = [_vp]
        g.voxgpu_launch_count.restype = C.c_ulonglong
        g.voxgpu_present_bytes_copied.restype = C.c_ulonglong
        g.voxgpu_set_present_spans.argtypes = [C.c_int, C.c_int, C.c_longlong]
        g.voxgpu_set_present_spans.restype = None
        g.voxgpu_timer_create.restype = _vp
        g.voxgpu_timer_destroy.argtypes = [_vp]
        g.voxgpu_timer_start.argtypes = [_vp, _vp]
        g.voxgpu_timer_stop.argtypes = [_vp, _vp]
        g.voxgpu_timer_elapsed_ms.restype = _f; g.voxgpu_timer_elapsed_ms.argtypes = [_vp]

        t.vox_last_error.restype = C.c_char_p
        t.vox_make_tree.restype = _vp; t.vox_make_tree.argtypes = [_vp, sz]
        t.vox_destroy_tree.argtypes = [_vp]
        t.vox_voxels_in_tree.restype = sz; t.vox_voxels_in_tree.argtypes = [_vp]
        t.vox_bounding_box.argtypes = [_vp, C.POINTER(_Box)]
        t.vox_make_dense_leaf.restype = _vp; t.vox_make_dense_leaf.argtypes = [C.POINTER(_Box)]
        t.vox_set_voxel_size.argtypes = [_f, _f, _f]
        t.vox_insert_voxel.restype = _i; t.vox_insert_voxel.argtypes = [C.POINTER(_vp), _vp]
        t.vox_delete_voxel.restype = _i; t.vox_delete_voxel.argtypes = [C.POINTER(_vp), _vp]
        t.vox_insert_voxel_coord.restype = _i; t.vox_insert_voxel_coord.argtypes = [C.POINTER(_vp), _f, _f, _f]
        t.vox_delete_voxel_coord.restype = _i; t.vox_delete_voxel_coord.argtypes = [C.POINTER(_vp), _f, _f, _f]
        t.vox_rebuild_tree.restype = _vp; t.vox_rebuild_tree.argtypes = [_vp]
        t.vox_dump_tree.argtypes = [_vp]
        t.vox_tree_get_info.argtypes = [_vp, C.POINTER(_TreeInfo)]
        t.vox_tree_export.argtypes = [_vp, C.POINTER(_TreeArrays)]
        t.vox_tree_import.restype = _vp; t.vox_tree_import.argtypes = [C.POINTER(_TreeArrays), u64, u64]
        t.vox_tree_node.restype = _vp; t.vox_tree_node.argtypes = [_vp, u32]
        t.vox_ray_tree_intersection.restype = _vp
        t.vox_ray_tree_intersection.argtypes = [_vp, _vp, _vp, _vp]
        t.vox_tree_ball_collidep.restype = _i; t.vox_tree_ball_collidep.argtypes = [_vp, _vp, _f]
        t.vox_ray_tree_intersection_batch.restype = l
        t.vox_ray_tree_intersection_batch.argtypes = [_vp, _vp, _vp, sz, _vp, _vp, _vp]
        t.vox_tree_ball_collidep_batch.restype = l
        t.vox_tree_ball_collidep_batch.argtypes = [_vp, _vp, _vp, sz, _vp]
        t.vox_tree_device_handle.restype = _vp; t.vox_tree_device_handle.argtypes = [_vp]
        t.vox_read_raw_data_threshold.restype = _vp
        t.vox_read_raw_data_threshold.argtypes = [C.c_char_p, C.POINTER(C.c_uint * 3), C.c_uint, C.c_uint,
                                                  C.POINTER(C.c_char_p)]
        t.vox_read_raw_data_range.restype = _vp
        t.vox_read_raw_data_range.argtypes = [C.c_char_p, C.POINTER(C.c_uint * 3), C.c_uint, C.c_uint, C.c_uint,
                                              C.POINTER(C.c_char_p)]
        t.vox_abs_metric.restype = _f; t.vox_abs_metric.argtypes = [_vp, _vp]
        t.vox_sqr_metric.restype = _f; t.vox_sqr_metric.argtypes = [_vp, _vp]
        t.vox_sqr_norm.restype = _f; t.vox_sqr_norm.argtypes = [_vp]

        r.vox_camera_methods.restype = C.POINTER(_CameraIface); r.vox_camera_methods.argtypes = [C.c_char_p]
        r.vox_make_surface.restype = C.POINTER(_Surface); r.vox_make_surface.argtypes = [C.c_uint, C.c_uint]
        r.vox_free_surface.argtypes = [C.POINTER(_Surface)]
        r.vox_map_rgb.restype = u32
        r.vox_map_rgb.argtypes = [C.POINTER(_PixelFormat), C.c_uint8, C.c_uint8, C.c_uint8]
        r.vox_make_context_from_surface.restype = C.POINTER(_Ctx)
        r.vox_make_context_from_surface.argtypes = [C.POINTER(_Surface)]
        r.vox_make_context_and_window.restype = C.POINTER(_Ctx)
        r.vox_make_context_and_window.argtypes = [C.c_uint, C.c_uint]
        r.vox_destroy_context.argtypes = [C.POINTER(_Ctx)]
        r.vox_context_set_scene.argtypes = [C.POINTER(_Ctx), _vp]
        r.vox_context_set_camera.argtypes = [C.POINTER(_Ctx), _vp]
        r.vox_context_set_quality.restype = _i; r.vox_context_set_quality.argtypes = [C.POINTER(_Ctx), C.c_uint]
        r.vox_context_set_texture.argtypes = [C.POINTER(_Ctx), _vp]
        r.vox_context_get_texture.restype = _vp; r.vox_context_get_texture.argtypes = [C.POINTER(_Ctx)]
        r.vox_context_last_kernel_ms.restype = _f; r.vox_context_last_kernel_ms.argtypes = [C.POINTER(_Ctx)]
        r.vox_context_set_devices.restype = _i; r.vox_context_set_devices.argtypes = [C.POINTER(_Ctx), _i]
        r.vox_context_set_light_manager.argtypes = [C.POINTER(_Ctx), _vp]
        r.vox_make_light_manager.restype = _vp
        r.vox_destroy_light_manager.argtypes = [_vp]
        r.vox_insert_shadowless_light.restype = _i; r.vox_insert_shadowless_light.argtypes = [_vp, _vp, _f, _vp]
        r.vox_delete_shadowless_light.restype = _i; r.vox_delete_shadowless_light.argtypes = [_vp, _vp, _f]
        r.vox_delete_shadowless_lights.argtypes = [_vp]
        r.vox_shadowless_lights_number.restype = _i; r.vox_shadowless_lights_number.argtypes = [_vp]
        r.vox_set_ambient_light.restype = _i; r.vox_set_ambient_light.argtypes = [_vp, _vp]
        r.vox_get_light.argtypes = [_vp, _vp, _vp]
        r.vox_light_manager_flatten.argtypes = [_vp, C.POINTER(u32), C.POINTER(u32), _vp, _vp, _vp, _vp]
        r.vox_render.argtypes = [C.POINTER(_Ctx)]
        r.vox_redraw.argtypes = [C.POINTER(_Ctx)]
        r.vox_make_cd.restype = _vp
        r.vox_destroy_cd.argtypes = [_vp]
        r.vox_cd_attach_camera.argtypes = [_vp, _vp, _f]
        r.vox_cd_attach_context.argtypes = [_vp, C.POINTER(_Ctx)]
        r.vox_cd_collide.argtypes = [_vp]
        r.vox_rotate_vector.argtypes = [_vp, _vp, _vp]
        r.vox_quat_mul.argtypes = [_vp, _vp, _vp]


libs = _Libs()


def last_error():
    """Error text of the last voxtrees/voxrnd call on this thread ("" if it succeeded)."""
    return libs.trees.vox_last_error().decode()


def device_count():
    return libs.gpu.voxgpu_device_count()


def gpu_available():
    return device_count() > 0


def set_device(d):
    if libs.gpu.voxgpu_set_device(d):
        raise VoxError(libs.gpu.voxgpu_last_error().decode())


def launch_count():
    return int(libs.gpu.voxgpu_launch_count())


def set_voxel_size(x, y, z):
    libs.trees.vox_set_voxel_size(x, y, z)


def get_voxel_size():
    v = (C.c_float * 4).in_dll(libs.trees, "vox_voxel")
    return np.array([v[0], v[1], v[2]], np.float32)


def dots4(xyz):
    """(n,3) floats -> 16-byte aligned (n,4) float32 array (the vox_dot[] ABI)."""
    xyz = np.asarray(xyz, np.float32).reshape(-1, 3)
    n = xyz.shape[0]
    raw = np.zeros(n * 16 + 64, np.uint8)
    off = (-raw.ctypes.data) % 64
    out = raw[off:off + n * 16].view(np.float32).reshape(n, 4)
    out[:, :3] = xyz
    return out


class DeviceBuffer:
    """A cudaMalloc'ed buffer (through voxgpu_malloc)."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = libs.gpu.voxgpu_malloc(self.nbytes)
        if not self.ptr:
            raise VoxError(libs.gpu.voxgpu_last_error().decode())

    @classmethod
    def from_array(cls, a, stream=None):
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        b.upload(a, stream)
        return b

    def upload(self, a, stream=None):
        a = np.ascontiguousarray(a)
        if libs.gpu.voxgpu_memcpy_h2d(self.ptr, a.ctypes.data, a.nbytes, stream) or \
                libs.gpu.voxgpu_synchronize(stream):
            raise VoxError(libs.gpu.voxgpu_last_error().decode())

    def download(self, dtype, shape):
        out = np.empty(shape, dtype)
        if libs.gpu.voxgpu_memcpy_d2h(out.ctypes.data, self.ptr, out.nbytes, None) or \
                libs.gpu.voxgpu_synchronize(None):
            raise VoxError(libs.gpu.voxgpu_last_error().decode())
        return out

    def free(self):
        if self.ptr:
            libs.gpu.voxgpu_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Timer:
    """CUDA-event timer on a stream (voxgpu_timer)."""

    def __init__(self):
        self.h = libs.gpu.voxgpu_timer_create()
        if not self.h:
            raise VoxError(libs.gpu.voxgpu_last_error().decode())

    def start(self, stream=None): libs.gpu.voxgpu_timer_start(self.h, stream)
    def stop(self, stream=None): libs.gpu.voxgpu_timer_stop(self.h, stream)
    def elapsed_ms(self): return float(libs.gpu.voxgpu_timer_elapsed_ms(self.h))


class Tree:
    """struct vox_node* (root) from vox_make_tree / vox_tree_import."""

    def __init__(self, handle, owner=True, reordered=None):
        self.h = handle
        self.owner = owner
        self.reordered = reordered

    @classmethod
    def build(cls, xyz):
        """vox_make_tree on a copy of xyz; the reordered copy is kept as .reordered"""
        d = dots4(xyz)
        h = libs.trees.vox_make_tree(d.ctypes.data, d.shape[0])
        if not h and d.shape[0]:
            raise VoxError(last_error())
        return cls(h, reordered=d)

    @classmethod
    def from_arrays(cls, a):
        """vox_tree_import from the canonical flat arrays (dict as oracle's flatten())."""
        keep = {k: np.ascontiguousarray(a[k]) for k in ("bb_min", "bb_max", "flags", "dots_num", "center",
                                                        "first_child", "child_mask", "dots_off", "dots")}
        arr = _TreeArrays(*[keep[k].ctypes.data for k, _ in _TreeArrays._fields_])
        h = libs.trees.vox_tree_import(C.byref(arr), keep["flags"].shape[0], keep["dots"].shape[0])
        if not h:
            raise VoxError(last_error())
        return cls(h)

    @classmethod
    def from_raw_file(cls, path, dim, samplesize, threshold=None, sample_range=None):
        """vox_read_raw_data_threshold (sample > threshold) or vox_read_raw_data_range (min <= sample < max)."""
        err = C.c_char_p()
        d = (C.c_uint * 3)(*dim)
        if sample_range is not None:
            h = libs.trees.vox_read_raw_data_range(path.encode(), C.byref(d), samplesize, sample_range[0], sample_range[1], C.byref(err))
        else:
            h = libs.trees.vox_read_raw_data_threshold(path.encode(), C.byref(d), samplesize, threshold, C.byref(err))
        if not h:
            raise VoxError((err.value or b"?").decode())
        return cls(h)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def destroy(self):
        if self.h and self.owner:
            libs.trees.vox_destroy_tree(self.h)
        self.h = None

    def info(self):
        ti = _TreeInfo()
        libs.trees.vox_tree_get_info(self.h, C.byref(ti))
        return ti

    def voxels(self):
        return int(libs.trees.vox_voxels_in_tree(self.h))

    def bounding_box(self):
        b = _Box()
        libs.trees.vox_bounding_box(self.h, C.byref(b))
        return np.array(b.min[:3], np.float32), np.array(b.max[:3], np.float32)

    def export(self):
        ti = self.info()
        n, nd = int(ti.n_nodes), int(ti.n_dots)
        a = {
            "bb_min": np.zeros((n, 3), np.float32), "bb_max": np.zeros((n, 3), np.float32),
            "flags": np.zeros(n, np.uint32), "dots_num": np.zeros(n, np.uint32),
            "center": np.zeros((n, 3), np.float32), "first_child": np.zeros(n, np.uint32),
            "child_mask": np.zeros(n, np.uint32), "dots_off": np.zeros(n, np.uint64),
            "dots": np.zeros((nd, 3), np.float32),
        }
        arr = _TreeArrays(*[a[k].ctypes.data for k, _ in _TreeArrays._fields_])
        libs.trees.vox_tree_export(self.h, C.byref(arr))
        a["depth"] = int(ti.depth)
        return a

    def node(self, node_id):
        return libs.trees.vox_tree_node(self.h, node_id)

    # ---- online edits (host; the device copy is refreshed by the next query) ----
    def edit(self, insert, xyz):
        """vox_insert_voxel (insert[i] true) / vox_delete_voxel per row of xyz -> their return values.
        The root pointer is passed by reference as in C; an emptied tree leaves self.h = None."""
        d = dots4(xyz)
        ins = np.broadcast_to(np.asarray(insert, bool), (d.shape[0],))
        res = np.zeros(d.shape[0], np.uint8)
        root = _vp(self.h)
        for i in range(d.shape[0]):
            fn = libs.trees.vox_insert_voxel if ins[i] else libs.trees.vox_delete_voxel
            res[i] = fn(C.byref(root), d[i].ctypes.data)
        self.h = root.value
        return res

    def insert(self, xyz):
        return self.edit(True, xyz)

    def delete(self, xyz):
        return self.edit(False, xyz)

    def rebuild(self):
        """vox_rebuild_tree: a new, independent tree."""
        h = libs.trees.vox_rebuild_tree(self.h)
        if not h and self.h:
            raise VoxError(last_error())
        return Tree(h)

    # ---- queries (GPU) ----
    def ray(self, origin, direction, start=None):
        """vox_ray_tree_intersection: (leaf_id, point) or (None, None)."""
        o, d, res = dots4([origin]), dots4([direction]), dots4([[0, 0, 0]])
        libs.trees.vox_last_error()
        leaf = libs.trees.vox_ray_tree_intersection(start or self.h, o.ctypes.data, d.ctypes.data, res.ctypes.data)
        if not leaf:
            err = libs.trees.vox_last_error().decode()
            if err and not gpu_available():
                raise VoxError(err)
            return None, None
        ti = _TreeInfo()
        libs.trees.vox_tree_get_info(leaf, C.byref(ti))
        return int(ti.node_id), res[0, :3].copy()

    def ball(self, center, radius):
        c = dots4([center])
        if not gpu_available():
            libs.trees.vox_tree_ball_collidep(self.h, c.ctypes.data, radius)
            raise VoxError(last_error())
        return int(libs.trees.vox_tree_ball_collidep(self.h, c.ctypes.data, radius))

    def rays(self, origins, dirs, start_node=0):
        """vox_ray_tree_intersection_batch with host buffers -> (leaf_ids, points, voxels)."""
        o, d = dots4(origins), dots4(dirs)
        n = o.shape[0]
        leaf = np.full(n, -1, np.int32)
        pts = dots4(np.zeros((n, 3), np.float32))
        vox = np.full(n, -1, np.int32)
        start = self.node(start_node) if start_node else self.h
        rc = libs.trees.vox_ray_tree_intersection_batch(start, o.ctypes.data, d.ctypes.data, n,
                                                        leaf.ctypes.data, pts.ctypes.data, vox.ctypes.data)
        if rc < 0:
            raise VoxError(last_error())
        return leaf, pts[:, :3].copy(), vox

    def balls(self, centers, radii):
        c = dots4(centers)
        r = np.ascontiguousarray(radii, np.float32)
        out = np.zeros(c.shape[0], np.uint8)
        rc = libs.trees.vox_tree_ball_collidep_batch(self.h, c.ctypes.data, r.ctypes.data, c.shape[0], out.ctypes.data)
        if rc < 0:
            raise VoxError(last_error())
        return out

    def device_handle(self):
        h = libs.trees.vox_tree_device_handle(self.h)
        if not h:
            raise VoxError(last_error())
        return h


class Camera:
    """A camera object driven through its vtable (struct vox_camera_interface)."""

    def __init__(self, name="simple-camera"):
        self.methods = libs.rnd.vox_camera_methods(name.encode())
        if not self.methods:
            raise VoxError(last_error())
        self.h = self.methods.contents.construct_camera(None)
        self.obj = C.cast(self.h, C.POINTER(_CameraObj))

    @property
    def iface(self):
        return self.obj.contents.iface.contents

    def __del__(self):
        try:
            if self.h:
                self.iface.destroy_camera(self.h)
                self.h = None
        except Exception:
            pass

    def _dot(self, fn, v, *pre):
        d = dots4([v])                       # named: the buffer must outlive the call
        fn(self.h, *pre, d.ctypes.data)

    def set_position(self, p): self._dot(self.iface.set_property_dot, p, b"position")
    def set_rotation(self, a): self._dot(self.iface.set_property_dot, a, b"rotation")
    def set_fov(self, fov): self.iface.set_property_number(self.h, b"fov", fov)
    def look_at(self, p): self._dot(self.iface.look_at, p)
    def rotate(self, d): self._dot(self.iface.rotate_camera, d)
    def move(self, d): self._dot(self.iface.move_camera, d)
    def set_window_size(self, w, h): self.iface.set_window_size(self.h, w, h)

    def get_position(self):
        a = dots4([[0, 0, 0]])
        self.iface.get_position(self.h, a.ctypes.data)
        return a[0, :3].copy()

    def screen2world(self, sx, sy):
        a = dots4([[0, 0, 0]])
        self.iface.screen2world(self.h, a.ctypes.data, sx, sy)
        return a[0, :3].copy()

    def describe(self):
        d = _CameraDesc()
        if not self.iface.describe(self.h, C.byref(d)):
            raise VoxError("camera cannot be described")
        return d

    def gpu_camera(self):
        d = self.describe()
        g = GpuCamera()
        g.kind = d.kind
        for i in range(3):
            g.position[i] = d.position[i]
        for i in range(4):
            g.rotation[i] = d.rotation[i]
        g.mul, g.xsub, g.ysub = d.mul, d.xsub, d.ysub
        g.sinphi, g.cosphi, g.k = d.sinphi, d.cosphi, d.k
        return g

    def state(self):
        """POD state as 11 floats: pos[3], quat[4] | (sinphi, cosphi, k, 0), mul, xsub, ysub, 0."""
        d = self.describe()
        s = np.zeros(11, np.float32)
        s[0:3] = d.position[:]
        if d.kind == 1:
            s[3], s[4], s[5] = d.sinphi, d.cosphi, d.k
        else:
            s[3:7] = d.rotation[:]
        s[7], s[8], s[9] = d.mul, d.xsub, d.ysub
        return s


class LightManager:
    """struct vox_light_manager* (include/voxrnd/lights.h)."""

    def __init__(self):
        self.h = libs.rnd.vox_make_light_manager()

    def __del__(self):
        try:
            if self.h:
                libs.rnd.vox_destroy_light_manager(self.h)
                self.h = None
        except Exception:
            pass

    def insert(self, center, radius, color):
        c, col = dots4([center]), dots4([color])      # named: the buffers must outlive the call
        return int(libs.rnd.vox_insert_shadowless_light(self.h, c.ctypes.data, radius, col.ctypes.data))

    def delete(self, center, radius):
        c = dots4([center])
        return int(libs.rnd.vox_delete_shadowless_light(self.h, c.ctypes.data, radius))

    def clear(self):
        libs.rnd.vox_delete_shadowless_lights(self.h)

    def count(self):
        return int(libs.rnd.vox_shadowless_lights_number(self.h))

    def set_ambient(self, color):
        c = dots4([color])
        return int(libs.rnd.vox_set_ambient_light(self.h, c.ctypes.data))

    def get_light(self, points):
        """vox_get_light on the HOST for every row of points -> (n, 3)."""
        p = dots4(points)
        out = dots4(np.zeros((p.shape[0], 3), np.float32))
        for i in range(p.shape[0]):
            libs.rnd.vox_get_light(self.h, p[i].ctypes.data, out[i].ctypes.data)
        return out[:, :3].copy()

    def flatten(self):
        """The flat light set vox_render uploads (dict of arrays + 'ambient')."""
        nn, ns = C.c_uint32(0), C.c_uint32(0)
        libs.rnd.vox_light_manager_flatten(self.h, C.byref(nn), C.byref(ns), None, None, None, None)
        a = {"node_bounds": np.zeros((nn.value, 4), np.float32), "node_link": np.zeros((nn.value, 4), np.uint32),
             "sphere_geom": np.zeros((ns.value, 4), np.float32), "sphere_light": np.zeros((ns.value, 4), np.float32)}
        libs.rnd.vox_light_manager_flatten(self.h, None, None, a["node_bounds"].ctypes.data, a["node_link"].ctypes.data,
                                           a["sphere_geom"].ctypes.data, a["sphere_light"].ctypes.data)
        a["ambient"] = self.get_light([[1e30, 1e30, 1e30]])[0]     # far from every light: the ambient colour
        return a


def upload_lights(flat):
    """voxgpu_lights_upload of a flat light set (dict as LightManager.flatten()) -> device handle."""
    keep = {k: np.ascontiguousarray(flat[k]) for k in ("node_bounds", "node_link", "sphere_geom", "sphere_light")}
    f = GpuLightsFlat((C.c_float * 3)(*[float(x) for x in flat["ambient"]]), keep["node_bounds"].shape[0],
                      keep["sphere_geom"].shape[0], keep["node_bounds"].ctypes.data, keep["node_link"].ctypes.data,
                      keep["sphere_geom"].ctypes.data, keep["sphere_light"].ctypes.data)
    h = libs.gpu.voxgpu_lights_upload(C.byref(f))
    if not h:
        raise VoxError(libs.gpu.voxgpu_last_error().decode())
    return h


class Context:
    """struct vox_rnd_ctx* with a headless surface."""

    def __init__(self, w, h):
        self.h = libs.rnd.vox_make_context_and_window(w, h)
        if not self.h:
            raise VoxError(f"bad geometry {w}x{h} (renderer.c:97-100) or no memory")
        self.w, self.hgt = w, h
        self._keep = []

    def __del__(self):
        try:
            if self.h:
                libs.rnd.vox_destroy_context(self.h)
                self.h = None
        except Exception:
            pass

    def set_scene(self, tree):
        self._scene = tree
        libs.rnd.vox_context_set_scene(self.h, tree.h if tree is not None else None)

    def set_camera(self, cam):
        self._camera = cam
        libs.rnd.vox_context_set_camera(self.h, cam.h)

    def set_light_manager(self, lm):
        self._lights = lm
        libs.rnd.vox_context_set_light_manager(self.h, lm.h if lm is not None else None)

    def set_quality(self, q):
        return bool(libs.rnd.vox_context_set_quality(self.h, q))

    def set_devices(self, n):
        """vox_context_set_devices: split every frame over up to n devices of this process (0 = all)."""
        return int(libs.rnd.vox_context_set_devices(self.h, n))

    def set_texture(self, tex):
        tex = np.ascontiguousarray(tex, np.uint8)
        assert tex.size == 32 * 32 * 32
        libs.rnd.vox_context_set_texture(self.h, tex.ctypes.data)

    def texture(self):
        p = libs.rnd.vox_context_get_texture(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(32 * 32 * 32,)).copy()

    def render(self):
        libs.rnd.vox_render(self.h)

    def redraw(self):
        libs.rnd.vox_redraw(self.h)

    def kernel_ms(self):
        return float(libs.rnd.vox_context_last_kernel_ms(self.h))

    def pixels(self):
        """View of surface->pixels (h, w) uint32 -- valid after redraw()."""
        s = self.h.contents.surface.contents
        buf = (C.c_uint32 * (s.w * s.h)).from_address(s.pixels)
        return np.frombuffer(buf, np.uint32).reshape(s.h, s.w)

    def pixfmt(self):
        f = self.h.contents.surface.contents.format.contents
        return GpuPixFmt(f.Amask, f.Rloss, f.Gloss, f.Bloss, 0, f.Rshift, f.Gshift, f.Bshift, 0)
