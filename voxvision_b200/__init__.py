"""voxvision_b200 -- B200-native hot path of voxvision behind its own C API.

The product is the set of shared libraries under ``voxvision_b200/lib``
(built in-tree by ``make -C voxvision_b200`` / ``__graft_entry__.build()``):

    libvoxgpu.so     CUDA kernels (sm_100a) + thin C-ABI     include/voxgpu.h
    libvoxtrees.so   vox_make_tree, vox_ray_tree_intersection,
                     vox_tree_ball_collidep (+ batched forms)  include/voxtrees.h
    libvoxrnd.so     cameras, contexts, vox_render/vox_redraw  include/voxrnd.h
    simple-camera.so, doom-camera.so                           camera modules

This Python package is only a ctypes binding over that C API, used by the
tests and by bench.py; it contains no compute of its own and never touches
``oracle/``.  Everything that searches a tree runs on the GPU: without a CUDA
device those calls fail loudly (VoxError) -- there is no CPU fallback.
"""
from .api import (  # noqa: F401
    LightManager, upload_lights,
    VoxError, libs, gpu_available, device_count, set_device, last_error, launch_count,
    set_voxel_size, get_voxel_size, Tree, Camera, Context, DeviceBuffer, Timer,
    QUALITY_BEST, QUALITY_FAST, QUALITY_ADAPTIVE, QUALITY_RAY_MERGE, QUALITY_RAY_MERGE_ACCURATE,
    LIB_DIR, INCLUDE_DIR,
)
