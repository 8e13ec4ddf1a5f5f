"""ctypes binding of the C++ host front-end (include/cutrace_host.h, cutrace_b200/lib/libcutrace_host.so):
the scene JSON/STL loader and JPEG writer the `cutrace` CLI uses.  No CUDA involved."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .scene import FlatScene, SceneError, cutrace_scene_desc

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(HERE, "lib", "libcutrace_host.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise FileNotFoundError(f"{HOST_LIB_PATH} not found: run `make host`")
        lib = C.CDLL(HOST_LIB_PATH)
        lib.cutrace_host_load_scene.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(cutrace_scene_desc),
                                                C.c_char_p, C.c_size_t]
        lib.cutrace_host_free_scene.argtypes = [C.c_void_p]
        lib.cutrace_host_free_scene.restype = None
        lib.cutrace_host_write_jpeg.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        lib.cutrace_host_look_at.argtypes = [C.POINTER(C.c_float)] * 6
        lib.cutrace_host_look_at.restype = None
        _lib = lib
    return _lib


def _arr(ptr, n, dtype, k=0):
    if not ptr or n == 0:
        return np.zeros((0, k) if k else 0, dtype)
    ct = C.c_float if dtype == np.float32 else C.c_uint32
    a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n * (k or 1),)).copy()
    return a.reshape(-1, k) if k else a


def load_scene(path, base_dir="", accept_aliases=False) -> FlatScene:
    """C++ loader (`default_schema::load_file` equivalent) -> FlatScene (arrays copied out)."""
    lib = load()
    h = C.c_void_p()
    d = cutrace_scene_desc()
    err = C.create_string_buffer(8192)
    rc = lib.cutrace_host_load_scene(os.fsencode(path), os.fsencode(base_dir or ""), int(accept_aliases), C.byref(h), C.byref(d), err, len(err))
    if rc:
        raise SceneError(err.value.decode("utf-8", "replace").strip() or f"load failed ({rc})")
    try:
        f32, u32 = np.float32, np.uint32
        return FlatScene(
            cam_pos=np.array(d.cam_pos[:], f32), cam_up=np.array(d.cam_up[:], f32), cam_forward=np.array(d.cam_forward[:], f32),
            cam_right=np.array(d.cam_right[:], f32), ambient=d.ambient, width=d.width, height=d.height,
            tri_p1=_arr(d.tri_p1, d.n_triangles, f32, 3), tri_p2=_arr(d.tri_p2, d.n_triangles, f32, 3),
            tri_p3=_arr(d.tri_p3, d.n_triangles, f32, 3), tri_object=_arr(d.tri_object, d.n_triangles, u32),
            sph_center=_arr(d.sph_center, d.n_spheres, f32, 3), sph_radius=_arr(d.sph_radius, d.n_spheres, f32),
            sph_object=_arr(d.sph_object, d.n_spheres, u32),
            pl_point=_arr(d.pl_point, d.n_planes, f32, 3), pl_normal=_arr(d.pl_normal, d.n_planes, f32, 3),
            pl_object=_arr(d.pl_object, d.n_planes, u32),
            obj_material=_arr(d.obj_material, d.n_objects, u32), obj_kind=_arr(d.obj_kind, d.n_objects, u32),
            mat_color=_arr(d.mat_color, d.n_materials, f32, 3), mat_specular=_arr(d.mat_specular, d.n_materials, f32),
            mat_reflect=_arr(d.mat_reflect, d.n_materials, f32), mat_phong=_arr(d.mat_phong, d.n_materials, f32),
            mat_transparency=_arr(d.mat_transparency, d.n_materials, f32),
            light_kind=_arr(d.light_kind, d.n_lights, u32), light_vec=_arr(d.light_vec, d.n_lights, f32, 3),
            light_color=_arr(d.light_color, d.n_lights, f32, 3),
        )
    finally:
        lib.cutrace_host_free_scene(h)


def write_jpeg(path, rgb, quality=90):
    """rgb: (h, w, 3) uint8"""
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    if load().cutrace_host_write_jpeg(os.fsencode(path), w, h, rgb.ctypes.data, quality):
        raise OSError(f"cannot write {path}")
