"""ctypes binding of the C-ABI in include/cutrace.h (cutrace_b200/lib/libcutrace_b200.so).

The library is CUDA-only: if the shared object is missing or no CUDA device is present every call
raises — there is no CPU fallback in the product path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CUTRACE_B200_LIB") or os.path.join(HERE, "lib", "libcutrace_b200.so")

# every symbol include/cutrace.h declares
SYMBOLS = (
    "cutrace_default_opts", "cutrace_upload_scene", "cutrace_render", "cutrace_download", "cutrace_render_download", "cutrace_download_bytes", "cutrace_free",
    "cutrace_last_error", "cutrace_set_camera", "cutrace_get_stats", "cutrace_get_phase_ms", "cutrace_device_buffers", "cutrace_frame_device",
    "cutrace_frame_ipc_export", "cutrace_frame_ipc_import", "cutrace_frame_attach", "cutrace_enable_peer_access",
    "cutrace_set_frame_max_depth",
    "cutrace_untile_device", "cutrace_encode_bytes_device", "cutrace_host_alloc", "cutrace_host_free", "cutrace_host_register", "cutrace_host_unregister", "cutrace_trim_memory",
    "cutrace_validate_bvh", "cutrace_debug_radix_sort", "cutrace_debug_phong_pow", "cutrace_abi_version", "cutrace_tile_size",
    "cutrace_debug_tile_of_slot", "cutrace_debug_slot_of_tile", "cutrace_debug_segment_length", "cutrace_debug_segment_work",
)

FLAG_NO_SMEM_TOP, FLAG_VALIDATE_BVH, FLAG_BRUTE_FORCE, FLAG_SERIALIZE, FLAG_FRAME_KERNEL, FLAG_LAUNCHES, \
    FLAG_PIXEL_KERNEL, FLAG_FAST_BUILD = 1, 2, 4, 8, 16, 32, 64, 128
IPC_HANDLE_BYTES = 80


class cutrace_opts(C.Structure):
    _fields_ = [
        ("fudge", C.c_float), ("bounces", C.c_uint32), ("device", C.c_int32), ("flags", C.c_uint32),
        ("tile_rank", C.c_uint32), ("tile_world", C.c_uint32), ("stream", C.c_void_p), ("leaf_size", C.c_uint32),
        ("reserved", C.c_uint32 * 7),
    ]


class cutrace_stats(C.Structure):
    _fields_ = [
        ("build_ms", C.c_float), ("render_ms", C.c_float), ("gather_ms", C.c_float), ("max_depth", C.c_float),
        ("rays_primary", C.c_uint64), ("rays_reflect", C.c_uint64), ("rays_transmit", C.c_uint64),
        ("rays_shadow", C.c_uint64), ("shadow_casts", C.c_uint64), ("local_pixels", C.c_uint64),
        ("kernel_launches", C.c_uint32), ("bvh_nodes", C.c_uint32), ("bvh_depth", C.c_uint32), ("smem_nodes", C.c_uint32),
        ("trace_ms", C.c_float), ("shade_ms", C.c_float), ("scheduler", C.c_uint32), ("reserved", C.c_uint32 * 5),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}
        d["rays_total"] = self.rays_primary + self.rays_reflect + self.rays_transmit + self.rays_shadow
        return d


class CutraceError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cutrace error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Loads the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not found: build the CUDA extension first (`make` or __graft_entry__.build()); "
            "cutrace_b200 has no CPU path")
    lib = C.CDLL(LIB_PATH)
    P = C.c_void_p
    lib.cutrace_default_opts.argtypes = [C.POINTER(cutrace_opts)]
    lib.cutrace_default_opts.restype = None
    lib.cutrace_upload_scene.argtypes = [P, C.POINTER(cutrace_opts), C.POINTER(P)]
    lib.cutrace_render.argtypes = [P, C.POINTER(cutrace_stats)]
    lib.cutrace_download.argtypes = [P, P, P, P, P, C.POINTER(C.c_float)]
    lib.cutrace_render_download.argtypes = [P, P, P, P, P, C.POINTER(C.c_float), C.POINTER(cutrace_stats)]
    lib.cutrace_download_bytes.argtypes = [P, P, P, P, C.POINTER(C.c_float)]
    lib.cutrace_free.argtypes = [P]
    lib.cutrace_free.restype = None
    lib.cutrace_last_error.restype = C.c_char_p
    lib.cutrace_set_camera.argtypes = [P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                       C.POINTER(C.c_float), C.c_float, C.c_uint32, C.c_uint32]
    lib.cutrace_get_stats.argtypes = [P, C.POINTER(cutrace_stats)]
    lib.cutrace_get_phase_ms.argtypes = [P, C.POINTER(C.c_float), C.c_uint32, C.POINTER(C.c_uint32)]
    lib.cutrace_device_buffers.argtypes = [P, C.POINTER(P), C.POINTER(P), C.POINTER(P), C.POINTER(P), C.POINTER(C.c_uint64)]
    lib.cutrace_frame_device.argtypes = [P, C.POINTER(P), C.POINTER(P), C.POINTER(P), C.POINTER(P)]
    lib.cutrace_frame_ipc_export.argtypes = [P, P]
    lib.cutrace_frame_ipc_import.argtypes = [P, P]
    lib.cutrace_frame_attach.argtypes = [P, P, C.c_uint32, C.c_uint32]
    lib.cutrace_host_register.argtypes = [P, C.c_size_t]
    lib.cutrace_host_unregister.argtypes = [P]
    lib.cutrace_trim_memory.argtypes = [C.c_int]
    lib.cutrace_enable_peer_access.argtypes = [C.c_int, C.c_int]
    lib.cutrace_set_frame_max_depth.argtypes = [P, C.c_float]
    lib.cutrace_untile_device.argtypes = [P, C.c_uint32, P, P, P, P, C.c_uint64, P, P, P, P]
    lib.cutrace_encode_bytes_device.argtypes = [P, P, P, P, C.c_float, C.c_uint64, P, P, P]
    lib.cutrace_host_alloc.argtypes = [C.c_size_t]
    lib.cutrace_host_alloc.restype = P
    lib.cutrace_host_free.argtypes = [P]
    lib.cutrace_host_free.restype = None
    lib.cutrace_validate_bvh.argtypes = [P]
    lib.cutrace_debug_radix_sort.argtypes = [P, P, C.c_uint32, C.c_int]
    lib.cutrace_debug_phong_pow.argtypes = [P, P, P, P, C.c_uint32, C.c_int]
    lib.cutrace_abi_version.restype = C.c_uint32
    lib.cutrace_tile_size.restype = C.c_uint32
    lib.cutrace_debug_tile_of_slot.argtypes = [C.c_uint32] * 5 + [P, P]
    lib.cutrace_debug_slot_of_tile.argtypes = [C.c_uint32] * 6
    lib.cutrace_debug_slot_of_tile.restype = C.c_uint32
    for fn in (lib.cutrace_debug_segment_length, lib.cutrace_debug_segment_work):
        fn.argtypes = [C.c_uint32] * 3
        fn.restype = C.c_uint32
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise CutraceError(rc, load().cutrace_last_error().decode("utf-8", "replace"))
