"""cutrace_b200 — B200-native render path for cutrace scenes.

Host-side mirror of the reference's render operator (``cutrace::gpu::render``,
/root/reference/inc/kernel.hpp:86-130) on top of the C-ABI in include/cutrace.h:

    r = Renderer(scene)            # cutrace_upload_scene  (default_to_gpu, inc/default_schema.hpp:935)
    stats = r.render()             # cutrace_render        (launch + sync, inc/kernel.hpp:103-108)
    out = r.download()             # cutrace_download      (D2H + max-depth, inc/kernel.hpp:110-125)

or ``render(scene, fudge=1e-3, bounces=5)`` which does all three and returns the same outputs as the
reference operator (depth_map, color_map, normal_map, max, render_ms, total_ms).
CUDA only — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import _lib
from ._lib import (CutraceError, FLAG_BRUTE_FORCE, FLAG_FAST_BUILD, FLAG_FRAME_KERNEL, FLAG_LAUNCHES, FLAG_NO_SMEM_TOP, FLAG_PIXEL_KERNEL, FLAG_SERIALIZE, FLAG_VALIDATE_BVH,
                   cutrace_opts, cutrace_stats)
from .scene import FlatScene, SceneError, load_scene_json, look_at

__all__ = ["Renderer", "render", "FlatScene", "SceneError", "CutraceError", "load_scene_json", "look_at",
           "FLAG_BRUTE_FORCE", "FLAG_NO_SMEM_TOP", "FLAG_VALIDATE_BVH", "FLAG_SERIALIZE", "FLAG_FRAME_KERNEL", "FLAG_LAUNCHES", "FLAG_PIXEL_KERNEL", "FLAG_FAST_BUILD"]


class Renderer:
    """One uploaded scene on one CUDA device (a ``cutrace_ctx``)."""

    def __init__(self, scene: FlatScene, fudge=1e-3, bounces=5, device=-1, flags=0, tile_rank=0, tile_world=1,
                 stream=None, leaf_size=0):
        self._lib = _lib.load()
        self.scene = scene
        o = cutrace_opts()
        self._lib.cutrace_default_opts(C.byref(o))
        o.fudge, o.bounces, o.device, o.flags = fudge, bounces, device, flags
        o.tile_rank, o.tile_world, o.stream, o.leaf_size = tile_rank, tile_world, stream, leaf_size
        self.opts = o
        self._ctx = C.c_void_p()
        desc = scene.as_desc()
        _lib.check(self._lib.cutrace_upload_scene(C.byref(desc), C.byref(o), C.byref(self._ctx)))
        self.width, self.height = scene.width, scene.height

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.cutrace_free(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- the three calls ----------------------------------------------------------------------------
    def render(self):
        st = cutrace_stats()
        _lib.check(self._lib.cutrace_render(self._ctx, C.byref(st)))
        return st.as_dict()

    def stats(self):
        st = cutrace_stats()
        _lib.check(self._lib.cutrace_get_stats(self._ctx, C.byref(st)))
        return st.as_dict()

    def phase_ms(self):
        """cutrace_get_phase_ms: when each bounce level's rays were all traced / the frame was assembled (ms from kernel start)."""
        buf = (C.c_float * 18)()
        n = C.c_uint32()
        _lib.check(self._lib.cutrace_get_phase_ms(self._ctx, buf, 18, C.byref(n)))
        return [float(buf[i]) for i in range(n.value)]

    def download(self, into=None, want=("depth", "normal", "color", "hit_id")):
        n = self.width * self.height
        out = into or {}
        shapes = {"depth": ((n,), np.float32), "normal": ((n, 3), np.float32), "color": ((n, 3), np.float32),
                  "hit_id": ((n,), np.uint32)}
        ptr = {}
        for k, (shape, dt) in shapes.items():
            if k in want:
                if k not in out:
                    out[k] = np.empty(shape, dt)
                ptr[k] = out[k].ctypes.data
            else:
                ptr[k] = None
        md = C.c_float()
        _lib.check(self._lib.cutrace_download(self._ctx, ptr["depth"], ptr["normal"], ptr["color"], ptr["hit_id"], C.byref(md)))
        out["max_depth"] = md.value
        return out

    def render_download(self, into=None):
        """cutrace_render_download: render and fill host images in one call (G-buffer copies overlap the bounce levels)."""
        n = self.width * self.height
        out = into or {}
        shapes = {"depth": ((n,), np.float32), "normal": ((n, 3), np.float32), "color": ((n, 3), np.float32), "hit_id": ((n,), np.uint32)}
        for k, (shape, dt) in shapes.items():
            if k not in out:
                out[k] = np.empty(shape, dt)
        md, st = C.c_float(), cutrace_stats()
        _lib.check(self._lib.cutrace_render_download(self._ctx, out["depth"].ctypes.data, out["normal"].ctypes.data, out["color"].ctypes.data,
                                                     out["hit_id"].ctypes.data, C.byref(md), C.byref(st)))
        out["max_depth"] = md.value
        return out, st.as_dict()

    def download_bytes(self):
        """The three 8-bit RGB images of the output stage (inc/images.hpp:26-88), encoded on the device."""
        n = self.width * self.height
        d8, n8, c8 = (np.empty((n, 3), np.uint8) for _ in range(3))
        md = C.c_float()
        _lib.check(self._lib.cutrace_download_bytes(self._ctx, d8.ctypes.data, n8.ctypes.data, c8.ctypes.data, C.byref(md)))
        return dict(depth_rgb=d8, normal_rgb=n8, color_rgb=c8, max_depth=md.value)

    # -- helpers --------------------------------------------------------------------------------------
    def set_camera(self, pos, up, forward, right, ambient, width, height):
        arr = [(C.c_float * 3)(*[float(x) for x in v]) for v in (pos, up, forward, right)]
        _lib.check(self._lib.cutrace_set_camera(self._ctx, arr[0], arr[1], arr[2], arr[3], ambient, width, height))
        self.width, self.height = width, height

    def set_resolution(self, width, height):
        s = self.scene
        self.set_camera(s.cam_pos, s.cam_up, s.cam_forward, s.cam_right, s.ambient, width, height)

    def validate_bvh(self):
        _lib.check(self._lib.cutrace_validate_bvh(self._ctx))

    def device_buffers(self):
        """(depth, normal, color, hit_id) device pointers (ints) of the local tile-major buffers + padded pixel count."""
        p = [C.c_void_p() for _ in range(4)]
        n = C.c_uint64()
        _lib.check(self._lib.cutrace_device_buffers(self._ctx, *[C.byref(x) for x in p], C.byref(n)))
        return [x.value for x in p], n.value

    def frame_device(self):
        """(depth, normal, color, hit_id) device pointers of this ctx's own row-major frame."""
        p = [C.c_void_p() for _ in range(4)]
        _lib.check(self._lib.cutrace_frame_device(self._ctx, *[C.byref(x) for x in p]))
        return [x.value for x in p]

    def frame_ipc_export(self) -> bytes:
        buf = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        _lib.check(self._lib.cutrace_frame_ipc_export(self._ctx, buf))
        return buf.raw

    def frame_ipc_import(self, handle: bytes):
        _lib.check(self._lib.cutrace_frame_ipc_import(self._ctx, C.create_string_buffer(handle, _lib.IPC_HANDLE_BYTES)))

    def frame_attach(self, block_ptr, width=None, height=None):
        """Kernels of this ctx store their tiles into the row-major frame block at ``block_ptr`` (another ctx's frame, or
        pinned / registered host memory); ``None`` detaches.  width/height describe the block (default: this ctx's)."""
        _lib.check(self._lib.cutrace_frame_attach(self._ctx, block_ptr, self.width if width is None else width,
                                                  self.height if height is None else height))

    def untile_device(self, world, g_depth, g_normal, g_color, g_id, stride_px, depth, normal, color, hit_id):
        _lib.check(self._lib.cutrace_untile_device(self._ctx, world, g_depth, g_normal, g_color, g_id, stride_px,
                                                   depth, normal, color, hit_id))

    def encode_bytes_device(self, depth, normal, color, max_depth, n_px, depth_rgb, normal_rgb, color_rgb):
        _lib.check(self._lib.cutrace_encode_bytes_device(self._ctx, depth, normal, color, max_depth, n_px,
                                                         depth_rgb, normal_rgb, color_rgb))


def render(scene: FlatScene, fudge=1e-3, bounces=5, **kw):
    """``cutrace::gpu::render`` equivalent (inc/kernel.hpp:86-130): returns a dict with depth_map (h,w),
    color_map (h,w,3), normal_map (h,w,3), hit_id (h,w), max, render_ms, total_ms and the ray statistics."""
    t0 = time.perf_counter()
    with Renderer(scene, fudge=fudge, bounces=bounces, **kw) as r:
        st = r.render()
        out = r.download()
    total_ms = (time.perf_counter() - t0) * 1e3
    h, w = scene.height, scene.width
    return dict(depth_map=out["depth"].reshape(h, w), color_map=out["color"].reshape(h, w, 3),
                normal_map=out["normal"].reshape(h, w, 3), hit_id=out["hit_id"].reshape(h, w), max=out["max_depth"],
                render_ms=st["render_ms"], total_ms=total_ms, stats=st)
