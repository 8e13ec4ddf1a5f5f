"""ctypes bindings for the parity checkers in oracle/.  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module.  The product package ``cutrace_b200`` never does.

  oracle_render(...)    plain-C restatement           oracle/libcutrace_oracle.so
  ref_host_render(...)  reference headers on the host oracle/_ref/libcutrace_ref_host.so
  ref_gpu_render(...)   reference kernel for sm_100a  oracle/_ref/libcutrace_ref_gpu.so
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libcutrace_oracle.so")
REF_HOST_SO = os.path.join(HERE, "_ref", "libcutrace_ref_host.so")
REF_GPU_SO = os.path.join(HERE, "_ref", "libcutrace_ref_gpu.so")

_libs = {}


def build(quiet=True):
    """Runs oracle/Makefile (the reference parts only when /root/reference or $CUTRACE_REF exists)."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _load(path):
    if path not in _libs:
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        _libs[path] = C.CDLL(path)
    return _libs[path]


def have_oracle():
    return os.path.exists(ORACLE_SO)


def have_ref_host():
    return os.path.exists(REF_HOST_SO)


def have_ref_gpu():
    return os.path.exists(REF_GPU_SO)


def _outputs(n):
    return (np.empty(n, np.float32), np.empty((n, 3), np.float32), np.empty((n, 3), np.float32),
            np.empty(n, np.uint32))


def _px_arg(scene, px):
    if px is None:
        return scene.width * scene.height, None, None
    px = np.ascontiguousarray(np.asarray(px, dtype=np.uint64))
    return len(px), px, px.ctypes.data_as(C.c_void_p)


def oracle_render(scene, fudge=1e-3, bounces=5, px=None, threads=None):
    """Returns dict(depth, normal, color, hit_id, counters) from the C restatement."""
    lib = _load(ORACLE_SO)
    fn = lib.cutrace_oracle_render
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.c_void_p, C.c_void_p, C.c_int]
    n, keep, pxp = _px_arg(scene, px)
    depth, normal, color, hid = _outputs(n)
    counters = np.zeros(6, np.uint64)
    desc = scene.as_desc()
    rc = fn(C.byref(desc), fudge, bounces, n, pxp, depth.ctypes.data, normal.ctypes.data, color.ctypes.data,
            hid.ctypes.data, counters.ctypes.data, threads or os.cpu_count() or 1)
    if rc:
        raise RuntimeError(f"cutrace_oracle_render failed: {rc}")
    names = ("casts", "rays_primary", "rays_reflect", "rays_transmit", "rays_shadow", "shadow_casts")
    return dict(depth=depth, normal=normal, color=color, hit_id=hid,
                counters={k: int(v) for k, v in zip(names, counters)})


def ref_host_render(scene, fudge=1e-3, px=None, threads=None):
    lib = _load(REF_HOST_SO)
    fn = lib.cutrace_ref_host_render
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.c_void_p, C.c_int]
    n, keep, pxp = _px_arg(scene, px)
    depth, normal, color, hid = _outputs(n)
    desc = scene.as_desc()
    rc = fn(C.byref(desc), fudge, 5, n, pxp, depth.ctypes.data, normal.ctypes.data, color.ctypes.data,
            hid.ctypes.data, threads or os.cpu_count() or 1)
    if rc:
        raise RuntimeError(f"cutrace_ref_host_render failed: {rc}")
    return dict(depth=depth, normal=normal, color=color, hit_id=hid)


def ref_gpu_render(scene, fudge=1e-3, px=None, iters=1, warmup=0):
    """Reference render_kernel<S,5> rebuilt for sm_100a (inc/kernel.hpp:35-60), launched
    <<<w*h/256+1,256>>> like inc/kernel.hpp:103-106.  px != None renders only those pixels with an
    oracle-side kernel that calls the reference's ray_cast/ray_color (config-5 subset parity).
    Returns outputs + ``render_ms`` (CUDA events, mean over ``iters``) and ``total_ms``."""
    lib = _load(REF_GPU_SO)
    fn = lib.cutrace_ref_gpu_render
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_float, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.c_int, C.c_int, C.c_void_p]
    n, keep, pxp = _px_arg(scene, px)
    depth, normal, color, hid = _outputs(n)
    ms = np.zeros(2, np.float32)
    desc = scene.as_desc()
    rc = fn(C.byref(desc), fudge, n, pxp, depth.ctypes.data, normal.ctypes.data, color.ctypes.data,
            hid.ctypes.data, iters, warmup, ms.ctypes.data)
    if rc:
        raise RuntimeError(f"cutrace_ref_gpu_render failed: {rc}")
    return dict(depth=depth, normal=normal, color=color, hit_id=hid, render_ms=float(ms[0]), total_ms=float(ms[1]))


def encode_bytes(depth, normal, color, max_d):
    lib = _load(ORACLE_SO)
    fn = lib.cutrace_oracle_encode_bytes
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    n = depth.size
    d8, n8, c8 = (np.empty((n, 3), np.uint8) for _ in range(3))
    depth = np.ascontiguousarray(depth, np.float32)
    normal = np.ascontiguousarray(normal, np.float32)
    color = np.ascontiguousarray(color, np.float32)
    fn(depth.ctypes.data, normal.ctypes.data, color.ctypes.data, max_d, n, d8.ctypes.data, n8.ctypes.data, c8.ctypes.data)
    return d8, n8, c8


def max_depth(depth):
    lib = _load(ORACLE_SO)
    fn = lib.cutrace_oracle_max_depth
    fn.restype = C.c_float
    fn.argtypes = [C.c_void_p, C.c_uint64]
    depth = np.ascontiguousarray(depth, np.float32)
    return float(fn(depth.ctypes.data, depth.size))
