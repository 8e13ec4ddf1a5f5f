#!/usr/bin/env python
"""Times the phases of the end-to-end call sequence (upload / render / download / free) per frame."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cutrace_b200 as ct
import bench
scene, wl = bench.load_workload(sys.argv[1] if len(sys.argv) > 1 else "bunny4k")
lib = ct._lib.load()
n = scene.width * scene.height
bufs = {}
for k, m in (("depth", 1), ("normal", 3), ("color", 3)):
    p = lib.cutrace_host_alloc(n * m * 4)
    bufs[k] = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n * m,))
for i in range(6):
    t0 = time.perf_counter(); r = ct.Renderer(scene)
    t1 = time.perf_counter(); md = C.c_float(); cst = ct._lib.cutrace_stats()
    ct._lib.check(lib.cutrace_render_download(r._ctx, bufs["depth"].ctypes.data, bufs["normal"].ctypes.data, bufs["color"].ctypes.data, None, C.byref(md), C.byref(cst)))
    st = cst.as_dict(); t2 = t1
    t3 = time.perf_counter(); r.close()
    t4 = time.perf_counter()
    print(f"frame {i}: upload {1e3*(t1-t0):7.2f}  render+download {1e3*(t3-t1):7.2f} (device {st['render_ms']:.2f}, build {st['build_ms']:.2f})  free {1e3*(t4-t3):7.2f}  total {1e3*(t4-t0):7.2f} ms")
