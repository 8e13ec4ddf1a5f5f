#!/usr/bin/env python
"""-DCTB_PIXEL_STAMPS build (CUTRACE_B200_LIB=cutrace_b200/lib/variants/libcutrace_b200_stamps.so): in-kernel %globaltimer stamps of triangle.json."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import cutrace_b200 as ct
s, _ = bench.load_workload("triangle")
import ctypes, numpy as np
lib = ct._lib.load()
print("library", ct._lib.LIB_PATH, flush=True)
lib.cutrace_debug_pixel_stamps.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
names = ["warm", "hit0", "phong0", "path", "loop-exit", "stats-out"]
with ct.Renderer(s) as r:
    for _ in range(12):
        ms = r.render()["render_ms"]
        st = np.zeros(7, dtype=np.uint64)
        lib.cutrace_debug_pixel_stamps(r._ctx, st.ctypes.data)
        d = (st[1:].astype(np.int64) - np.int64(st[0]))
        print("render", round(ms * 1e3, 2), "us; ns after kernel entry:", "  ".join(f"{n} {int(v)}" for n, v in zip(names, d)), flush=True)
