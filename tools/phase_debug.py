#!/usr/bin/env python
"""Per-phase timeline of the persistent frame kernel from a -DCTB_PHASE_DEBUG build (tools/build_variant.sh dbg -DCTB_PHASE_DEBUG):
for every phase, when the first / last warp ran out of trace work, when the last warp had arrived, when the first / last warp saw
the phase open.  usage: CUTRACE_B200_LIB=cutrace_b200/lib/variants/libcutrace_b200_dbg.so tools/phase_debug.py <workload> <world>"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import cutrace_b200 as ct  # noqa: E402

wl, world = sys.argv[1], int(sys.argv[2])
scene, _ = bench.load_workload(wl)
lib = ct._lib.load()
buf = np.zeros((6, 18), np.uint64)
with ct.Renderer(scene, tile_rank=0, tile_world=world) as r:
    for _ in range(3):
        r.render()
    lib.cutrace_debug_phase_dump(buf.ctypes.data_as(C.c_void_p))
    st = r.render()
    lib.cutrace_debug_phase_dump(buf.ctypes.data_as(C.c_void_p))
    ph = r.phase_ms()
t0 = int(buf[5, 0]) if buf[5, 0] else int(buf[0][buf[0] < 2**63].min())
print(f"{wl} world={world} render={st['render_ms']:.4f} ms   phases(ms)={' '.join(f'{x:.3f}' for x in ph)}")
print("  (us after the last CTA finished staging)  p: first-out-of-trace  last-out-of-trace  last-arrived  first-saw-open  last-saw-open")
for p in range(len(ph) - 1):
    v = [(int(buf[k, p]) - t0) / 1e3 for k in range(5)]
    print(f"  p={p}: {v[0]:9.1f} {v[1]:9.1f} {v[2]:9.1f} {v[3]:9.1f} {v[4]:9.1f}")
