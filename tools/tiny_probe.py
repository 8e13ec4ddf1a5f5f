#!/usr/bin/env python
"""triangle.json 20x20 and sphere_plane.json 1080p: ms/frame of this path (render_ms of 40 frames, median) and of the reference's own
kernel (oracle/_ref, 5 launches); run under `ncu --metrics gpu__time_duration.sum` for the kernels' own durations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
import cutrace_b200 as ct
from oracle import pyoracle as po
for name in sys.argv[1:] or ("triangle", "spheres1080"):
    s, _ = bench.load_workload(name)
    with ct.Renderer(s) as r:
        ms = [r.render()["render_ms"] for _ in range(40)]
    ref = po.ref_gpu_render(s, iters=5, warmup=2)
    print(name, "this path median", round(float(np.median(ms[5:])) * 1e3, 2), "us  min", round(float(np.min(ms[5:])) * 1e3, 2), "us   reference", round(ref["render_ms"] * 1e3, 2), "us", flush=True)
