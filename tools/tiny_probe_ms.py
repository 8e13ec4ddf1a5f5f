#!/usr/bin/env python
"""render_ms (median of 12 frames) of bench workloads with the library $CUTRACE_B200_LIB selects"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import cutrace_b200 as ct
for name in sys.argv[1:]:
    s, _ = bench.load_workload(name)
    with ct.Renderer(s) as r:
        ms = [r.render()["render_ms"] for _ in range(12 if name != "synthetic10m" else 5)]
    print(f"  {name:14s} median {np.median(ms[2:]):9.4f} ms   min {np.min(ms[2:]):9.4f} ms", flush=True)
