#!/usr/bin/env python
"""ms/frame of bench workloads for leaf sizes 1..8 (SAH treelet build): tools/leaf_sweep.py [workload ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import cutrace_b200 as ct
for name in (sys.argv[1:] or ("bunny4k", "synthetic10m")):
    s, _ = bench.load_workload(name)
    for leaf in (1, 2, 3, 4, 5, 6, 8):
        with ct.Renderer(s, leaf_size=leaf) as r:
            ms = [r.render() for _ in range(6 if name != "synthetic10m" else 4)]
        print(name, "leaf", leaf, "render_ms", round(float(np.median([m["render_ms"] for m in ms[1:]])), 3), "nodes", ms[-1]["bvh_nodes"], "depth", ms[-1]["bvh_depth"], flush=True)
