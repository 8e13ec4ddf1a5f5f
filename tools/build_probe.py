#!/usr/bin/env python
"""Uploads the 10 M-triangle scene a few times (LBVH build) — with the -DCTB_TIMING variant it prints the phases."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cutrace_b200 as ct
import bench
scene, wl = bench.load_workload(sys.argv[1] if len(sys.argv) > 1 else "synthetic10m")
scene = scene.with_resolution(64, 64)
for i in range(3):
    t = time.perf_counter()
    with ct.Renderer(scene) as r:
        print("upload", round(1e3 * (time.perf_counter() - t), 2), "ms  build_ms", r.stats()["build_ms"], "nodes", r.stats()["bvh_nodes"], flush=True)
