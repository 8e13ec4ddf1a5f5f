#!/usr/bin/env python
"""One upload + build of a bench workload (argv[1], default synthetic10m): for `ncu -k regex:sah|karras|...` launch lists of the build."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import cutrace_b200 as ct
s, _ = bench.load_workload(sys.argv[1] if len(sys.argv) > 1 else "synthetic10m")
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    with ct.Renderer(s) as r:
        print("build_ms", r.stats()["build_ms"] if hasattr(r, "stats") else None, flush=True)
