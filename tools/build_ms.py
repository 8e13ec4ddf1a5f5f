#!/usr/bin/env python
"""build_ms (upload + LBVH + SAH treelets) and resident ms/frame of the bunny 4K and hall workloads, with and without the SAH pass."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import cutrace_b200 as ct
for name in (sys.argv[1:] or ("bunny4k", "synthetic10m")):
    s, _ = bench.load_workload(name)
    for sah in (1, 0):
        os.environ.pop("CUTRACE_DEBUG_NO_SAH", None)
        if not sah:
            os.environ["CUTRACE_DEBUG_NO_SAH"] = "1"
        b = []
        for rep in range(3):
            with ct.Renderer(s, flags=ct.FLAG_VALIDATE_BVH if rep == 0 else 0) as r:
                ms = [r.render() for _ in range(5)]
                b.append(ms[-1]["build_ms"])
        print(name, "sah", sah, "build_ms", [round(x, 3) for x in b], "render_ms", round(float(np.median([m["render_ms"] for m in ms[1:]])), 3), flush=True)
