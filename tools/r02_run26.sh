#!/bin/bash
python - <<'PY'
import sys, os; sys.path.insert(0,'.')
import numpy as np, bench, cutrace_b200 as ct
for wl in ("mirror1080","bunny4k","synthetic10m"):
    s,_=bench.load_workload(wl)
    for leaf in (2,3,4,6,8):
        with ct.Renderer(s, leaf_size=leaf) as r:
            ms=[r.render()["render_ms"] for _ in range(5 if wl!="synthetic10m" else 3)]
            st=r.stats()
        print(wl, "leaf", leaf, "render", round(float(np.median(ms[1:])),4), "nodes", st["bvh_nodes"], "depth", st["bvh_depth"], "build", round(st["build_ms"],2), flush=True)
PY
