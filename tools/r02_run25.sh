#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_gputest25.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r02_gputest25.log | cut -c1-200
grep -E "^E  " gpurun_out/r02_gputest25.log | cut -c1-400 | head -6
export PROBE_SCHEDS=pixel PROBE_WORLDS=1
echo "== SAH treelets"; timeout 600 python tools/r02_probe.py mirror1080 bunny4k synthetic10m
echo "== Karras only"; CUTRACE_DEBUG_NO_SAH=1 timeout 600 python tools/r02_probe.py mirror1080 bunny4k synthetic10m
python - <<'PY'
import sys, os; sys.path.insert(0,'.')
import bench, cutrace_b200 as ct
for wl in ("bunny4k","synthetic10m"):
    s,_=bench.load_workload(wl)
    for env in ("", "1"):
        if env: os.environ["CUTRACE_DEBUG_NO_SAH"]="1"
        else: os.environ.pop("CUTRACE_DEBUG_NO_SAH",None)
        with ct.Renderer(s) as r: pass
        with ct.Renderer(s) as r:
            st=r.stats(); print(wl, "NO_SAH" if env else "SAH", "build_ms", round(st["build_ms"],3), "nodes", st["bvh_nodes"], "depth", st["bvh_depth"])
PY
