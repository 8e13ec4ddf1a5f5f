#!/bin/bash
mkdir -p gpurun_out
export PROBE_SCHEDS=pixel PROBE_WORLDS=8
{
for g in 296 222 148 74; do echo "== grid $g"; CUTRACE_DEBUG_PIXEL_GRID=$g timeout 300 python tools/r02_probe.py bunny4k mirror1080 spheres1080 synthetic10m; done
echo "== N=1 grid 148"; PROBE_WORLDS=1 CUTRACE_DEBUG_PIXEL_GRID=148 timeout 300 python tools/r02_probe.py bunny4k
} > gpurun_out/r02_probe18.log 2>&1
cat gpurun_out/r02_probe18.log
M=gpu__time_duration.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_triangle.csv python tools/one_frame.py triangle 5 1 0 > /dev/null 2>&1
grep -E "pixel_kernel" gpurun_out/r02_ncu_triangle.csv | tail -3 | cut -c1-300
python - <<'PY'
import sys; sys.path.insert(0,'.')
import bench
from oracle import pyoracle as po
s,w=bench.load_workload("triangle")
import subprocess
PY
