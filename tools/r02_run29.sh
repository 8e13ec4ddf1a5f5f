#!/bin/bash
# per-CTA work cursors x super-tile slot order; launch floor of a tiny frame
mkdir -p gpurun_out
L=gpurun_out/r02_run29.log
{
timeout 600 python tools/seg_probe.py synthetic10m bunny4k mirror1080 spheres1080
PROBE_WORLDS=1 timeout 100 python tools/seg_probe.py triangle
echo "== tiny frames"
timeout 120 python tools/tiny_probe.py triangle
CUTRACE_DEBUG_NO_HOST_STATS=1 timeout 120 python tools/tiny_probe.py triangle
echo "== launch floor"
timeout 60 tools/launch_floor.bin
} > $L 2>&1
cat $L
