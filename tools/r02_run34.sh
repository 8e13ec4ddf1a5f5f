#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_run34.log
{
PROBE_COMBOS=dd timeout 400 python tools/seg_probe.py synthetic10m bunny4k mirror1080 spheres1080
PROBE_COMBOS=dd PROBE_WORLDS=1 timeout 100 python tools/seg_probe.py triangle
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "soup or goldens or edge_cases or frame_kernel_equals" 2>&1 | tail -4
} > $L 2>&1
cat $L
