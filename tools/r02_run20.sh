#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest20.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r02_gputest20.log | cut -c1-200
grep -E "^E  " gpurun_out/r02_gputest20.log | cut -c1-300 | head -5
export PROBE_SCHEDS=pixel,launches PROBE_WORLDS=1
{
echo "== prepass"; timeout 300 python tools/r02_probe.py triangle spheres1080 mirror1080 bunny4k synthetic10m
echo "== no prepass"; CUTRACE_B200_LIB=$PWD/cutrace_b200/lib/variants/libcutrace_b200_noprepass.so timeout 300 python tools/r02_probe.py mirror1080 bunny4k synthetic10m
} > gpurun_out/r02_probe20.log 2>&1
cat gpurun_out/r02_probe20.log
