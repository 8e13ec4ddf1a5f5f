#!/bin/bash
mkdir -p gpurun_out
export CUTRACE_B200_LIB=$PWD/cutrace_b200/lib/variants/libcutrace_b200_bvh4.so
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not cli and not integration" > gpurun_out/r02_gputest24.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r02_gputest24.log | cut -c1-200
grep -E "^E  " gpurun_out/r02_gputest24.log | cut -c1-300 | head -5
export PROBE_SCHEDS=pixel,launches
timeout 600 python tools/r02_probe.py mirror1080 bunny4k synthetic10m
