#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02_gputest3.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_gputest3.log
timeout 900 python tools/r02_probe.py > gpurun_out/r02_probe3.log 2>&1
CUTRACE_B200_LIB=$PWD/cutrace_b200/lib/variants/libcutrace_b200_nogate.so timeout 600 python tools/r02_probe.py bunny4k synthetic10m > gpurun_out/r02_probe3_nogate.log 2>&1
export CUTRACE_B200_LIB=$PWD/cutrace_b200/lib/variants/libcutrace_b200_dbg.so
for w in "triangle 1" "mirror1080 8" "bunny4k 8"; do timeout 120 python tools/phase_debug.py $w; done > gpurun_out/r02_phase_debug3.txt 2>&1
grep -E "passed|failed|^FAILED|^E  .*Error" gpurun_out/r02_gputest3.log | cut -c1-600
cat gpurun_out/r02_probe3.log; echo NOGATE; cat gpurun_out/r02_probe3_nogate.log; cat gpurun_out/r02_phase_debug3.txt
