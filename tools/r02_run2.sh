#!/bin/bash
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/atomic_bench tools/atomic_bench.cu && /tmp/atomic_bench > gpurun_out/r02_atomics.txt 2>&1
export CUTRACE_B200_LIB=$PWD/cutrace_b200/lib/variants/libcutrace_b200_dbg.so
for w in "triangle 1" "mirror1080 8" "mirror1080 1" "bunny4k 8" "bunny4k 1"; do
  timeout 120 python tools/phase_debug.py $w
done > gpurun_out/r02_phase_debug.txt 2>&1
unset CUTRACE_B200_LIB
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r02_gputest2.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_gputest2.log
cat gpurun_out/r02_atomics.txt gpurun_out/r02_phase_debug.txt; tail -30 gpurun_out/r02_gputest2.log
