#!/bin/bash
mkdir -p gpurun_out
for v in share1 share2 share3; do
  echo "== $v"
  CUTRACE_B200_LIB=$PWD/cutrace_b200/lib/variants/libcutrace_b200_$v.so timeout 600 python tools/r02_probe.py mirror1080 bunny4k synthetic10m
done > gpurun_out/r02_probe5.log 2>&1
cat gpurun_out/r02_probe5.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest5.log 2>&1
grep -E "passed|failed|^FAILED|^E  .*Error" gpurun_out/r02_gputest5.log | cut -c1-600
