#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_run35.log
{
echo "== default"; PROBE_COMBOS=dd timeout 300 python tools/seg_probe.py synthetic10m bunny4k mirror1080
for v in "$@"; do
  echo "== $v"; CUTRACE_B200_LIB=$PWD/cutrace_b200/lib/variants/libcutrace_b200_$v.so PROBE_COMBOS=dd timeout 300 python tools/seg_probe.py synthetic10m bunny4k mirror1080
done
} > $L 2>&1
cat $L
