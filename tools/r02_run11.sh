#!/bin/bash
mkdir -p gpurun_out
export PROBE_SCHEDS=pixel
{
echo "== default (256x4, smem)"; timeout 300 python tools/r02_probe.py triangle spheres1080 mirror1080 bunny4k
echo "== default, global (FLAG_NO_SMEM_TOP)"; PROBE_FLAGS=1 timeout 300 python tools/r02_probe.py mirror1080 bunny4k
for v in p3 p2 p512 p384; do
  echo "== $v"; CUTRACE_B200_LIB=$PWD/cutrace_b200/lib/variants/libcutrace_b200_$v.so timeout 300 python tools/r02_probe.py spheres1080 mirror1080 bunny4k
  echo "== $v global"; PROBE_FLAGS=1 CUTRACE_B200_LIB=$PWD/cutrace_b200/lib/variants/libcutrace_b200_$v.so timeout 300 python tools/r02_probe.py bunny4k
done
echo "== synthetic pixel vs others"; PROBE_SCHEDS=pixel,launches,frame PROBE_WORLDS=1 timeout 300 python tools/r02_probe.py synthetic10m
} > gpurun_out/r02_probe11.log 2>&1
cat gpurun_out/r02_probe11.log
