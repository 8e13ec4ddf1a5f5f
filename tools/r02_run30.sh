#!/bin/bash
# e2e phase breakdown (timing build) for bunny.json 4K, default build sanity on all workloads
mkdir -p gpurun_out
L=gpurun_out/r02_run30.log
{
echo "== e2e phases, bunny4k (timing build)"
CUTRACE_B200_LIB=$PWD/cutrace_b200/lib/variants/libcutrace_b200_timing.so timeout 120 python tools/e2e_probe.py bunny4k 2>&1 | tail -75
echo "== e2e, default build"
timeout 120 python tools/e2e_probe.py bunny4k 2>&1 | tail -4
echo "== defaults"
timeout 300 python tools/tiny_probe_ms.py bunny4k mirror1080 spheres1080 triangle synthetic10m
} > $L 2>&1
cat $L
