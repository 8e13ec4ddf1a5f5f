#!/bin/bash
# lane refill of the pixel kernel (hall, and what it costs the staged scenes), CTA shapes on the hall, threads per CTA on tiny frames
mkdir -p gpurun_out
L=gpurun_out/r02_run28.log
{
echo "== refill 0/1"
timeout 600 python tools/refill_probe.py synthetic10m bunny4k mirror1080 spheres1080
PROBE_WORLDS=1 timeout 100 python tools/refill_probe.py triangle
echo "== hall, CTA shapes"
for v in t512x1 t768x1 t512x2; do
  CUTRACE_B200_LIB=$PWD/cutrace_b200/lib/variants/libcutrace_b200_$v.so PROBE_WORLDS=1 timeout 300 python tools/refill_probe.py synthetic10m
done
echo "== tiny frames: threads per CTA"
for t in 1024 512 256 128; do echo "-- threads $t"; CUTRACE_DEBUG_PIXEL_THREADS=$t timeout 120 python tools/tiny_probe.py triangle spheres1080 mirror1080; done
echo "== stamps (triangle.json)"
CUTRACE_B200_LIB=$PWD/cutrace_b200/lib/variants/libcutrace_b200_stamps.so timeout 100 python tools/stamps_probe.py | tail -4
for t in 256; do CUTRACE_DEBUG_PIXEL_THREADS=$t CUTRACE_B200_LIB=$PWD/cutrace_b200/lib/variants/libcutrace_b200_stamps.so timeout 100 python tools/stamps_probe.py | tail -3; done
} > $L 2>&1
cat $L
