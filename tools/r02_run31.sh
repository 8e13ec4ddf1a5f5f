#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_run31.log
{
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "render_download or peer_frame or tile_sharding or untile or goldens or pixel_batches or set_camera" 2>&1 | tail -5
echo "== e2e, default build"
timeout 120 python tools/e2e_probe.py bunny4k 2>&1 | tail -3
echo "== defaults"
timeout 300 python tools/tiny_probe_ms.py bunny4k mirror1080 spheres1080 triangle synthetic10m
} > $L 2>&1
cat $L
