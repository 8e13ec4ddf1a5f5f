#!/bin/bash
# ncu --set full of the pixel kernel on the hall and on bunny.json 4K (source-level counters for the tuning log)
mkdir -p gpurun_out
python tools/one_frame.py synthetic10m 2 1 0 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:pixel_kernel -s 1 -c 1 -f -o gpurun_out/r02c_pixel_synthetic10m python tools/one_frame.py synthetic10m 2 1 0 > gpurun_out/r02c_ncu_full2.log 2>&1; echo "ncu hall exit $?"
python tools/one_frame.py bunny4k 2 1 0 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:pixel_kernel -s 1 -c 1 -f -o gpurun_out/r02c_pixel_bunny4k python tools/one_frame.py bunny4k 2 1 0 > gpurun_out/r02c_ncu_full.log 2>&1; echo "ncu bunny exit $?"
ls -la gpurun_out/*.ncu-rep
