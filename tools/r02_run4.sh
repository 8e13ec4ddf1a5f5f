#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/r02_probe.py > gpurun_out/r02_probe4.log 2>&1
cat gpurun_out/r02_probe4.log
