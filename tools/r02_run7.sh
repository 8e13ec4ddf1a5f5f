#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gputest7.log 2>&1
grep -E "passed|failed|^FAILED|^E  .*Error" gpurun_out/r02_gputest7.log | cut -c1-700
timeout 600 python tools/r02_probe.py triangle spheres1080 mirror1080 > gpurun_out/r02_probe7.log 2>&1
cat gpurun_out/r02_probe7.log
