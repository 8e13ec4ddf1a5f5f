#!/bin/bash
mkdir -p gpurun_out
PROBE_SCHEDS=launches,pixel timeout 900 python tools/r02_probe.py bunny4k synthetic10m > gpurun_out/r02_probe8.log 2>&1
cat gpurun_out/r02_probe8.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest8.log 2>&1
grep -E "passed|failed|^FAILED|^E  .*Error" gpurun_out/r02_gputest8.log | cut -c1-900
