#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest9.log 2>&1
grep -E "passed|failed|^FAILED|^E  .*Error" gpurun_out/r02_gputest9.log | cut -c1-900
PROBE_SCHEDS=launches,pixel timeout 900 python tools/r02_probe.py triangle mirror1080 bunny4k > gpurun_out/r02_probe9.log 2>&1
cat gpurun_out/r02_probe9.log
