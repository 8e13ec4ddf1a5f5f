#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest23.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r02_gputest23.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python tools/config_report.py > gpurun_out/r02_configs.log 2>&1; tail -12 gpurun_out/r02_configs.log
