#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest10.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r02_gputest10.log | cut -c1-200
grep -E "^E  .*Error" gpurun_out/r02_gputest10.log | cut -c1-420 | sort | uniq | head
