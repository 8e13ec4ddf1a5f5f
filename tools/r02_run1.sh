#!/bin/bash
# first round-2 GPU call: environment facts, the GPU test suite, the scheduler probe, a short bench
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
  nproc; free -g | head -2
  df -h /dev/shm | tail -1
  python - <<'PY'
import os, mmap
try:
    fd = os.memfd_create("ctb_probe")
    os.ftruncate(fd, 1 << 30)
    m = mmap.mmap(fd, 1 << 30)
    m[0:4] = b"abcd"
    p = f"/proc/{os.getpid()}/fd/{fd}"
    fd2 = os.open(p, os.O_RDWR)
    m2 = mmap.mmap(fd2, 1 << 30)
    print("memfd 1 GiB ok, reopen via /proc ok:", m2[0:4])
except Exception as e:
    print("memfd failed:", e)
PY
} > gpurun_out/r02_env.txt 2>&1
timeout 2000 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02_gputest1.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_gputest1.log
timeout 1200 python tools/r02_probe.py > gpurun_out/r02_probe1.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
tail -5 gpurun_out/r02_gputest1.log
cat gpurun_out/r02_probe1.log
