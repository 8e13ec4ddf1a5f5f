#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest17.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r02_gputest17.log | cut -c1-200
grep -E "^E  " gpurun_out/r02_gputest17.log | cut -c1-300 | head -5
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench17_n1.json 2> gpurun_out/r02_bench17_n1.err; echo "bench n1 exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench17_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["ms_per_frame"], d["kernel_ms"])
for k,v in d["extra_workloads"].items(): print(k, round(v["ms_per_frame"],4), round(v["e2e_ms_per_frame"],3))
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for spec in "bunny4k 1 0" "bunny4k 8 0" "synthetic10m 1 0" "bunny4k 1 32"; do
  set -- $spec
  python tools/one_frame.py $1 2 $2 $3 > /dev/null 2>&1 || { echo "one_frame $spec failed"; continue; }
  timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_$1_w$2_f$3.csv python tools/one_frame.py $1 2 $2 $3 > gpurun_out/r02_ncu_$1_w$2_f$3.log 2>&1
  echo "ncu $spec exit $?"; tail -1 gpurun_out/r02_ncu_$1_w$2_f$3.log
done
python tools/one_frame.py bunny4k 2 1 0 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:pixel_kernel -s 1 -c 1 -o gpurun_out/r02_pixel_bunny4k python tools/one_frame.py bunny4k 2 1 0 > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/*.ncu-rep
