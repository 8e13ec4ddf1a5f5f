#!/bin/bash
# closing measurement of round 2 on one GPU: GPU tests, bench.py (both arms), per-frame ncu counters, ncu --set full of the pixel kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_close_gputest.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r02_close_gputest.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_close_n1.json 2> gpurun_out/r02_close_n1.err; echo "bench n1 exit $?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_close_ref.json 2> gpurun_out/r02_close_ref.err; echo "bench ref exit $?"
python - <<'PY'
import json
for n in ("n1", "ref"):
    try:
        d = json.loads(open(f"gpurun_out/r02_close_{n}.json").read().strip().splitlines()[-1])
        print(n, {k: round(d[k], 3) for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"].get("ms_per_frame", 0), 3), "clocks", d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("reasons"))
        for k, v in d.get("extra_workloads", {}).items():
            print("   ", k, round(v["ms_per_frame"], 4), round(v.get("e2e_ms_per_frame", 0), 3), v.get("parity_check"))
        if "roofline" in d:
            print("    roofline", round(d["roofline"]["frac"], 4), (d["roofline"].get("issue") or {}).get("frac"))
    except Exception as e:
        print(n, "ERR", e)
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for spec in "bunny4k 1" "bunny4k 2" "bunny4k 4" "bunny4k 8" "synthetic10m 1" "synthetic10m 8"; do
  set -- $spec
  python tools/one_frame.py $1 2 $2 0 > /dev/null 2>&1 || { echo "one_frame $spec failed"; continue; }
  timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_close_ncu_$1_w$2.csv python tools/one_frame.py $1 2 $2 0 > gpurun_out/r02_close_ncu_$1_w$2.log 2>&1
  echo "ncu $spec exit $?"; tail -1 gpurun_out/r02_close_ncu_$1_w$2.log
done
python tools/one_frame.py bunny4k 2 1 0 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:pixel_kernel -s 1 -c 1 -f -o gpurun_out/r02_close_pixel_bunny4k python tools/one_frame.py bunny4k 2 1 0 > gpurun_out/r02_close_ncu_full.log 2>&1; echo "ncu full exit $?"
python tools/one_frame.py synthetic10m 2 1 0 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:pixel_kernel -s 1 -c 1 -f -o gpurun_out/r02_close_pixel_synthetic10m python tools/one_frame.py synthetic10m 2 1 0 > gpurun_out/r02_close_ncu_full2.log 2>&1; echo "ncu full2 exit $?"
