#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest6.log 2>&1
grep -E "passed|failed|^FAILED|^E  .*Error" gpurun_out/r02_gputest6.log | cut -c1-600
timeout 900 python bench.py --steps 10 --warmup 3 --verbose > gpurun_out/r02_bench6_n1.json 2> gpurun_out/r02_bench6_n1.err; echo "bench n1 exit $?"; tail -3 gpurun_out/r02_bench6_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench6_ref.json 2> gpurun_out/r02_bench6_ref.err; echo "bench ref exit $?"; tail -3 gpurun_out/r02_bench6_ref.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench6_n1.json","gpurun_out/r02_bench6_ref.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d[k] for k in ("value","ms_per_step") if k in d}, "e2e", d.get("e2e",{}).get("ms_per_frame"))
        print("  extra", {k:(round(v["ms_per_frame"],4), round(v.get("e2e_ms_per_frame",0),3)) for k,v in d.get("extra_workloads",{}).items()})
        print("  roofline", {k:v for k,v in d.get("roofline",{}).items() if k in ("achieved","frac","compulsory_bytes_per_frame")}, d.get("kernel_ms"))
    except Exception as e: print(f, "ERR", e)
PY
