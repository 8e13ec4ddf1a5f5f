#!/bin/bash
# the round's closing measurement: GPU tests, then bench.py at N = 8, 4, 2, 1 and the reference arm, back to back on one box
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_final_gputest.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r02_final_gputest.log | cut -c1-200
for n in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/r02_final_n$n.json 2> gpurun_out/r02_final_n$n.err; echo "bench n$n exit $?"
done
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r02_final_n1.json 2> gpurun_out/r02_final_n1.err; echo "bench n1 exit $?"
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_final_ref.json 2> gpurun_out/r02_final_ref.err; echo "bench ref exit $?"
python - <<'PY'
import json
for n in ("n8","n4","n2","n1","ref"):
    try:
        d=json.loads(open(f"gpurun_out/r02_final_{n}.json").read().strip().splitlines()[-1])
        print(n, {k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["ms_per_frame"],3), d.get("parity_check",{}).get("n_gpu_equals_1_gpu"), d["e2e"].get("parity_check",{}).get("n_gpu_equals_1_gpu"), "clocks", d.get("clocks",{}).get("sm_mhz"), d.get("clocks",{}).get("reasons"))
        for k,v in d.get("extra_workloads",{}).items(): print("   ", k, round(v["ms_per_frame"],4), round(v.get("e2e_ms_per_frame",0),3), v.get("parity_check",{}).get("n_gpu_equals_1_gpu"), v.get("e2e_parity_check",{}).get("n_gpu_equals_1_gpu"))
        if "roofline" in d: print("    roofline", round(d["roofline"]["frac"],4), (d["roofline"].get("issue") or {}).get("frac"))
    except Exception as e: print(n, "ERR", e)
PY
