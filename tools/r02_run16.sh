#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest16.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r02_gputest16.log | cut -c1-200
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 5 --verbose > gpurun_out/r02_bench16_n$n.json 2> gpurun_out/r02_bench16_n$n.err; echo "bench n$n exit $?"
grep -E "^\[rank 0|Error" gpurun_out/r02_bench16_n$n.err | head -8
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench16_n$n.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["ms_per_frame"], d.get("parity_check",{}).get("n_gpu_equals_1_gpu"), d["e2e"].get("parity_check",{}).get("n_gpu_equals_1_gpu"))
    for k,v in d["extra_workloads"].items(): print(k, round(v["ms_per_frame"],4), round(v["e2e_ms_per_frame"],3), v.get("parity_check",{}).get("n_gpu_equals_1_gpu"), v.get("e2e_parity_check",{}).get("n_gpu_equals_1_gpu"))
except Exception as e: print("ERR", e)
PY
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench16_n1.json 2> gpurun_out/r02_bench16_n1.err; echo "bench n1 exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench16_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["ms_per_frame"], d["kernel_ms"])
for k,v in d["extra_workloads"].items(): print(k, round(v["ms_per_frame"],4), round(v["e2e_ms_per_frame"],3))
PY
