#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest13.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r02_gputest13.log | cut -c1-200
grep -E "^E  .*Error" gpurun_out/r02_gputest13.log | cut -c1-420 | sort | uniq | head
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --verbose > gpurun_out/r02_bench13_n2.json 2> gpurun_out/r02_bench13_n2.err; echo "bench n2 exit $?"
tail -25 gpurun_out/r02_bench13_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_bench13_n2.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["ms_per_frame"], d.get("parity_check"), d["e2e"].get("parity_check"))
    for k,v in d["extra_workloads"].items(): print(k, round(v["ms_per_frame"],4), round(v["e2e_ms_per_frame"],3), v.get("parity_check",{}).get("n_gpu_equals_1_gpu"), v.get("e2e_parity_check",{}).get("n_gpu_equals_1_gpu"))
except Exception as e: print("ERR", e)
PY
