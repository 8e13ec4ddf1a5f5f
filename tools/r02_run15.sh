#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 900 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --verbose > gpurun_out/r02_bench15_$tag.json 2> gpurun_out/r02_bench15_$tag.err; echo "bench $tag exit $?"
grep -E "^\[rank|Error" gpurun_out/r02_bench15_$tag.err | head -12
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench15_$tag.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["ms_per_frame"], d.get("parity_check",{}).get("n_gpu_equals_1_gpu"), d["e2e"].get("parity_check",{}).get("n_gpu_equals_1_gpu"))
    for k,v in d["extra_workloads"].items(): print(k, round(v["ms_per_frame"],4), round(v["e2e_ms_per_frame"],3), v.get("parity_check",{}).get("n_gpu_equals_1_gpu"), v.get("e2e_parity_check",{}).get("n_gpu_equals_1_gpu"))
except Exception as e: print("ERR", e)
PY
}
run wide X=1
run narrow CUTRACE_DEBUG_NARROW_WARPS=1
