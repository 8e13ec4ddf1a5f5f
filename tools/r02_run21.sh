#!/bin/bash
mkdir -p gpurun_out
for variant in "host" "nccl --nccl-frame-barrier"; do
set -- $variant; tag=$1; shift
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 30 --warmup 5 --verbose --no-extra "$@" > gpurun_out/r02_bench21_n8_$tag.json 2> gpurun_out/r02_bench21_n8_$tag.err; echo "bench n8 $tag exit $?"
grep -E "^\[rank 0|Error" gpurun_out/r02_bench21_n8_$tag.err | head -4
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench21_n8_$tag.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["ms_per_frame"], d.get("parity_check",{}).get("n_gpu_equals_1_gpu"), d["e2e"].get("parity_check",{}).get("n_gpu_equals_1_gpu"))
except Exception as e: print("ERR", e)
PY
done
