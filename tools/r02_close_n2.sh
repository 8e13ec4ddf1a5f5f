#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_close_n2.json 2> gpurun_out/r02_close_n2.err; echo "bench n2 exit $?"
tail -3 gpurun_out/r02_close_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_close_n2.json").read().strip().splitlines()[-1])
    print({k: round(d[k], 3) for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"]["ms_per_frame"], 3), d.get("parity_check", {}).get("n_gpu_equals_1_gpu"), d["e2e"].get("parity_check", {}).get("n_gpu_equals_1_gpu"), d.get("clocks", {}).get("sm_mhz"))
    for k, v in d.get("extra_workloads", {}).items():
        print("   ", k, round(v["ms_per_frame"], 4), round(v.get("render_device_ms", 0), 4), round(v.get("e2e_ms_per_frame", 0), 3), v.get("parity_check", {}).get("n_gpu_equals_1_gpu"), v.get("e2e_parity_check", {}).get("n_gpu_equals_1_gpu"))
except Exception as e:
    print("ERR", e)
PY
