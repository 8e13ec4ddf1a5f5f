#!/bin/bash
export PROBE_SCHEDS=pixel PROBE_WORLDS=1
echo "== default"; timeout 300 python tools/r02_probe.py triangle spheres1080
echo "== brute (flag 4)"; PROBE_FLAGS=4 timeout 300 python tools/r02_probe.py triangle spheres1080
M=gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_spheres.csv python tools/one_frame.py spheres1080 2 1 0 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.DictReader([l for l in open("gpurun_out/r02_ncu_spheres.csv") if l.startswith('"')]))
L={}
for r in rows:
    if "pixel_kernel" not in r["Kernel Name"]: continue
    d=L.setdefault(int(r["ID"]),{}); d[r["Metric Name"]]=float(r["Metric Value"].replace(",",""))
x=L[sorted(L)[-1]]
print("spheres1080 pixel kernel:", x['gpu__time_duration.sum']/1e6, "ms warp-inst", x['smsp__inst_executed.sum']/1e9, "G lanes", x['smsp__thread_inst_executed.sum']/x['smsp__inst_executed.sum'], "issue", x['smsp__issue_active.avg.pct_of_peak_sustained_active'])
PY
