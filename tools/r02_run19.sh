#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
python tools/one_frame.py synthetic10m 2 1 32 && timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_synthetic10m_w1_f32.csv python tools/one_frame.py synthetic10m 2 1 32 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.DictReader([l for l in open("gpurun_out/r02_ncu_synthetic10m_w1_f32.csv") if l.startswith('"')]))
L={}
for r in rows:
    if not any(k in r["Kernel Name"] for k in ("trace_kernel","shade_kernel","combine")): continue
    d=L.setdefault(int(r["ID"]),{"name":r["Kernel Name"][:28]}); d[r["Metric Name"]]=float(r["Metric Value"].replace(",",""))
ids=sorted(L)[-13:]
for i in ids:
    x=L[i]; print(f"{x['name']:30s} {x['gpu__time_duration.sum']/1e6:8.3f} ms  warp-inst {x['smsp__inst_executed.sum']/1e9:7.3f} G  lanes {x['smsp__thread_inst_executed.sum']/x['smsp__inst_executed.sum']:5.1f}  issue {x['smsp__issue_active.avg.pct_of_peak_sustained_active']:5.1f}%  dram {(x['dram__bytes_read.sum']+x['dram__bytes_write.sum'])/1e9:6.2f} GB")
PY
