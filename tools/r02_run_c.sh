#!/bin/bash
# round-2 check run C: new tests, build time, bunny/hall frame time, tiny-scene kernel durations
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/build_ms.py ${BUILD_MS_ARGS:-bunny4k} 2>&1 | tail -4
python tools/tiny_probe.py 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tiny_launches.csv python tools/tiny_probe.py > gpurun_out/tiny_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/tiny_launches.csv")) if len(r) > 10 and r[0].isdigit()]
d = collections.defaultdict(list)
for r in rows:
    d[r[4][:60]].append(float(r[-1]))
for k, v in d.items():
    v = sorted(v)
    print(f"{k:62s} n={len(v):3d} min {v[0]/1e3:9.2f} us  median {v[len(v)//2]/1e3:9.2f} us")
PY
