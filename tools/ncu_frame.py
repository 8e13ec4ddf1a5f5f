#!/usr/bin/env python
"""Per-FRAME totals of an `ncu --metrics ... --csv` launch list of tools/one_frame.py (the launches of the last rendered frame):
thread / warp instructions, DRAM bytes, per-kernel time — merged into profiles/ncu_frame.json, which bench.py reads for
roofline.issue / roofline.traffic.
usage: tools/ncu_frame.py <metrics.csv> <workload name> <n gpus the shard stands for> "<note>" """
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

RENDER = ("pixel_kernel", "frame_kernel", "trace_kernel", "shade_kernel", "combine_levels_kernel", "export_gbuffer_kernel")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}


def main():
    path, wl, n, note = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.DictReader(lines))
    # long format: one row per (launch ID, metric)
    launches = {}
    for r in rows:
        name = r["Kernel Name"]
        if not any(k in name for k in RENDER):
            continue
        d = launches.setdefault(int(r["ID"]), {"name": next(k for k in RENDER if k in name)})
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        d[r["Metric Name"]] = v
    ids = sorted(launches)
    # the last frame: walk back until the first render kernel of a frame (pixel/frame kernel, or trace level 0 = first trace after a combine)
    names = [launches[i]["name"] for i in ids]
    per_frame = len(ids)
    for k in range(1, len(ids) + 1):
        if len(ids) % k == 0 and names[:k] * (len(ids) // k) == names:
            per_frame = k
            break
    last = [launches[i] for i in ids[-per_frame:]]
    tot = lambda m: sum(x.get(m, 0.0) for x in last)  # noqa: E731
    kernels = {}
    for x in last:
        k = kernels.setdefault(x["name"], {"launches": 0, "ms": 0.0, "thread_inst": 0.0, "warp_inst": 0.0, "dram_bytes": 0.0, "issue_busy_pct_time_weighted": 0.0})
        k["launches"] += 1
        k["ms"] += x.get("gpu__time_duration.sum", 0.0) / 1e6
        k["thread_inst"] += x.get("smsp__thread_inst_executed.sum", 0.0)
        k["warp_inst"] += x.get("smsp__inst_executed.sum", 0.0)
        k["dram_bytes"] += x.get("dram__bytes_read.sum", 0.0) + x.get("dram__bytes_write.sum", 0.0)
        k["issue_busy_pct_time_weighted"] += x.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0) * x.get("gpu__time_duration.sum", 0.0)
    for k in kernels.values():
        k["issue_busy_pct_time_weighted"] = k["issue_busy_pct_time_weighted"] / max(1e-9, k["ms"] * 1e6)
        k["active_lanes"] = k["thread_inst"] / max(1.0, k["warp_inst"])
    out = {"launches_per_frame": per_frame, "frames_captured": len(ids) // per_frame, "kernel_ms_under_ncu": tot("gpu__time_duration.sum") / 1e6,
           "thread_inst_per_frame": tot("smsp__thread_inst_executed.sum"), "warp_inst_per_frame": tot("smsp__inst_executed.sum"),
           "active_lanes": tot("smsp__thread_inst_executed.sum") / max(1.0, tot("smsp__inst_executed.sum")),
           "dram_bytes_per_frame": tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum"), "kernels": kernels, "source": note}
    label = bench.WORKLOADS[wl]["label"]
    jp = os.path.join(ROOT, "profiles", "ncu_frame.json")
    allw = json.load(open(jp)) if os.path.exists(jp) else {}
    allw.setdefault(label, {})[f"n{n}"] = out
    json.dump(allw, open(jp, "w"), indent=1)
    print(json.dumps({label: {f"n{n}": out}}, indent=1))


if __name__ == "__main__":
    main()
