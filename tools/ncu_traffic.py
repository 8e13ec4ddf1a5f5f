#!/usr/bin/env python
"""Per-kernel means of DRAM traffic and the issue / divergence metrics from an .ncu-rep, merged into
profiles/ncu_traffic.json under a workload label (bench.py reads roofline.traffic from there).
usage: tools/ncu_traffic.py <report.ncu-rep> "<workload label>" "<source note>" """
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = {
    "dram__bytes_read.sum": "read", "dram__bytes_write.sum": "write",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_warp_instruction",
    "smsp__inst_executed.sum": "warp_instructions_per_launch",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "gpu__time_duration.sum": "duration_ns",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}


def main():
    rep, label, note = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    acc = {}
    for r in data:
        name = r[col["Kernel Name"]]
        key = "shade_kernel" if "shade_kernel" in name else "trace_kernel" if "trace_kernel" in name else None
        if not key:
            continue
        d = acc.setdefault(key, {v: [] for v in METRICS.values()})
        for m, v in METRICS.items():
            x = float(r[col[m]].replace(",", "")) * UNIT.get(units[col[m]], 1.0)
            d[v].append(x)
    out = {}
    for key, d in acc.items():
        n = len(d["read"])
        mean = {k: sum(v) / n for k, v in d.items()}
        out[key] = {"mean_traffic_bytes": mean["read"] + mean["write"], "issue_slots_busy_pct": mean["issue_slots_busy_pct"],
                    "active_threads_per_warp_instruction": round(mean["active_threads_per_warp_instruction"], 2),
                    "warp_instructions_per_launch": mean["warp_instructions_per_launch"], "dram_throughput_pct": mean["dram_throughput_pct"],
                    "mean_duration_ms": mean["duration_ns"] / 1e6, "launches_captured": n}
    out["source"] = note
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    allw = json.load(open(path)) if os.path.exists(path) else {}
    allw[label] = out
    json.dump(allw, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
