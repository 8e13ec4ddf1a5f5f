#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, no GPU needed) into a markdown table for profiles/.
usage: tools/ncu_summary.py <report.ncu-rep> [title]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers)"), ("launch__occupancy_limit_shared_mem", "occupancy limit (smem)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy % of peak warps"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instruction (divergence)"),
    ("smsp__thread_inst_executed_per_inst_executed.pct", "... as % of 32"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_bytes.sum", "L1 bytes"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local loads (stack + spills)"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local stores"),
]
STALLS = "smsp__average_warps_issue_stalled_{}_per_issue_active.ratio"
STALL_NAMES = ["no_instruction", "wait", "branch_resolving", "long_scoreboard", "short_scoreboard", "math_pipe_throttle",
               "mio_throttle", "lg_throttle", "barrier", "not_selected", "dispatch_stall", "imc_miss", "tex_throttle", "drain", "sleeping", "membar"]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# {title}\n")
    print(f"source: `{rep}` (`ncu --set full --clock-control none --import-source on`); values per launch\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## {name[:110]}\n")
        print("| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in hdr:
                v = r[hdr.index(k)]
                try:
                    f = float(v)
                    v = f"{f:,.0f}" if abs(f) >= 1000 else f"{f:.3f}".rstrip("0").rstrip(".")
                except ValueError:
                    pass
                print(f"| {label} (`{k}`) | {v} {units[hdr.index(k)]} |")
        print("\n| stall reason (warp-cycles per issued instruction) | value |\n|---|---|")
        st = []
        for s in STALL_NAMES:
            k = STALLS.format(s)
            if k in hdr:
                try:
                    st.append((float(r[hdr.index(k)]), s))
                except ValueError:
                    pass
        for v, s in sorted(st, reverse=True):
            if v >= 0.01:
                print(f"| {s} | {v:.3f} |")
        print()


if __name__ == "__main__":
    main()
