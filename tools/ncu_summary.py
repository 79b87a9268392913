#!/usr/bin/env python3
"""Condense `ncu --page raw --csv` dumps into the one-table summary kept under profiles/.
usage: ncu_summary.py out.csv tag=raw.csv [tag=raw.csv ...]"""
import csv
import sys

COLS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
out = csv.writer(open(sys.argv[1], "w"))
out.writerow(["capture", "kernel"] + COLS)
units_row = None
for arg in sys.argv[2:]:
    tag, path = arg.rsplit("=", 1)
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    if units_row is None:
        units_row = ["", "(units)"] + [units[hdr.index(c)] if c in hdr else "" for c in COLS]
        out.writerow(units_row)
    for r in rows[2:]:
        out.writerow([tag, r[hdr.index("Kernel Name")].split("(")[0]] + [r[hdr.index(c)] if c in hdr else "" for c in COLS])
