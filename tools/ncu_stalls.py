#!/usr/bin/env python3
"""Aggregate the warp-stall samples of an `ncu --page source --csv --print-source sass` dump
per SASS opcode (one table per kernel).  usage: ncu_stalls.py <source.csv> [kernel-substring]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
rows = list(csv.reader(open(path)))
kernel, hdr = None, None
agg = {}
for r in rows:
    if r and r[0] == "Kernel Name":
        kernel = r[1]
        hdr = None
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if not hdr or not r or kernel is None or want not in kernel:
        continue
    d = dict(zip(hdr, r))
    src = d.get("Source", "").strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1] if len(src.split()) > 1 else op
    a = agg.setdefault(kernel, defaultdict(lambda: defaultdict(float)))
    a[op]["n"] += 1
    for k in ("# Samples", "Instructions Executed", "stall_dispatch", "stall_math", "stall_wait",
              "stall_long_sb", "stall_not_selected", "stall_selected", "stall_short_sb", "stall_lg",
              "stall_no_inst", "stall_branch_resolving", "stall_mio", "stall_barrier"):
        try:
            a[op][k] += float(d.get(k, 0) or 0)
        except ValueError:
            pass
for kernel, a in agg.items():
    tot = sum(v["# Samples"] for v in a.values())
    print("==", kernel[:100], " total samples", int(tot))
    keys = ["stall_dispatch", "stall_math", "stall_wait", "stall_long_sb", "stall_not_selected", "stall_selected",
            "stall_short_sb", "stall_mio", "stall_barrier", "stall_no_inst"]
    print("totals:", ", ".join("%s=%d" % (k, sum(v[k] for v in a.values())) for k in keys))
    print("%-22s %6s %12s %8s %6s  dispatch/math/wait/long_sb/not_sel/selected" % ("opcode", "#sass", "executed", "samples", "share"))
    for op, v in sorted(a.items(), key=lambda kv: -kv[1]["# Samples"])[:14]:
        print("%-22s %6d %12d %8d %5.1f%%  %d/%d/%d/%d/%d/%d" % (
            op, v["n"], v["Instructions Executed"], v["# Samples"], 100 * v["# Samples"] / max(tot, 1),
            v["stall_dispatch"], v["stall_math"], v["stall_wait"], v["stall_long_sb"],
            v["stall_not_selected"], v["stall_selected"]))
