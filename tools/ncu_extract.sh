#!/bin/bash
# Run on the GPU box: capture `ncu --set full` for a kernel regex, keep only the raw-metric
# CSV page and the per-opcode stall summary of the source page (tools/ncu_stalls.py), so
# that gpurun_out/ stays far below its 64 MiB limit.
#   tools/ncu_extract.sh <tag> <kernel-regex> <skip> <count> <cmd...>
set -e
tag=$1; regex=$2; skip=$3; count=$4; shift 4
rep=/tmp/$tag
ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c "$count" -f -o "$rep" "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
ncu -i $rep.ncu-rep --page source --csv --print-source sass > /tmp/${tag}_source.csv 2>/dev/null || true
python tools/ncu_stalls.py /tmp/${tag}_source.csv > gpurun_out/${tag}_stalls.txt 2>/dev/null || true
ls -la gpurun_out/${tag}_*
