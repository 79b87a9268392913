#!/bin/bash
# configuration 1 (2^16 decompress -> mul -> compress): CTA size / residency sweep
mkdir -p gpurun_out
run() { echo "== block=$1 smem=$2"; D377_SM_BLOCK=$1 D377_SM_SMEM=$2 timeout 200 python bench.py --workload pipeline --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
l = json.loads([x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1])
print(round(l['value'], 2), 'Melem/s', round(l['ms_per_step'], 3), 'ms', l.get('verified_vs_oracle'))"; }
run 128 0
run 64 0
run 64 29696
run 96 0
run 32 14848
