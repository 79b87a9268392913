#!/bin/bash
# Window-size sweep in the pipelined regime (tails hidden): the round-1 choice was tuned on the
# blocking call.  Output: gpurun_out/r2_sweep_c.log
mkdir -p gpurun_out
for spec in "20 15 16 17" "21 16 17 18" "22 17 18 19" "23 18 19 21" "24 18 19 20 21"; do
  set -- $spec; logn=$1; shift
  for c in "$@"; do
    echo "== 2^$logn c=$c"; timeout 120 python tools/msm_loop.py --fast -c=$c $logn 2>&1 | grep "^msm"
  done
done > gpurun_out/r2_sweep_c.log 2>&1
cat gpurun_out/r2_sweep_c.log
