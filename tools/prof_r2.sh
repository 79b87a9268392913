#!/bin/bash
# Round-2 profile captures (run on the GPU box under gpurun): launch list of the default
# bench command, ncu --set full of the MSM kernels (one 2^24 MSM as one window group and as
# the pipelined groups) and of the codec kernels at 2^20.  CSV pages only.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_msm24.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r2_launches_bench.log 2>&1
D377_MSM_GROUPS=1 tools/ncu_extract.sh r2_msm24_one_group "k_msm" 0 40 python tools/prof_msm.py 24 msm1
tools/ncu_extract.sh r2_msm24_pipelined "k_msm_accumulate|k_msm_normalize" 0 8 python tools/prof_msm.py 24 msm1
tools/ncu_extract.sh r2_codec20 "k_compress|k_decompress|k_elligator|k_hash_encode|k_fixed_base" 0 12 python tools/prof_msm.py 20 codec
ls -la gpurun_out/ | tail -20
