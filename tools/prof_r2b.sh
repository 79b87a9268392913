#!/bin/bash
# Round-2 final captures of the kernels that changed after tools/prof_r2.sh ran: the fused
# hash_to_curve kernel, the config-1 scalar multiplication, and the TMA-staged accumulation.
set -x
tools/ncu_extract.sh r2b_codec20 "k_hash_encode|k_elligator_encode|k_scalar_mul" 0 6 python tools/prof_msm.py 20 codec2
D377_MSM_GROUPS=1 D377_MSM_ACC_TMA=1 tools/ncu_extract.sh r2b_msm24_tma "k_msm_accumulate" 0 1 python tools/prof_msm.py 24 msm1
ls -la gpurun_out | grep r2b
