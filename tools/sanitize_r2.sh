#!/bin/bash
# compute-sanitizer over the round-2 paths (run on the GPU box); summary lines only.
set -x
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -m gpu -x -q \
  -k "async or multi or montgomery or 64_byte or 80_bytes or sub_neg or on_curve or noncanonical or registry or threads" \
  > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q \
  -k "64_byte or encode_and_hash or msm_async or montgomery or element_sum" \
  > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2_racecheck.log
tail -6 gpurun_out/r2_memcheck.log gpurun_out/r2_racecheck.log
