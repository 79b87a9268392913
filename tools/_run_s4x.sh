set +e
export PATH=/usr/local/cuda/bin:$PATH
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm_matches_oracle or msm_window_group or fixed_base_matches or encode_and_hash or normalize_batch or msm_edge or msm_timeline" > gpurun_out/s4x_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -15 gpurun_out/s4x_memcheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm_matches_oracle or encode_and_hash or element_sum" > gpurun_out/s4x_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -15 gpurun_out/s4x_racecheck.log | cut -c1-200
