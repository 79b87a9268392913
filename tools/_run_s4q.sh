set +e
D377_MSM_GROUPS=1 timeout 900 bash tools/ncu_extract.sh r1c_msm24_g1 "k_msm|k_wsum|k_finish|k_scan" 16 16 python tools/prof_msm.py 24 msm
timeout 900 bash tools/ncu_extract.sh r1c_msm24_pipe "k_msm_accumulate|k_msm_normalize" 4 4 python tools/prof_msm.py 24 msm
timeout 900 bash tools/ncu_extract.sh r1c_codec20 "k_compress|k_decompress|k_elligator|k_fixed_base" 1 4 python tools/prof_msm.py 20 codec
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1c_launches_bench_msm24.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/s4q_ncu_bench.log 2>&1
ls -la gpurun_out/r1c_*
