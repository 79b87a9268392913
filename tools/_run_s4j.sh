set +e
timeout 600 python bench.py 2>gpurun_out/s4j_bench_err.log | tail -1 > gpurun_out/s4j_bench.json
python -c "
import json; j=json.loads(open('gpurun_out/s4j_bench.json').read()); print('msm24', round(j['value'],1), 'ms', round(j['ms_per_step'],3), 'frac', round(j['roofline']['frac'],3), 'e2e', round(j['e2e']['value'],1), j['msm_stage_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1b_launches_bench_msm24.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/s4j_ncu_bench.log 2>&1
D377_MSM_GROUPS=1 timeout 900 bash tools/ncu_extract.sh r1b_msm24_g1 "k_msm|k_wsum|k_finish|k_scan" 16 16 python tools/prof_msm.py 24 msm
timeout 900 bash tools/ncu_extract.sh r1b_msm24_pipe "k_msm_accumulate|k_msm_normalize" 4 4 python tools/prof_msm.py 24 msm
timeout 900 bash tools/ncu_extract.sh r1b_codec20 "k_compress|k_decompress|k_elligator|k_fixed_base" 1 4 python tools/prof_msm.py 20 codec
rm -f gpurun_out/*_source.csv.tmp
ls -la gpurun_out/ | tail -20
