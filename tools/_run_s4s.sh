set +e
(timeout 900 python -m pytest tests -m gpu -x -q -k "fixed_base or config3" 2>&1 | tail -4) > gpurun_out/s4s_tests.log; cat gpurun_out/s4s_tests.log
for Qv in 0 1; do
D377_FB_QUARTIC=$Qv timeout 400 python bench.py --workload fixed_base --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/s4s_bench_fb$Qv.json
python -c "
import json; j=json.loads(open('gpurun_out/s4s_bench_fb$Qv.json').read()); print('fixed_base quartic=$Qv', round(j['value'],2), j['unit'], 'ms', round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value'],2), j['verified_vs_oracle'])"
done
