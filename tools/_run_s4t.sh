set +e
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/s4t_tests.log; cat gpurun_out/s4t_tests.log
for WL in encode hash fixed_base; do
timeout 400 python bench.py --workload $WL --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/s4t_bench_$WL.json
python -c "
import json; j=json.loads(open('gpurun_out/s4t_bench_$WL.json').read()); print('$WL', round(j['value'],2), j['unit'], 'ms', round(j['ms_per_step'],3), 'frac', round(j['roofline']['frac'],3), 'e2e', round(j['e2e']['value'],2), j['verified_vs_oracle'])"
done
