set +e
(timeout 900 python -m pytest tests -m gpu -x -q -k "ellig or encode or config2 or canonical or host_api" 2>&1 | tail -4) > gpurun_out/s4p_tests.log; cat gpurun_out/s4p_tests.log
timeout 400 python bench.py --workload encode 2>&1 | tail -1 > gpurun_out/s4p_bench_encode.json
python -c "
import json; j=json.loads(open('gpurun_out/s4p_bench_encode.json').read()); print('encode', round(j['value'],2), j['unit'], 'ms', round(j['ms_per_step'],3), 'frac', round(j['roofline']['frac'],3), 'issued', round(j['roofline']['issued_frac'],3), 'e2e', round(j['e2e']['value'],2), j['verified_vs_oracle'])"
