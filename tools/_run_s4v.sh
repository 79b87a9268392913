set +e
(timeout 900 python -m pytest tests -m gpu -x -q -k "ellig or encode or hash or config2 or host_api or compress" 2>&1 | tail -3) > gpurun_out/s4v_tests.log; cat gpurun_out/s4v_tests.log
for WL in encode compress decompress; do timeout 400 python bench.py --workload $WL --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/s4v_bench_$WL.json
python -c "
import json; j=json.loads(open('gpurun_out/s4v_bench_$WL.json').read()); print('$WL', round(j['value'],2), 'ms', round(j['ms_per_step'],3), 'frac', round(j['roofline']['frac'],3), 'issued', round(j['roofline']['issued_frac'],3), 'e2e', round(j['e2e']['value'],2), j['verified_vs_oracle'])"
done
