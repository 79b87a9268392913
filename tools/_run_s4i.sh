set +e
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/s4i_tests.log; cat gpurun_out/s4i_tests.log
run() { echo "== $*"; env "$@" timeout 200 python tools/tune_msm.py $L 2>&1 | grep "^n=\|group"; }
{
for L in 20 22 24; do run D377_X=0; done
} > gpurun_out/s4i_tune.log 2>&1
sed -E 's/run=auto seg=auto: //; s/scan=0.00. scatter=0.00. //' gpurun_out/s4i_tune.log
for WL in compress decompress encode fixed_base pipeline; do
timeout 300 python bench.py --workload $WL --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/s4i_bench_$WL.json
python -c "
import sys,json; j=json.loads(open('gpurun_out/s4i_bench_$WL.json').read()); print('$WL', round(j['value'],2), j['unit'], 'ms', round(j['ms_per_step'],3), 'frac', round(j['roofline']['frac'],3), 'e2e', round(j['e2e']['value'],2))"
done 2>&1 | tee gpurun_out/s4i_codec.log
