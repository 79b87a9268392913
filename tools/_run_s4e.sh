set +e
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/s4e_tests.log; cat gpurun_out/s4e_tests.log
run() { echo "== $*"; env "$@" timeout 200 python tools/tune_msm.py $L 2>&1 | grep "^n="; }
{
for L in 21 22 24; do
  run D377_GCD_INV=0
  run D377_GCD_INV=1
  run D377_GCD_INV=1 D377_MSM_NORM_WAVE=2
done
L=24; run D377_MSM_GROUPS=2; run D377_MSM_GROUPS=3; run D377_MSM_GROUPS=5
} > gpurun_out/s4e_tune.log 2>&1
cat gpurun_out/s4e_tune.log
for B in 128 32 64; do for WL in pipeline compress; do
D377_CODEC_BLOCK=$B timeout 300 python bench.py --workload $WL --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; j=json.loads(sys.stdin.read()); print('$WL block=$B', round(j['value'],2), j['unit'], 'ms', round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value'],2))"
done; done 2>&1 | tee gpurun_out/s4e_codec.log
