set +e
(timeout 1500 python -m pytest tests -m gpu -x -q -k "msm or element_sum" 2>&1 | tail -6) > gpurun_out/s4f_tests.log; cat gpurun_out/s4f_tests.log
run() { echo "== $*"; env "$@" timeout 200 python tools/tune_msm.py $L 2>&1 | grep "^n="; }
{
for L in 20 21 22 23 24; do
  run D377_X=0
  run D377_MSM_GROUPS=2
  run D377_MSM_GROUPS=3
  run D377_MSM_GROUPS=4
done
L=26; run D377_X=0
} > gpurun_out/s4f_tune.log 2>&1
sed -E 's/run=auto seg=auto: //; s/scan=0.00. scatter=0.00. //' gpurun_out/s4f_tune.log
