set +e
(timeout 900 python -m pytest tests -m gpu -x -q -k "msm" 2>&1 | tail -4) > gpurun_out/s4k_tests.log; cat gpurun_out/s4k_tests.log
run() { echo "== $*"; env "$@" timeout 200 python tools/tune_msm.py $L 2>&1 | grep "^n=\|group"; }
{
for L in 24 22 20; do run D377_ACC_REGPIPE=0; run D377_ACC_REGPIPE=1; done
L=24; run D377_ACC_REGPIPE=0 D377_MSM_GROUPS=1; run D377_ACC_REGPIPE=1 D377_MSM_GROUPS=1
} > gpurun_out/s4k_tune.log 2>&1
sed -E 's/run=auto seg=auto: //; s/scan=0.00. scatter=0.00. //' gpurun_out/s4k_tune.log
