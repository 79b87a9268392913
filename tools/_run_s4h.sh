set +e
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/s4h_tests.log; cat gpurun_out/s4h_tests.log
export D377_TIMELINE=1
run() { echo "== $*"; env "$@" timeout 200 python tools/tune_msm.py $L 2>&1 | grep "^n=\|group"; }
{
for L in 20 21 22 23 24; do run D377_X=0; done
L=24; run D377_MSM_GROUPS=2; run D377_MSM_GROUPS=4; run D377_MSM_GW=3,5,6 ; run D377_MSM_GW=2,5,7; run D377_MSM_GW=3,3,4,4
L=20; run D377_MSM_GW=8,8; run D377_MSM_GW=4,12
} > gpurun_out/s4h_tune.log 2>&1
sed -E 's/run=auto seg=auto: //; s/scan=0.00. scatter=0.00. //' gpurun_out/s4h_tune.log
for B in 0 32; do D377_CODEC_BLOCK=$B timeout 300 python tools/codec_balance.py; done 2>&1 | grep "^block" | tee gpurun_out/s4h_balance.log
