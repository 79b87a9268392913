set +e
(timeout 900 python -m pytest tests -m gpu -x -q -k "msm or element_sum" 2>&1 | tail -6) > gpurun_out/s4d_tests.log; cat gpurun_out/s4d_tests.log
run() { echo "== $*"; env "$@" timeout 200 python tools/tune_msm.py $L 2>&1 | grep "^n="; }
{
for L in 20 24; do
  run D377_MSM_NORM_WAVE=0 D377_MSM_WTREE=0 D377_MSM_STITCH_WARP=0
  run D377_MSM_NORM_WAVE=0
  run D377_MSM_NORM_WAVE=3
  run D377_MSM_NORM_WAVE=2
  run D377_MSM_STITCH_WARP=4096
  run D377_MSM_STITCH_WARP=2000000
done
L=22; run D377_MSM_NORM_WAVE=0; run D377_MSM_NORM_WAVE=3; run D377_MSM_NORM_WAVE=2
for L in 21 22 23; do for G in 2 3 4; do run D377_MSM_GROUPS=$G; done; done
L=26; run D377_MSM_NORM_WAVE=0; run D377_MSM_NORM_WAVE=3
} > gpurun_out/s4d_tune.log 2>&1
cat gpurun_out/s4d_tune.log
