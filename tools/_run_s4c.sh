set +e
(timeout 900 python -m pytest tests -m gpu -x -q -k "msm or element_sum" 2>&1 | tail -6) > gpurun_out/s4c_tests.log; cat gpurun_out/s4c_tests.log
for L in 20 22 24; do
 for SW in 0 1; do for WT in 0 1; do
  echo "== logn=$L stitch_warp=$SW wtree=$WT"
  D377_MSM_STITCH_WARP=$SW D377_MSM_WTREE=$WT timeout 200 python tools/tune_msm.py $L 2>&1 | grep "^n="
 done; done
done > gpurun_out/s4c_tune.log 2>&1
for SEG in 4 8 32; do echo "== logn=20/24 wtree seg=$SEG"; D377_REDUCE_SEG=$SEG timeout 200 python tools/tune_msm.py 20 2>&1 | grep "^n="; D377_REDUCE_SEG=$SEG timeout 200 python tools/tune_msm.py 24 2>&1 | grep "^n="; done >> gpurun_out/s4c_tune.log 2>&1
for G in 2 3; do echo "== logn=20/21 groups=$G"; D377_MSM_GROUPS=$G timeout 200 python tools/tune_msm.py 20 2>&1 | grep "^n=";  D377_MSM_GROUPS=$G timeout 200 python tools/tune_msm.py 21 2>&1 | grep "^n="; done >> gpurun_out/s4c_tune.log 2>&1
for G in 0 3; do echo "== logn=22 groups=$G"; D377_MSM_GROUPS=$G timeout 200 python tools/tune_msm.py 22 2>&1 | grep "^n="; done >> gpurun_out/s4c_tune.log 2>&1
cat gpurun_out/s4c_tune.log
