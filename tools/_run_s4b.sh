set +e
for L in 20 22; do
 for G in 1 2; do for NM in 0 1; do
  echo "== logn=$L groups=$G norm=$NM"
  D377_MSM_GROUPS=$G D377_MSM_NORMALIZE=$NM timeout 200 python tools/tune_msm.py $L 15 16 17 18 2>&1 | grep "^n="
 done; done
done > gpurun_out/s4b_tune.log 2>&1
timeout 400 python bench.py --workload msm --logn 26 --no-cpu-baseline --steps 3 2>&1 | tail -1 > gpurun_out/s4b_bench_msm26.json
cat gpurun_out/s4b_tune.log; cut -c1-600 gpurun_out/s4b_bench_msm26.json
