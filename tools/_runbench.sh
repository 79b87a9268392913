set +e
(timeout 900 python -m pytest tests -m gpu -x -q -k "msm" 2>&1 | tail -5) > gpurun_out/s14_tests.log; cat gpurun_out/s14_tests.log
for G in 0 1 3; do D377_MSM_GROUPS=$G timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 6 2>&1 | tail -1 > gpurun_out/s14_bench_g$G.log; done
timeout 300 python bench.py --workload msm --logn 20 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/s14_bench_msm20.log
timeout 300 python bench.py --workload msm --logn 22 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/s14_bench_msm22.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/s14_bench*.log")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["value"],1), round(j["ms_per_step"],3), j["msm_stage_ms"], round(j["roofline"]["frac"],3))
    except Exception as e: print(f, "ERR", open(f).read()[-300:])
PY
