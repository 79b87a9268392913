set +e
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/s12_tests.log; cat gpurun_out/s12_tests.log
timeout 400 python bench.py 2>&1 | tail -1 > gpurun_out/s12_bench.log
D377_MSM_GROUPS=1 timeout 400 python bench.py --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/s12_bench_g1.log
./tools/ub_field > gpurun_out/s12_ub_field.txt 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/s12_bench*.log")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["value"],1), round(j["ms_per_step"],3), j["msm_stage_ms"], j["roofline"])
        for k in ("e2e","e2e_element","e2e_sync","e2e_affine","cpu_baseline"): print("  ",k,j.get(k))
    except Exception as e: print(f, "ERR", open(f).read()[-1500:])
PY
cat gpurun_out/s12_ub_field.txt
