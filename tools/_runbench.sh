set +e
(timeout 900 python -m pytest tests -m gpu -x -q -k "msm or pipeline or chunk" 2>&1 | tail -5) > gpurun_out/s11_tests.log; cat gpurun_out/s11_tests.log
timeout 400 python bench.py --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/s11_bench.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/s11_bench*.log")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["value"],1), round(j["ms_per_step"],3), j["msm_stage_ms"], round(j["roofline"]["frac"],3))
        for k in ("e2e","e2e_element","e2e_sync","e2e_affine"): print("  ",k,j.get(k))
    except Exception as e: print(f, "ERR", open(f).read()[-1500:])
PY
