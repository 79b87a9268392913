set +e
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/s4o_tests.log; cat gpurun_out/s4o_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py 2>gpurun_out/s4o_bench_err.log | tail -1 > gpurun_out/s4o_bench_msm24.json
for L in 20 22 25 26; do timeout 400 python bench.py --workload msm --logn $L --no-cpu-baseline --steps 6 2>&1 | tail -1 > gpurun_out/s4o_bench_msm$L.json; done
for WL in compress decompress encode fixed_base pipeline; do timeout 400 python bench.py --workload $WL 2>&1 | tail -1 > gpurun_out/s4o_bench_$WL.json; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s4o_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]; cb=j.get("cpu_baseline") or {}
        print(f.split("s4o_bench_")[1], round(j["value"],1), j["unit"], "ms", round(j["ms_per_step"],3), "frac", round(r["frac"],3), "issued", round(r.get("issued_frac") or 0,3),
              "e2e", round(j["e2e"]["value"],1), "e2e_el", round((j.get("e2e_element") or {}).get("value",0),1), "e2e_aff", round((j.get("e2e_affine") or {}).get("value",0),1), "cpu", round(cb.get("value",0),3), cb.get("cores"), j.get("msm_stage_ms"))
    except Exception as e: print(f, "ERR", e, open(f).read()[-300:])
PY
