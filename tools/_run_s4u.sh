set +e
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/s4u_tests.log; cat gpurun_out/s4u_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py 2>gpurun_out/s4u_bench_err.log | tail -1 > gpurun_out/s4u_bench_msm24.json
for L in 20 22; do timeout 400 python bench.py --workload msm --logn $L --no-cpu-baseline --steps 6 2>&1 | tail -1 > gpurun_out/s4u_bench_msm$L.json; done
for WL in compress decompress encode hash fixed_base pipeline; do timeout 400 python bench.py --workload $WL 2>&1 | tail -1 > gpurun_out/s4u_bench_$WL.json; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s4u_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]; cb=j.get("cpu_baseline") or {}
        print(f.split("s4u_bench_")[1], round(j["value"],1), j["unit"], "ms", round(j["ms_per_step"],3), "frac", round(r["frac"],3), "issued", round(r.get("issued_frac") or 0,3),
              "e2e", round(j["e2e"]["value"],1), "cpu", round(cb.get("value",0),3), cb.get("cores"), j.get("verified_vs_oracle"))
    except Exception as e: print(f, "ERR", e, open(f).read()[-300:])
PY
timeout 900 bash tools/ncu_extract.sh r1d_codec20 "k_compress|k_decompress|k_elligator_encode|k_fixed_base_jq|k_hash_encode" 1 6 python tools/prof_msm.py 20 codec
