set +e
for w in encode fixed_base pipeline decompress compress; do timeout 300 python bench.py --workload $w 2>&1 | tail -1 > gpurun_out/s13_bench_$w.log; done
timeout 300 python bench.py --workload msm --logn 20 2>&1 | tail -1 > gpurun_out/s13_bench_msm20.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/s13_bench*.log")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1], round(j["value"],2), j["unit"], round(j["ms_per_step"],3), "frac", round(j["roofline"]["frac"],3), "issued", j["roofline"].get("issued_frac"), "e2e", round(j["e2e"]["value"],2), "cpu", round(j["cpu_baseline"]["value"],3), j["cpu_baseline"]["cores"])
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
PY
