set +e
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/s4a_tests.log; cat gpurun_out/s4a_tests.log
timeout 600 python bench.py 2>gpurun_out/s4a_bench_err.log | tail -1 > gpurun_out/s4a_bench.json; cat gpurun_out/s4a_bench.json
timeout 300 python bench.py --workload msm --logn 20 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/s4a_bench_msm20.json; cat gpurun_out/s4a_bench_msm20.json
