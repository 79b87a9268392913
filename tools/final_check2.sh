#!/bin/bash
# final validation of the round-2 tree on one B200: GPU suite, smoke, default bench, memcheck of
# the newest kernels, fresh launch list of the MSM bench command
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2x_bench1.json 2> gpurun_out/r2x_bench1.err) 2>&1 | grep real
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "affine" > gpurun_out/r2x_memcheck_affine.log 2>&1; tail -3 gpurun_out/r2x_memcheck_affine.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2x_launches_bench_msm24.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r2x_launches_bench.log 2>&1
wc -l gpurun_out/r2x_launches_bench_msm24.csv
