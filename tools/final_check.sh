#!/bin/bash
# final validation of a build on one B200: full GPU suite, msm loops, default bench, smoke
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/msm_loop.py 20 21 22 24 2>&1 | grep -v stages
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2s_bench1.json 2> gpurun_out/r2s_bench1.err) 2>&1 | grep real
