set +e
(timeout 900 python -m pytest tests -m gpu -x -q -k "timeline or normalize_batch" 2>&1 | tail -4) > gpurun_out/s4n_tests.log; cat gpurun_out/s4n_tests.log
{
timeout 300 python tools/tune_msm.py 20 14 15 16 17 | grep "^n="
timeout 300 python tools/tune_msm.py 21 15 16 17 18 | grep "^n="
timeout 300 python tools/tune_msm.py 22 16 17 18 19 | grep "^n="
timeout 300 python tools/tune_msm.py 23 17 18 19 20 | grep "^n="
timeout 300 python tools/tune_msm.py 24 17 18 19 20 21 | grep "^n="
timeout 300 python tools/tune_msm.py 25 18 19 20 21 | grep "^n="
timeout 400 python tools/tune_msm.py 26 18 20 21 22 | grep "^n="
} > gpurun_out/s4n_tune.log 2>&1
sed -E 's/run=auto seg=auto: //; s/scan=0.00. scatter=0.00. //' gpurun_out/s4n_tune.log
