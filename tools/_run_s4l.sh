set +e
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline 2>gpurun_out/s4l_err.log | tail -1 > gpurun_out/s4l_bench_n2.json
python -c "
import json; j=json.loads(open('gpurun_out/s4l_bench_n2.json').read()); print('N=2', round(j['value'],1), j['unit'], 'ms', round(j['ms_per_step'],3), 'e2e', j.get('e2e',{}).get('value'), j['config'])"
tail -5 gpurun_out/s4l_err.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>&1 | tail -2 | cut -c1-300
