set +e
nvidia-smi -L | wc -l
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline 2>gpurun_out/s4m_err_n$N.log | tail -1 > gpurun_out/s4m_bench_n$N.json
python -c "
import json; j=json.loads(open('gpurun_out/s4m_bench_n$N.json').read()); print('N=$N', round(j['value'],1), j['unit'], 'ms', round(j['ms_per_step'],3), 'e2e', j.get('e2e',{}).get('value'), j['config']['parallelism'], j['clocks'])"
done
