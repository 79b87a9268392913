set +e
for N in 8 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline 2>gpurun_out/s4w_err_n$N.log | tail -1 > gpurun_out/s4w_bench_n$N.json
python -c "
import json; j=json.loads(open('gpurun_out/s4w_bench_n$N.json').read()); print('N=$N', round(j['value'],1), j['unit'], 'ms', round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value'],1), 'strong', j.get('strong_scaling_2p24'))"
tail -3 gpurun_out/s4w_err_n$N.log | cut -c1-200
done
