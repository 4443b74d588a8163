mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-bitset > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -2 gpurun_out/bench_n8.err | cut -c1-300; python -c "
import json
d=json.load(open('gpurun_out/bench_n8.json'))
print('n_gpus',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'i64',d['extra']['e2e_int64_offsets'])
"
