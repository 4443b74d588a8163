# run "w": 2-GPU sanity of the bench contract (torchrun launch exactly as the driver does)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-bitset > gpurun_out/bench_w_n2.json 2> gpurun_out/bench_w_n2.err; tail -3 gpurun_out/bench_w_n2.err; python -c "
import json
d=json.load(open('gpurun_out/bench_w_n2.json'))
print('n_gpus',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['config'].get('n_queries'))
"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_w_ref_n2.json 2>/dev/null; head -c 300 gpurun_out/bench_w_ref_n2.json
