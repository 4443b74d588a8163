# round 2: the bench under torchrun exactly as the driver launches it, N = $1 ranks (tight timeout: a hang must not eat the budget)
N=${1:-2}
mkdir -p gpurun_out
( time timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02n_bench_n$N.json 2> gpurun_out/r02n_bench_n$N.err ) 2>&1 | tail -3
strings gpurun_out/r02n_bench_n$N.err | grep "^\[bench\]" | tail -6
python - <<PY
import json
d = json.load(open('gpurun_out/r02n_bench_n$N.json'))
print('N', d['n_gpus'], 'value', '%.4g' % d['value'], 'ms', round(d['ms_per_step'], 4))
print('e2e', {k: (round(v, 3) if isinstance(v, float) and v < 1e6 else v) for k, v in d['e2e'].items() if k != 'timing'})
r = d['roofline']
print({k: r[k] for k in r if k.startswith(('c4_', 'c5_', 'bitset_and_')) and not k.endswith('workload')})
print('probe', d['extra']['copy_probe'])
PY
( time timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02n_ref_n$N.json 2> /dev/null ) 2>&1 | tail -3
cut -c1-200 gpurun_out/r02n_ref_n$N.json
