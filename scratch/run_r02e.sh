# round 2, run "e" (2 GPUs): the multi-rank bench with NCCL (tight timeout)
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err ) 2>&1 | tail -3
strings gpurun_out/r02e_bench_n2.err | grep "^\[bench\]" | tail -8
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02e_bench_n2.json'))
print('N=2 value', '%.4g' % d['value'], 'ms', d['ms_per_step'], 'n_gpus', d['n_gpus'])
print('e2e', {k: v for k, v in d['e2e'].items() if k != 'timing'})
r = d['roofline']
print({k: r[k] for k in r if k.startswith(('c4_', 'c5_', 'bitset_and_')) and not k.endswith('workload')})
print('probe', d['extra']['copy_probe'])
print('strong', d['extra']['strong_scaling'])
PY
