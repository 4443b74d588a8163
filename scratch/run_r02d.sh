# round 2, run "d" (2 GPUs): the multi-rank bench legs with NCCL, then single-GPU A/Bs on GPU 0
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err ) 2>&1 | tail -3
grep -v "^W\|^\*" gpurun_out/r02d_bench_n2.err | tail -8
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02d_bench_n2.json'))
print('N=2 value', '%.4g' % d['value'], 'ms', d['ms_per_step'], 'n_gpus', d['n_gpus'])
print('e2e', {k: v for k, v in d['e2e'].items() if k != 'timing'})
r = d['roofline']
print({k: r[k] for k in r if k.startswith(('c4_', 'c5_', 'bitset_and_')) and not k.endswith('workload')})
print('probe', d['extra']['copy_probe'])
print('strong', d['extra']['strong_scaling'])
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02d_ref_n2.json 2> gpurun_out/r02d_ref_n2.err ) 2>&1 | tail -3
cut -c1-400 gpurun_out/r02d_ref_n2.json
echo
legs() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null > gpurun_out/r02d_legs_$name.json
  python -c "
import json
d=json.load(open('gpurun_out/r02d_legs_$name.json'))
r=d['roofline']
print('$name', 'find', round(d['ms_per_step'],4), {k[:14]:v['avg_ms'] for k,v in d['extra']['kernels'].items()}, 'c4', round(r['c4_ms'],3), {k[:18]:v['ms'] for k,v in d['extra']['bed_intersect']['kernels_rank0'].items()}, 'c5', round(r['c5_ms'],3), 'ok', r['c4_parity_ok'], r['c5_parity_ok'], 'scalar', d['extra']['scalar_api'])"
}
legs hint1
BXB200_NVCC_FLAGS="-DCOUNT_L2_HINT=0 -DBXS_WALK_PREFETCH" python -m bx_python_b200.build > /dev/null 2>&1
legs hint0_prefetch
