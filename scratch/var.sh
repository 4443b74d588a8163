timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "find or forest or tree or intersect or neighb or modes" 2>&1 | tail -2
for v in 0 1; do
  BXB200_FILL_STAGED=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-bitset --no-cpu 2>/dev/null > gpurun_out/fs_$v.json
  python -c "
import json
d=json.load(open('gpurun_out/fs_$v.json'))
print('STAGED=$v', d['ms_per_step'], {k:v['avg_ms'] for k,v in d['extra']['kernels'].items()}, d['extra']['single_pass_kernel_ms_per_step'], d['e2e']['value'])"
done
