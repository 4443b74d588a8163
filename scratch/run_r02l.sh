# round 2, run "l": speculative fill launch -- find tests + bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -k "find or forest or c2 or tree or neighb or small or offsets" 2>&1 | tail -4 ) 2>&1
( time timeout 600 python bench.py > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err ) 2>&1 | tail -4
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02l_bench.json'))
print('value', '%.4g' % d['value'], 'ms', round(d['ms_per_step'], 4), 'launches', d['gpu_launches'])
print('e2e', {k: v for k, v in d['e2e'].items() if k != 'timing'})
ro = d['roofline']
print({k: ro[k] for k in ('frac', 'avg_launch_ms', 'bitset_and_frac', 'c4_ms', 'c5_ms', 'c4_parity_ok', 'c5_parity_ok', 'bitset_and_parity_ok')})
print('kernels', d['extra']['kernels'], d['extra']['single_pass_kernel_ms_per_step'], d['extra']['sorted_queries_ms_per_step'])
PY
