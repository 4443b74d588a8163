# run "g": fused kernel with staged emission
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scores.py -m gpu -q -x -k "find or forest or tree or intersect or modes or small or join or aggregate_nonnan" 2>&1 | tail -3
run() {
  timeout 300 python bench.py --steps 10 --warmup 3 --no-bitset --no-cpu 2>/dev/null > gpurun_out/g_$1.json
  python -c "
import json
d=json.load(open('gpurun_out/g_$1.json'))
print('$1', round(d['ms_per_step'],4), {k:v['avg_ms'] for k,v in d['extra']['kernels'].items()}, 'single', round(d['extra']['single_pass_kernel_ms_per_step'],4), 'sorted', round(d['extra']['sorted_queries_ms_per_step'],4), 'e2e', d['e2e']['value'], 'i64', d['extra']['e2e_int64_offsets'])"
}
run default
