# run "r": L2 hints A/B (default build has hints on), tests on the default build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "find or forest or tree or intersect or neighb or modes or small" 2>&1 | tail -3
run() {
  timeout 300 python bench.py --steps 10 --warmup 3 --no-bitset --no-cpu 2>/dev/null > gpurun_out/l2_$1.json
  python -c "
import json
d=json.load(open('gpurun_out/l2_$1.json'))
print('$1', d['ms_per_step'], {k:v['avg_ms'] for k,v in d['extra']['kernels'].items()}, 'single', d['extra']['single_pass_kernel_ms_per_step'], 'sorted', d['extra']['sorted_queries_ms_per_step'], 'e2e', d['e2e']['value'])"
}
run hints1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^(k_find|k_fill_staged)$' -s 8 -c 4 -o gpurun_out/prof_find_r -f python bench.py --steps 2 --warmup 1 --no-bitset --no-cpu > gpurun_out/ncu_find_r.log 2>&1; tail -1 gpurun_out/ncu_find_r.log
BXB200_NVCC_FLAGS="-DFIND_L2_HINTS=0" python -m bx_python_b200.build > /dev/null 2>&1
run hints0
