# run "q": probe search A/B, find tests, fresh ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "find or forest or tree or intersect or neighb or modes or small" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_scores.py -m gpu -q -x 2>&1 | tail -2
for v in 0 1; do
  BXB200_FIND_PROBE=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-bitset --no-cpu 2>/dev/null > gpurun_out/probe_$v.json
  python -c "
import json
d=json.load(open('gpurun_out/probe_$v.json'))
print('PROBE=$v', d['ms_per_step'], {k:v['avg_ms'] for k,v in d['extra']['kernels'].items()}, 'single', d['extra']['single_pass_kernel_ms_per_step'], 'sorted', d['extra']['sorted_queries_ms_per_step'], 'e2e', d['e2e']['value'], d['extra']['scalar_api'])"
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; python -c "
import json
d=json.load(open('gpurun_out/bench_q.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline'])
ss=d['extra'].get('score_sources') or {}
print({k:(v.get('ms')) for k,v in ss.items()})
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^(k_find|k_fill_staged|k_find_fused)$' -s 8 -c 6 -o gpurun_out/prof_find_q -f python bench.py --steps 2 --warmup 1 --no-bitset --no-cpu > gpurun_out/ncu_find_q.log 2>&1; tail -2 gpurun_out/ncu_find_q.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q.csv python bench.py --steps 2 --warmup 1 --no-bitset --no-cpu > gpurun_out/ncu_launch_q.log 2>&1
