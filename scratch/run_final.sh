# final validation of the round: all gpu tests, smoke, both bench arms, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_f.json 2> /dev/null
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; python -c "
import json
d=json.load(open('gpurun_out/bench_f.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],d['roofline']['kernel'], d['roofline']['traffic'])
print('bitset', d['roofline'].get('bitset_and'))
print(json.dumps(d['extra']['kernels']))
print('clocks',d['clocks'], 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline']['value'])
r=json.load(open('gpurun_out/bench_ref_f.json')); print('reference arm', r['value'], r['cpu_baseline']['cores'])
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_f.csv python bench.py --steps 2 --warmup 1 --no-bitset --no-cpu > gpurun_out/ncu_launch_f.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^(k_find|k_fill_staged|k_find_fused)$' -s 8 -c 6 -o gpurun_out/prof_find_f -f python bench.py --steps 2 --warmup 1 --no-bitset --no-cpu > gpurun_out/ncu_find_f.log 2>&1; tail -1 gpurun_out/ncu_find_f.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^(k_binop_batch|k_set_ranges_multi|k_spans_write|k_summarize|k_join)$' -c 8 -o gpurun_out/prof_bits_f -f python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bits_f.log 2>&1; tail -1 gpurun_out/ncu_bits_f.log
