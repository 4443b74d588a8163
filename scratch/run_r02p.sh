# round 2, run "k": end-of-round validation -- all GPU tests, smoke, default bench + reference arm, launch list, ncu summaries
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/r02p_pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
( time timeout 600 python bench.py > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err ) 2>&1 | tail -4
tail -6 gpurun_out/r02p_bench.err
( time timeout 600 python bench.py --impl reference > gpurun_out/r02p_bench_reference.json 2> gpurun_out/r02p_ref.err ) 2>&1 | tail -4
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02p_bench.json'))
r = json.load(open('gpurun_out/r02p_bench_reference.json'))
print('value', '%.4g' % d['value'], 'ms', round(d['ms_per_step'], 4), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('e2e', {k: v for k, v in d['e2e'].items() if k != 'timing'})
print('reference', '%.4g' % r['value'], 'e2e ratio', d['e2e']['value'] / r['value'])
ro = d['roofline']
print({k: ro[k] for k in ro if not k.endswith(('workload', 'note', 'source'))})
print('kernels', d['extra']['kernels'])
for k in ('bed_intersect', 'aggregate'):
    print(k, d['extra'][k]['kernels_rank0'], d['extra'][k]['launches_per_step_rank0'])
print(d['extra']['scalar_api'])
print('cpu', d['cpu_baseline'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02p_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --c4-steps 1 --c5-steps 1 > gpurun_out/r02p_ncu_bench.log 2>&1; tail -c 300 gpurun_out/r02p_ncu_bench.log; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^(k_find|k_fill_staged)$' -s 8 -c 2 -o gpurun_out/r02p_prof_find -f python bench.py --steps 2 --warmup 1 --no-bitset --no-cpu > gpurun_out/r02p_ncu_find.log 2>&1; tail -1 gpurun_out/r02p_ncu_find.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^(k_binop_batch|k_set_ranges_bucketed|k_count_ranges_multi|k_aggregate_multi|k_group_stats|k_range_buckets)$' -s 4 -c 14 -o gpurun_out/r02p_prof_legs -f python bench.py --steps 2 --warmup 1 --no-cpu --c4-steps 1 --c5-steps 1 > gpurun_out/r02p_ncu_legs.log 2>&1; tail -1 gpurun_out/r02p_ncu_legs.log
