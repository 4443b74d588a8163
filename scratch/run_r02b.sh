# round 2, run "b": new tests, grid search A/B, ncu captures
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) 2>&1 | tee gpurun_out/r02b_pytest.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-bitset --no-cpu 2>/dev/null > gpurun_out/r02b_$name.json
  python -c "
import json
d=json.load(open('gpurun_out/r02b_$name.json'))
print('$name', 'ms', round(d['ms_per_step'],4), {k:v['avg_ms'] for k,v in d['extra']['kernels'].items()}, 'single', round(d['extra']['single_pass_kernel_ms_per_step'],4), 'sorted', round(d['extra']['sorted_queries_ms_per_step'],4), 'e2e', '%.3g'%d['e2e']['value'], 'count_only', '%.3g'%d['e2e']['count_only_value'], 'build', round(d['extra']['build_ms'],2), 'scalar', d['extra']['scalar_api']['find_us_per_call'], d['extra']['scalar_api']['count_range_us_per_call'])"
}
run grid1 BXB200_FIND_PROBE=3
run probe8 BXB200_FIND_PROBE=2
run grid0 BXB200_FIND_PROBE=3 BXB200_GRID_LOG2=0
run grid2 BXB200_FIND_PROBE=3 BXB200_GRID_LOG2=2
( time timeout 600 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err ) 2>&1 | tail -4
tail -5 gpurun_out/r02b_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02b_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['count_only_value'])
r = d['roofline']
print({k: r[k] for k in r if k.startswith(('c4_', 'c5_', 'bitset_and_')) and not k.endswith('workload')})
for k in ('bed_intersect', 'aggregate'):
    print(k, d['extra'][k]['kernels_rank0'])
print(d['extra']['scalar_api'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^(k_find|k_fill_staged)$' -s 8 -c 4 -o gpurun_out/r02b_prof_find -f python bench.py --steps 2 --warmup 1 --no-bitset --no-cpu > gpurun_out/r02b_ncu_find.log 2>&1; tail -1 gpurun_out/r02b_ncu_find.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^(k_set_ranges_multi|k_count_ranges_multi|k_aggregate_multi|k_group_stats)$' -c 6 -o gpurun_out/r02b_prof_legs -f python bench.py --steps 2 --warmup 1 --no-cpu --c4-steps 1 --c5-steps 1 > gpurun_out/r02b_ncu_legs.log 2>&1; tail -1 gpurun_out/r02b_ncu_legs.log
ls -la gpurun_out/*.ncu-rep
