mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^(k_set_ranges_bucketed|k_aggregate_multi|k_count_ranges_multi|k_group_stats|k_range_buckets|k_mark_touched_bins|k_rank_lines_multi)$' -c 9 -o gpurun_out/r02q_prof_legs -f python bench.py --steps 2 --warmup 1 --no-cpu --c4-steps 1 --c5-steps 1 > gpurun_out/r02q_ncu_legs.log 2>&1; tail -1 gpurun_out/r02q_ncu_legs.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null > gpurun_out/r02q_legs.json
python -c "
import json
d=json.load(open('gpurun_out/r02q_legs.json'))
r=d['roofline']
print('c4', round(r['c4_ms'],3), {k[:26]:v['ms'] for k,v in d['extra']['bed_intersect']['kernels_rank0'].items()}, 'ok', r['c4_parity_ok'])"
