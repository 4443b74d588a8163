# round 2, run "g": lane-local interiors + deferred bin marking, genome-wide rank-line build, unconditional line loads
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r02g_pytest.log
legs() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null > gpurun_out/r02g_legs_$name.json
  python -c "
import json
d=json.load(open('gpurun_out/r02g_legs_$name.json'))
r=d['roofline']
print('$name', 'find', round(d['ms_per_step'],4), 'c4', round(r['c4_ms'],3), {k[:26]:v['ms'] for k,v in d['extra']['bed_intersect']['kernels_rank0'].items()}, 'c5', round(r['c5_ms'],3), {k[:18]:v['ms'] for k,v in d['extra']['aggregate']['kernels_rank0'].items()}, 'ok', r['c4_parity_ok'], r['c5_parity_ok'], r['bitset_and_parity_ok'])"
}
legs default
legs nobuckets BXB200_SET_BUCKETS=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^(k_set_ranges_multi|k_count_ranges_multi|k_aggregate_multi|k_rank_lines_multi|k_line_popc_multi|k_mark_touched_bins|k_range_buckets)$' -c 14 -o gpurun_out/r02g_prof_legs -f python bench.py --steps 2 --warmup 1 --no-cpu --c4-steps 1 --c5-steps 1 > gpurun_out/r02g_ncu_legs.log 2>&1; tail -1 gpurun_out/r02g_ncu_legs.log
