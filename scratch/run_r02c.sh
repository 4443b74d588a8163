# round 2, run "c": full tests, PROBE 3/4, L2 fetch granularity, occupancy, ncu
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) 2>&1 | tee gpurun_out/r02c_pytest.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-bitset --no-cpu 2>/dev/null > gpurun_out/r02c_$name.json
  python -c "
import json
d=json.load(open('gpurun_out/r02c_$name.json'))
print('$name', 'ms', round(d['ms_per_step'],4), {k[:14]:v['avg_ms'] for k,v in d['extra']['kernels'].items()}, 'single', round(d['extra']['single_pass_kernel_ms_per_step'],4), 'sorted', round(d['extra']['sorted_queries_ms_per_step'],4), 'e2e', '%.3g'%d['e2e']['value'], 'cnt', '%.3g'%d['e2e']['count_only_value'], 'build', round(d['extra']['build_ms'],2), 'scalar', round(d['extra']['scalar_api']['find_us_per_call'],1), round(d['extra']['scalar_api']['count_range_us_per_call'],1))"
}
run p4_f32 BXB200_FIND_PROBE=4
run p3_f32 BXB200_FIND_PROBE=3
run p4_f64 BXB200_FIND_PROBE=4 BXB200_L2_FETCH=64
run p4_g2 BXB200_FIND_PROBE=4 BXB200_GRID_LOG2=2
run p4_g0 BXB200_FIND_PROBE=4 BXB200_GRID_LOG2=0
legs() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null > gpurun_out/r02c_legs_$name.json
  python -c "
import json
d=json.load(open('gpurun_out/r02c_legs_$name.json'))
r=d['roofline']
print('$name', 'and', round(r['bitset_and_gbs']), round(r['bitset_and_count_gbs']), round(r['bitset_and_per_pair_gbs']), 'c4', round(r['c4_ms'],3), {k[:18]:v['ms'] for k,v in d['extra']['bed_intersect']['kernels_rank0'].items()}, 'c5', round(r['c5_ms'],3), {k[:18]:v['ms'] for k,v in d['extra']['aggregate']['kernels_rank0'].items()}, 'ok', r['c4_parity_ok'], r['c5_parity_ok'], r['bitset_and_parity_ok'])"
}
legs f32
legs f64 BXB200_L2_FETCH=64
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^(k_find|k_fill_staged)$' -s 8 -c 2 -o gpurun_out/r02c_prof_find -f python bench.py --steps 2 --warmup 1 --no-bitset --no-cpu > gpurun_out/r02c_ncu_find.log 2>&1; tail -1 gpurun_out/r02c_ncu_find.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^(k_aggregate_multi|k_count_ranges_multi)$' -c 2 -o gpurun_out/r02c_prof_legs -f python bench.py --steps 2 --warmup 1 --no-cpu --c4-steps 1 --c5-steps 1 > gpurun_out/r02c_ncu_legs.log 2>&1; tail -1 gpurun_out/r02c_ncu_legs.log
BXB200_NVCC_FLAGS="-DFIND_MIN_CTAS=8 -DFILL_MIN_CTAS=8" python -m bx_python_b200.build > /dev/null 2>&1
run p4_occ8 BXB200_FIND_PROBE=4
BXB200_NVCC_FLAGS="-DFIND_MIN_CTAS=7" python -m bx_python_b200.build > /dev/null 2>&1
run p4_occ7 BXB200_FIND_PROBE=4
