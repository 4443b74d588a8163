# round 2, run "o": barrier-free bucketed set kernel
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_full_size.py tests/test_gpu_scores.py -m gpu -q -k "bucketed or c4 or set_ranges or clear or scores or binned or wiggle" 2>&1 | tail -4 ) 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null > gpurun_out/r02o_legs.json
python -c "
import json
d=json.load(open('gpurun_out/r02o_legs.json'))
r=d['roofline']
print('find', round(d['ms_per_step'],4), 'overlap', round(d['extra']['three_pass_overlapped_ms_per_step'],4), 'c4', round(r['c4_ms'],3), {k[:26]:v['ms'] for k,v in d['extra']['bed_intersect']['kernels_rank0'].items()}, 'c5', round(r['c5_ms'],3), 'ok', r['c4_parity_ok'], r['c5_parity_ok'], r['bitset_and_parity_ok'])"
