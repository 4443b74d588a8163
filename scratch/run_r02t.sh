mkdir -p gpurun_out
python scratch/ab_aggregate.py | tail -1
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scores.py tests/test_gpu_round2.py -m gpu -q -k "aggregate or bucketed or set_ranges" 2>&1 | tail -2 )
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null > gpurun_out/r02t_legs.json
python -c "
import json
d=json.load(open('gpurun_out/r02t_legs.json'))
r=d['roofline']
print('c4', round(r['c4_ms'],3), {k[:26]:v['ms'] for k,v in d['extra']['bed_intersect']['kernels_rank0'].items()}, 'c5', round(r['c5_ms'],3), {k[:18]:v['ms'] for k,v in d['extra']['aggregate']['kernels_rank0'].items()}, 'ok', r['c4_parity_ok'], r['c5_parity_ok'])"
