mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -k "bucketed or set_ranges or builders" 2>&1 | tail -2 )
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null > gpurun_out/r02x_legs.json
python -c "
import json
d=json.load(open('gpurun_out/r02x_legs.json'))
r=d['roofline']
print('c4', round(r['c4_ms'],3), {k[:26]:v['ms'] for k,v in d['extra']['bed_intersect']['kernels_rank0'].items()}, 'c5', round(r['c5_ms'],3), 'ok', r['c4_parity_ok'], r['c5_parity_ok'], r['bitset_and_parity_ok'])"
