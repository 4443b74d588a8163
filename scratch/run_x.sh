# run "x": set_ranges_multi, small-path kernel with staged splitters
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "set_ranges or builders or small_path or scalar or bitset_random or lotsa or node_root or neighbors_reference" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; python -c "
import json
d=json.load(open('gpurun_out/bench_x.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print(json.dumps(d['extra']['scalar_api'])[:100])
b=d['extra']['bed_intersect']
print({k:(v['ms'] if isinstance(v,dict) else v) for k,v in b.items()})
"
