# run "h": confirm the reverted fused kernel (half-group loads, direct stores) + everything else
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; python -c "
import json
d=json.load(open('gpurun_out/bench_h.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'i64',d['extra']['e2e_int64_offsets'],'single',d['extra']['single_pass_kernel_ms_per_step'])
print(json.dumps(d['extra']['kernels']))
print('pcie', d['extra']['pcie'], 'clocks', d['clocks'])
"
