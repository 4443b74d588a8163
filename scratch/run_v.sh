# run "v": int32-offset e2e path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "int32_offsets or c2_full or small_path or modes" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; python -c "
import json
d=json.load(open('gpurun_out/bench_v.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'],'i64',d['extra']['e2e_int64_offsets'],'serial',d['extra']['e2e_serial_copies'])
print('pcie',d['extra']['pcie'])
"
