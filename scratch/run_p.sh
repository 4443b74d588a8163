# round-1 run "p": all gpu tests (no -x), bench, proper ncu full capture of the find kernels, sanitizer on the new kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; tail -3 gpurun_out/bench_p.err; python -c "
import json
d=json.load(open('gpurun_out/bench_p.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],d['roofline']['kernel'])
print(json.dumps(d['extra']['kernels']))
print(json.dumps(d['extra']['scalar_api']))
ss=d['extra'].get('score_sources') or {}
print({k:(v.get('ms')) for k,v in ss.items()})
print('clocks',d['clocks'])
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^(k_find|k_fill_staged)$' -s 8 -c 4 -o gpurun_out/prof_find_p -f python bench.py --steps 2 --warmup 1 --no-bitset --no-cpu > gpurun_out/ncu_find_p.log 2>&1; tail -2 gpurun_out/ncu_find_p.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scores.py tests/test_gpu_parity.py -m gpu -q -x -k "scores or spans or summarize_golden or join or small_path or wiggle or binned or quicksect" > gpurun_out/sanitizer_p.log 2>&1; echo "sanitizer rc=$?"; tail -4 gpurun_out/sanitizer_p.log
ls -la gpurun_out | tail -8
