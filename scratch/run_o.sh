# round-1 run "o": gpu tests, smoke, bench (+reference arm), ncu launch list + full captures of the find kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_o.json 2> gpurun_out/bench_ref_o.err; tail -c 600 gpurun_out/bench_ref_o.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; tail -5 gpurun_out/bench_o.err; python -c "
import json
d=json.load(open('gpurun_out/bench_o.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],d['roofline']['kernel'])
print(json.dumps(d['extra']['kernels']))
print(json.dumps(d['extra'].get('score_sources'),indent=0))
print('clocks',d['clocks'])
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_o.csv python bench.py --steps 2 --warmup 1 --no-bitset --no-cpu > gpurun_out/ncu_launch_o.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_find|k_fill" -s 6 -c 4 -o gpurun_out/prof_find_o -f python bench.py --steps 2 --warmup 1 --no-bitset --no-cpu > gpurun_out/ncu_find_o.log 2>&1
ls -la gpurun_out | tail -12
