# run "u": full validation of the current state + e2e chunk-size sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_u.json 2> /dev/null
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err; python -c "
import json
d=json.load(open('gpurun_out/bench_u.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],d['roofline']['kernel'], d['roofline']['traffic'])
print(json.dumps(d['extra']['kernels']))
print(json.dumps(d['extra']['scalar_api'])[:120])
print('clocks',d['clocks'], 'launches', d['gpu_launches'])
"
for c in 262144 524288 2097152; do BXB200_CHUNK_QUERIES=$c timeout 200 python scratch/e2e_chunks.py 2>&1 | tail -1; done
timeout 200 python scratch/e2e_chunks.py 2>&1 | tail -1
BXB200_FIND_MODE=0 timeout 200 python scratch/e2e_chunks.py 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_u.csv python bench.py --steps 2 --warmup 1 --no-bitset --no-cpu > gpurun_out/ncu_launch_u.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^(k_find|k_fill_staged|k_find_fused)$' -s 8 -c 6 -o gpurun_out/prof_find_u -f python bench.py --steps 2 --warmup 1 --no-bitset --no-cpu > gpurun_out/ncu_find_u.log 2>&1; tail -1 gpurun_out/ncu_find_u.log
