# round 2, run "a": all GPU tests (incl. the full-size goldens), default bench, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r02a_pytest.log
( time timeout 600 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err ) 2>&1 | tail -4
tail -12 gpurun_out/r02a_bench.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02a_bench.json'))
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])
    print('roofline', json.dumps(d['roofline'], indent=0)[:3000])
    ex = d['extra']
    for k in ('bitset', 'bed_intersect', 'aggregate'):
        print(k, json.dumps(ex.get(k), indent=0)[:2500])
    print('kernels', ex['kernels'], 'copy_probe', ex['copy_probe'], 'scalar', ex['scalar_api'])
except Exception as e:
    print('bench parse failed', e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --c4-steps 1 --c5-steps 1 > gpurun_out/r02a_ncu_bench.log 2>&1; tail -2 gpurun_out/r02a_ncu_bench.log | cut -c1-300
