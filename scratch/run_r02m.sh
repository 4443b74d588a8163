# round 2, run "m": overlapped count/fill pipeline -- correctness + A/B of CTA shares and chunk counts
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "find or forest or tree" 2>&1 | tail -4 ) 2>&1
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-bitset --no-cpu 2>/dev/null > gpurun_out/r02m_$name.json
  python -c "
import json
d=json.load(open('gpurun_out/r02m_$name.json'))
print('$name', 'value_ms', round(d['ms_per_step'],4), 'serial', round(d['extra']['three_pass_serial_ms_per_step'],4), 'single', round(d['extra']['single_pass_kernel_ms_per_step'],4), 'sorted', round(d['extra']['sorted_queries_ms_per_step'],4), 'spot', d['extra']['parity_spot_check'][:40])"
}
run o43x4 BXB200_OVERLAP=4,3,4
run o43x8 BXB200_OVERLAP=4,3,8
run o33x4 BXB200_OVERLAP=3,3,4
run o52x4 BXB200_OVERLAP=5,2,4
run o44x6 BXB200_OVERLAP=4,4,6
run o34x4 BXB200_OVERLAP=3,4,4
run off BXB200_FIND_OVERLAP=0
