mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -k "bitset or c3 or flat" 2>&1 | tail -2 )
for v in 1 0; do
BXB200_PAIR_VIA_BATCH=$v timeout 300 python - <<'PY'
import sys, json, os
sys.path.insert(0, '.')
import bench
from bx_python_b200.dist import Comm
class A: pass
comm = Comm("nccl")
r = bench.leg_bitset(comm, 6547.5, A())
print('PAIR_VIA_BATCH', os.environ['BXB200_PAIR_VIA_BATCH'], {k: round(r[k]['frac'], 4) for k in ('and_genome', 'and_count_genome', 'and_per_pair', 'and_count_per_pair')}, r['parity_ok'])
PY
done
