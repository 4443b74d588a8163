"""A/B of the count_range kernel alone: 24 hg38 bit sets from the C4 law, 50 M shuffled count_range queries."""
import ctypes as C
import sys
import numpy as np
sys.path.insert(0, ".")
from bx_python_b200 import _lib, synth
from bx_python_b200._lib import check
from bx_python_b200.bitset import BinnedBitSet
L = _lib.lib()
n = 50_000_000
f2 = synth.genome_intervals(n // 4, 4002)          # lighter fill: the access pattern of the queries is what matters
f1 = synth.genome_intervals(n, 4001)
bits = [BinnedBitSet(int(sz)) for sz in synth.HG38_LENS]
for b, (s, e) in zip(bits, f2):
    b.set_ranges(s, e - s)
sets = (C.c_void_p * 24)(*[b._h for b in bits])
w = np.concatenate([np.full(len(f1[c][0]), c, np.int32) for c in range(24)])
s = np.concatenate([p[0] for p in f1])
c = np.concatenate([(p[1] - p[0]).astype(np.int32) for p in f1])
perm = np.random.default_rng(1).permutation(n)
d = [_lib.DeviceBuffer(a[perm]) for a in (w, s, c)]
out = _lib.DeviceBuffer(np.zeros(n, np.int32))
t = _lib.Timer()
def run():
    check(L.bxg_bits_count_ranges_multi(sets, 24, d[0].ptr, d[1].ptr, d[2].ptr, n, out.ptr, 1, _lib.DEVICE))
run(); _lib.sync()
t.start()
for _ in range(5):
    run()
t.stop()
got = np.empty(n, np.int32)
check(L.bxg_memcpy_d2h(got.ctypes.data_as(C.c_void_p), out.ptr, got.nbytes)); _lib.sync()
print("count_ranges_multi ms", round(t.elapsed_ms() / 5, 4), "checksum", int(got.astype(np.int64).sum()))
