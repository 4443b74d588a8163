"""The aggregate kernel of configs[4] alone (100 M scores, 5 M windows in shuffled BED order, one launch)."""
import ctypes as C
import sys
import numpy as np
sys.path.insert(0, ".")
from bx_python_b200 import _lib, synth
from bx_python_b200._lib import check, ptr
L = _lib.lib()
tracks = synth.genome_scores(100_000_000, 5_000_000, 5001)
hs = []
for origin, v, _, _ in tracks:
    h = C.c_void_p()
    check(L.bxg_scores_create(ptr(v), len(v), origin, _lib.HOST, C.byref(h)))
    hs.append(h)
ht = (C.c_void_p * 24)(*hs)
wt = np.concatenate([np.full(len(t[2]), k, np.int32) for k, t in enumerate(tracks)])
ws = np.concatenate([t[2] for t in tracks]); we = np.concatenate([t[3] for t in tracks])
perm = np.random.default_rng(50).permutation(len(wt))
d = [_lib.DeviceBuffer(a[perm]) for a in (wt, ws, we)]
nw = len(wt)
outs = [_lib.DeviceBuffer(np.zeros(nw, dt)) for dt in (np.float32, np.float32, np.int32, np.float32, np.float32)]
def run():
    check(L.bxg_aggregate_multi(ht, None, 24, d[0].ptr, d[1].ptr, d[2].ptr, nw, _lib.DEVICE, *[o.ptr for o in outs]))
run(); _lib.sync()
t = _lib.Timer(); t.start()
for _ in range(10):
    run()
t.stop()
print("aggregate_multi ms", round(t.elapsed_ms() / 10, 4))
