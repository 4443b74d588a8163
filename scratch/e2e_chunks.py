"""e2e (bxg_itree_find_host, pinned host arrays in / pinned CSR out) for the current BXB200_CHUNK_QUERIES."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bx_python_b200 import _lib, synth
from bx_python_b200.intervals import IntervalForest
n = 10_000_000
db, qq = synth.genome_intervals(n, 2001), synth.genome_intervals(n, 2002)
cat = lambda per, k: np.concatenate([p[k] for p in per])
tid = np.concatenate([np.full(len(d[0]), c, np.int32) for c, d in enumerate(db)])
qt = np.concatenate([np.full(len(q[0]), c, np.int32) for c, q in enumerate(qq)])
perm = np.random.default_rng(7).permutation(n)
f = IntervalForest(24).build(tid, cat(db, 0), cat(db, 1))
pin = [_lib.PinnedArray(n, np.int32) for _ in range(3)]
for p, a in zip(pin, (qt[perm], cat(qq, 0)[perm], cat(qq, 1)[perm])):
    p.array[:] = a
ts = []
for it in range(8):
    t0 = time.perf_counter()
    off, hits = f.find_batch(pin[0].array, pin[1].array, pin[2].array, copy=False)
    ts.append(time.perf_counter() - t0)
best = min(ts[2:]); med = sorted(ts[2:])[len(ts[2:]) // 2]
print(f"chunk={os.environ.get('BXB200_CHUNK_QUERIES','default')} mode={os.environ.get('BXB200_FIND_MODE','auto')} "
      f"median {med*1e3:.2f} ms ({n/med/1e9:.3f} Gq/s) best {best*1e3:.2f} ms; hits {len(hits)}")
