"""
GPU tests of the round-2 entry points: per-chromosome counters, genome-wide totals, clear, the per-sector rank table,
NULL handles in the genome-wide calls, > 1024 sets, the pipelined host count, the copy probe and the device all-reduce.
"""
import ctypes as C

import numpy as np
import pytest

from bx_python_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from bx_python_b200 import _lib
    _lib.lib()
    return _lib


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


def test_group_stats_vs_numpy(lib):
    from bx_python_b200._lib import check, ptr
    L = lib.lib()
    rng = np.random.default_rng(11)
    for n, nkeys, thr in ((1, 1, 1), (1000, 24, 1), (3_000_000, 24, 5), (1_000_003, 1024, 0), (500_000, 3, -7)):
        key = rng.integers(-2, nkeys + 2, n).astype(np.int32)              # some keys outside [0, nkeys): ignored
        val = rng.integers(-50, 3000, n).astype(np.int32)
        val[rng.integers(0, n, min(n, 100))] = 2**31 - 1                   # large values: the sum must not wrap
        val[rng.integers(0, n, min(n, 100))] = -(2**31)
        if n > 1000:
            key[: n // 2] = np.sort(key[: n // 2])                         # long runs of one key (sorted BED input)
        stats = np.arange(2 * nkeys, dtype=np.int64)                       # accumulated INTO
        expect = stats.copy()
        inside = (key >= 0) & (key < nkeys)
        np.add.at(expect, 2 * key[inside & (val >= thr)], 1)
        np.add.at(expect, 2 * key[inside] + 1, val[inside].astype(np.int64))
        check(L.bxg_group_stats_i32(ptr(key), ptr(val), n, nkeys, thr, ptr(stats), lib.HOST))
        assert np.array_equal(stats, expect), (n, nkeys)
    assert L.bxg_group_stats_i32(ptr(key), ptr(val), 10, 2000, 0, ptr(stats), lib.HOST) != 0       # nkeys > 1024


def test_clear_totals_and_null_handles(lib, orc):
    from bx_python_b200._lib import check, ptr
    from bx_python_b200.bitset import BinnedBitSet
    L = lib.lib()
    rng = np.random.default_rng(12)
    sizes = [100_000, 777, 5_000_001, 256, 64 * 1024]
    sets = [BinnedBitSet(sz) for sz in sizes]
    oras = [orc.OracleBinnedBitSet(sz) for sz in sizes]
    which, starts, counts = [], [], []
    for k, sz in enumerate(sizes):
        m = 300
        s = rng.integers(0, sz, m)
        c = np.minimum(rng.integers(0, 900, m), sz - s)
        which += [k] * m
        starts += s.tolist()
        counts += c.tolist()
        oras[k].set_ranges(s, c)
    which, starts, counts = (np.asarray(a, np.int32) for a in (which, starts, counts))
    perm = rng.permutation(len(which))
    which, starts, counts = which[perm], starts[perm], counts[perm]
    # sets 1 and 3 are "another rank's": NULL handles -> their ranges are skipped, their counts and totals are 0
    h = (C.c_void_p * len(sets))(*[(None if k in (1, 3) else b._h) for k, b in enumerate(sets)])
    check(L.bxg_bits_set_ranges_multi(h, len(sets), ptr(which), ptr(starts), ptr(counts), len(which), lib.HOST))
    out = np.empty(len(which), np.int32)
    check(L.bxg_bits_count_ranges_multi(h, len(sets), ptr(which), ptr(starts), ptr(counts), len(which), ptr(out), 1, lib.HOST))
    tot = np.empty(len(sets), np.int64)
    check(L.bxg_bits_count_all_multi(h, len(sets), ptr(tot), 1, lib.HOST))
    for k in range(len(sets)):
        sel = which == k
        if k in (1, 3):
            assert not out[sel].any() and tot[k] == 0 and sets[k].count_all() == 0
        else:
            assert np.array_equal(out[sel], oras[k].count_ranges(starts[sel], counts[sel])), k
            assert tot[k] == oras[k].count_range(0, sizes[k]) == sets[k].count_all()
    # clear: bits and bin states back to a fresh bitset; the rank table must be rebuilt
    check(L.bxg_bits_clear(sets[0]._h))
    assert sets[0].count_all() == 0 and not sets[0].bin_states().any()
    assert not sets[0].count_ranges(starts[which == 0], counts[which == 0]).any()
    sets[0].invert()
    fresh = orc.OracleBinnedBitSet(sizes[0])
    fresh.invert()
    assert np.array_equal(sets[0].count_ranges(starts[which == 0], counts[which == 0]),
                          fresh.count_ranges(starts[which == 0], counts[which == 0]))     # strict ALL_ONE arithmetic


@pytest.mark.parametrize("size", [1, 63, 64, 255, 256, 257, 511, 512, 1000, 4096, 100_003])
def test_count_ranges_sector_rank_table(lib, orc, size):
    """Every (start, count) of a small bitmap (or a dense sample): ranges inside one 256-bit sector, across sector and
    word boundaries, ending at `size`, with a bitmap length that is / is not a multiple of 64 and 256."""
    from bx_python_b200.bitset import BinnedBitSet, BitSet
    rng = np.random.default_rng(size)
    for cls, ocls in ((BinnedBitSet, orc.OracleBinnedBitSet), (BitSet, orc.OracleBitSet)):
        b, o = cls(size), ocls(size)
        for _ in range(max(1, size // 40)):
            s = int(rng.integers(0, size))
            c = int(rng.integers(0, min(size - s, 70) + 1))
            b.set_range(s, c)
            o.set_range(s, c)
        if size <= 600:
            ss, cc = np.meshgrid(np.arange(size), np.arange(size + 1), indexing="ij")
            keep = ss + cc <= size
            ss, cc = ss[keep].astype(np.int32), cc[keep].astype(np.int32)
        else:
            ss = rng.integers(0, size, 20000).astype(np.int32)
            cc = np.minimum(rng.integers(0, 700, 20000), size - ss).astype(np.int32)
            ss[:10], cc[:10] = 0, size
        got = b.count_ranges(ss, cc)
        if ocls is orc.OracleBinnedBitSet:
            exp = o.count_ranges(ss, cc)
        else:
            exp = np.array([o.count_range(int(s), int(c)) for s, c in zip(ss[:3000], cc[:3000])], np.int32)
            got = got[:3000]
        assert np.array_equal(got, exp), cls.__name__


def test_more_than_1024_sets(lib, orc):
    """a scaffold-level assembly has more chromosomes than one launch's descriptor table."""
    from bx_python_b200.bitset import BinnedBitSet, and_count_many, count_ranges_many, ior_many, set_ranges_many
    rng = np.random.default_rng(13)
    nsets, size = 1500, 5000
    A = [BinnedBitSet(size) for _ in range(nsets)]
    B = [BinnedBitSet(size) for _ in range(nsets)]
    n = 40_000
    w = rng.integers(0, nsets, n).astype(np.int32)
    s = rng.integers(0, size - 200, n).astype(np.int32)
    c = rng.integers(0, 200, n).astype(np.int32)
    set_ranges_many(A, w, s, c)
    set_ranges_many(B, w[::-1].copy(), s, c)
    dense_a = np.zeros((nsets, size), bool)
    dense_b = np.zeros((nsets, size), bool)
    for k in range(n):
        dense_a[w[k], s[k]:s[k] + c[k]] = True
        dense_b[w[n - 1 - k], s[k]:s[k] + c[k]] = True
    qs = rng.integers(0, size - 300, n).astype(np.int32)
    qc = rng.integers(0, 300, n).astype(np.int32)
    qw = rng.integers(-1, nsets + 1, n).astype(np.int32)             # -1 / nsets: chromosome without a bitset -> 0
    got = count_ranges_many(A, qw, qs, qc)
    csum = np.concatenate([np.zeros((nsets, 1), np.int64), np.cumsum(dense_a, axis=1)], axis=1)
    inside = (qw >= 0) & (qw < nsets)
    exp = np.where(inside, csum[np.clip(qw, 0, nsets - 1), qs + qc] - csum[np.clip(qw, 0, nsets - 1), qs], 0)
    assert np.array_equal(got, exp.astype(np.int32))
    counts = and_count_many(A, B)
    assert np.array_equal(counts, (dense_a & dense_b).sum(axis=1))
    ior_many(A, B)
    assert [a.count_all() for a in A[::97]] == dense_b.sum(axis=1)[::97].tolist()       # (a & b) | b == b
    with pytest.raises(IndexError):
        set_ranges_many(A, np.array([1400], np.int32), np.array([size - 5], np.int32), np.array([10], np.int32))


def test_count_host_pipelined(lib, orc):
    """bxg_itree_count with HOST arrays and >= 2 chunks takes the copy/compute-overlapped path; same counts."""
    from bx_python_b200.intervals import IntervalForest
    rng = np.random.default_rng(14)
    n, nq = 400_000, 2_600_000
    s, e = synth.uniform_intervals(rng, n, 20_000_000, 3000)
    tid = rng.integers(0, 3, n).astype(np.int32)
    qs, qe = synth.uniform_intervals(rng, nq, 20_000_000, 3000)
    qt = rng.integers(0, 3, nq).astype(np.int32)
    f = IntervalForest(3).build(tid, s, e)
    got = f.count_batch(qt, qs, qe)
    small = f.count_batch(qt[:50_000], qs[:50_000], qe[:50_000])      # one chunk: the serial path
    assert np.array_equal(got[:50_000], small)
    exp = np.zeros(nq, np.int64)
    for t in range(3):
        o = orc.OracleIntervalTree(s[tid == t], e[tid == t])
        sel = qt == t
        exp[sel] = np.diff(o.find(qs[sel], qe[sel])[0])
    assert np.array_equal(got, exp.astype(np.int32))


def test_copy_probe_and_device_allreduce_single_rank(lib):
    from bx_python_b200._lib import check
    L = lib.lib()
    h2d, d2h, bi = C.c_double(), C.c_double(), C.c_double()
    check(L.bxg_copy_probe(32 << 20, 3, C.byref(h2d), C.byref(d2h), C.byref(bi)))
    assert 1.0 < h2d.value < 200 and 1.0 < d2h.value < 200 and bi.value > 0.6 * max(h2d.value, d2h.value)
    buf = lib.DeviceBuffer(np.arange(10, dtype=np.int64))
    check(L.bxg_comm_allreduce_i64_dev(buf.ptr, 10))                   # single rank: identity, no communicator needed
    out = np.empty(10, np.int64)
    check(L.bxg_memcpy_d2h(out.ctypes.data_as(C.c_void_p), buf.ptr, 80))
    lib.sync()
    assert out.tolist() == list(range(10))
    check(L.bxg_dev_memset(buf.ptr, 0, 80))
    check(L.bxg_memcpy_d2h(out.ctypes.data_as(C.c_void_p), buf.ptr, 80))
    lib.sync()
    assert not out.any()


# ---- interval operations, array in / array out, against the reference's own functions ----------------------
def _ops_gold():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "operations.json")))


def _rows(primary, res):
    out = []
    for k, a, b in zip(res["src"].tolist(), res["start"].tolist(), res["end"].tolist()):
        f = list(primary.fields[k])
        f[1], f[2] = str(a), str(b)
        out.append(f)
    return out


@pytest.mark.parametrize("seed", range(4))
def test_interval_operations_vs_reference_functions(lib, seed):
    from bx_python_b200.intervals.io import read_bed
    from bx_python_b200.intervals.operations import arrays as ops
    g = _ops_gold()[seed]
    p, s2, s3, lens = synth.ops_case(seed)
    for pieces in (True, False):
        for mincols in (1, 40):
            key = f"pieces{int(pieces)}_min{mincols}"
            P = read_bed(p)
            assert _rows(P, ops.intersect(P, [read_bed(s2)], mincols=mincols, pieces=pieces, lens=lens)) == g["intersect_" + key], key
            assert _rows(P, ops.subtract(P, [read_bed(s2)], mincols=mincols, pieces=pieces, lens=lens)) == g["subtract_" + key], key
    P = read_bed(p)
    assert _rows(P, ops.intersect(P, [read_bed(s2), read_bed(s3)], lens=lens)) == g["intersect3"]
    assert _rows(P, ops.subtract(P, [read_bed(s2), read_bed(s3)], lens=lens)) == g["subtract3"]
    m = ops.merge(read_bed(p))
    assert [[m["names"][c], str(a), str(b)] for c, a, b in zip(m["chrom"].tolist(), m["start"].tolist(), m["end"].tolist())] == g["merge"]
    c = ops.complement(read_bed(s2), lens)
    assert [[c["names"][k], str(a), str(b)] for k, a, b in zip(c["chrom"].tolist(), c["start"].tolist(), c["end"].tolist())] == \
           [[r[0], r[1], r[2]] for r in g["complement"]]
    for name, others in (("coverage", [s2]), ("coverage3", [s2, s3])):
        cov = ops.coverage(P, [read_bed(o) for o in others])
        assert len(cov["src"]) == len(g[name])
        for k, n, pc, row in zip(cov["src"].tolist(), cov["bases_covered"].tolist(), cov["percent"].tolist(), g[name]):
            assert list(P.fields[k]) == row[:-2] and str(n) == row[-2] and pc == float(row[-1]), (name, k)
    assert ops.base_coverage(read_bed(p)) == g["base_coverage"]


def test_bits_in_range_generators_match_reference_loops(lib, orc):
    """operations/__init__.py:10-33 on the device bit set vs the same loops over the oracle, incl. where they raise."""
    from bx_python_b200.bitset import BinnedBitSet
    from bx_python_b200.intervals.operations import bits_clear_in_range, bits_set_in_range

    def ref_set(bits, a, b):
        end = a
        while True:
            start = bits.next_set(end)
            end = min(bits.next_clear(start), b)
            if start >= end:
                break
            yield start, end

    def ref_clear(bits, a, b):
        end = a
        while True:
            start = bits.next_clear(end)
            if start >= b:
                break
            end = min(bits.next_set(start), b)
            yield start, end

    class Checked:                      # the oracle with the reference's IndexError on next_*(>= size) (bitset.pyx:222-227)
        def __init__(self, o):
            self.o, self.size = o, o.size

        def next_set(self, s):
            if s >= self.size or s < 0:
                raise IndexError(s)
            return self.o.next_set(s)

        def next_clear(self, s):
            if s >= self.size or s < 0:
                raise IndexError(s)
            return self.o.next_clear(s)
    rng = np.random.default_rng(21)
    size = 3000
    for trial in range(6):
        b, o = BinnedBitSet(size), orc.OracleBinnedBitSet(size)
        for _ in range(30):
            s = int(rng.integers(0, size - 1))
            c = int(rng.integers(0, min(150, size - s)))
            b.set_range(s, c)
            o.set_range(s, c)
        if trial % 2:
            b.set_range(size - 40, 40)
            o.set_range(size - 40, 40)
        for (a, z) in [(0, size), (0, size + 500), (100, 900), (size - 10, size), (500, 500), (size, size + 5)] + \
                      [tuple(sorted(rng.integers(0, size, 2).tolist())) for _ in range(12)]:
            for mine, ref in ((bits_set_in_range, ref_set), (bits_clear_in_range, ref_clear)):
                got, exp, gerr, eerr = [], [], False, False
                try:
                    for x in mine(b, a, z):
                        got.append(x)
                except IndexError:
                    gerr = True
                try:
                    for x in ref(Checked(o), a, z):
                        exp.append(x)
                except IndexError:
                    eerr = True
                assert got == exp and gerr == eerr, (trial, a, z, mine.__name__, got[-3:], exp[-3:], gerr, eerr)


def test_interleaved_insert_find_uses_tail(lib, orc):
    """insert / find interleaved (the `if not tree.find(s, e): tree.insert(s, e)` idiom): hits of recent inserts are merged
    at their in-order position; the device index is rebuilt once per TAIL_MAX inserts, not per find."""
    from bx_python_b200.intervals import IntervalTree
    rng = np.random.default_rng(31)
    t = IntervalTree()
    t.TAIL_MAX = 50
    S, E, builds, last_built = [], [], 0, 0
    for k in range(700):
        s = int(rng.integers(0, 300))
        e = s + int(rng.integers(-2, 25))
        t.insert(s, e, k)
        S.append(s)
        E.append(e)
        a = int(rng.integers(-5, 305))
        b = a + int(rng.integers(-1, 30))
        got = t.find(a, b)
        _, h = orc.OracleIntervalTree(np.array(S), np.array(E)).find([a], [b])
        assert got == h.tolist(), k
        if t._built != last_built:
            builds, last_built = builds + 1, t._built
    assert builds <= 700 // 50 + 2
    assert t.before(150, 3, 100) is not None and t._built == 700          # neighbour queries need the full index


def test_scalar_next_far_and_near(lib, orc):
    from bx_python_b200.bitset import BinnedBitSet, BitSet
    size = 3_000_000
    for cls in (BinnedBitSet, BitSet):
        b = cls(size)
        assert b.next_set(0) == size and b.next_clear(5) == 5
        b.set_range(2_500_000, 10)
        assert b.next_set(0) == 2_500_000                      # far beyond the warp's near scan: falls through to the grid kernel
        assert b.next_set(2_499_990) == 2_500_000 and b.next_set(2_500_003) == 2_500_003
        assert b.next_clear(2_500_000) == 2_500_010 and b.next_set(2_500_010) == size
        b.set_range(0, 40_000)
        assert b.next_clear(0) == 40_000 and b.next_clear(39_999) == 40_000 and b.next_set(100) == 100
        b.set_range(40_000, size - 40_000)
        assert b.next_clear(7) == size                         # nothing clear: scans to the end
        if cls is BitSet:
            assert b.next_clear(7, 1000) == 1000 and b.next_set(7, 7) == 7


BUCKET_SNIPPET = r"""
import sys
import numpy as np
sys.path.insert(0, %r)
from bx_python_b200.bitset import BinnedBitSet, set_ranges_many
from oracle import oracle as orc
rng = np.random.default_rng(5)
for nsets, size in ((5, 3_000_000), (300, 40_000)):
    sets = [BinnedBitSet(size + 17 * k) for k in range(nsets)]
    n = 200_000
    w = rng.integers(0, nsets, n).astype(np.int32)
    s = rng.integers(0, size - 3000, n).astype(np.int32)
    c = rng.integers(0, 3000, n).astype(np.int32)
    c[:50] = size // 2                                    # a few very long ranges (CTA-wide interior sweep)
    s[:50] = rng.integers(0, size // 2 - 10, 50)
    set_ranges_many(sets, w, s, c)
    for k in range(0, nsets, max(1, nsets // 7)):
        o = orc.OracleBinnedBitSet(size + 17 * k)
        o.set_ranges(s[w == k], c[w == k])
        assert np.array_equal(sets[k].to_words(), o.words()), (nsets, k)
        assert np.array_equal(sets[k].bin_states(), o.states()), (nsets, k)
print("bucketed ok")
"""


def test_set_ranges_bucketed_path(lib):
    """BXB200_SET_BUCKETS=1 forces the locality-bucketed form of bxg_bits_set_ranges_multi (radix pass on (set, position)
    buckets, then the same kernel over the bucketed records) -- few large sets and > 128 small sets (grouped buckets)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", BUCKET_SNIPPET % root], capture_output=True, text=True, timeout=600,
                       env={**os.environ, "BXB200_SET_BUCKETS": "1"})
    assert r.returncode == 0 and "bucketed ok" in r.stdout, r.stderr[-3000:]
