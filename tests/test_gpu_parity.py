"""
GPU parity: the CUDA path (through the Python shim -> ctypes -> C ABI of libbxb200.so) against
  * the committed golden vectors produced by the compiled unmodified reference (tests/golden/),
  * the reference's own unit-test known answers and doctests,
  * the CPU restatement (oracle/) on seeded random inputs,
bit-exact everywhere (integer / index work) and bit-exact float32 for the aggregate path.
"""
import json
import os

import numpy as np
import pytest

from bx_python_b200 import synth

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="module")
def bx():
    import bx_python_b200.bitset as bitset
    import bx_python_b200.intervals.intersection as ix
    from bx_python_b200 import _lib, aggregate
    _lib.lib()   # raises if there is no device: GPU tests must never pass on a fallback

    class NS:
        pass
    ns = NS()
    ns.bitset, ns.ix, ns.aggregate, ns.lib = bitset, ix, aggregate, _lib
    return ns


def tree_of(bx, s, e):
    t = bx.ix.IntervalTree()
    t.insert_many(s, e)
    return t


# ---------------------------------------------------------------------------------------------------------------
# find
# ---------------------------------------------------------------------------------------------------------------
def test_find_c1_golden(bx):
    g = np.load(os.path.join(G, "find.npz"))
    s, e, qs, qe = synth.c1_intervals()
    t = tree_of(bx, s, e)
    off, hits = t.find_batch(qs, qe)
    assert np.array_equal(off, g["c1_offsets"])
    assert np.array_equal(hits, g["c1_hits"])
    assert np.array_equal(t.order(), g["c1_order"])
    assert np.array_equal(t.count_batch(qs, qe), np.diff(g["c1_offsets"]).astype(np.int32))


def test_find_single_pass_and_three_pass_agree(bx, orc):
    """Both find implementations (single-pass look-back kernel / count+scan+fill) give the oracle's ordered CSR, on the
    device-array path and on the chunk-pipelined host path (>1 chunk, hit-buffer regrowth on a fresh index)."""
    import ctypes as C
    rng = np.random.default_rng(4242)
    n, nq = 600_000, 2_500_000
    s, e = synth.uniform_intervals(rng, n, 30_000_000, 3000)
    qs, qe = synth.uniform_intervals(rng, nq, 30_000_000, 3000)
    ooff, ohits = orc.OracleIntervalTree(s, e).find(qs, qe)
    L = bx.lib.lib()
    try:
        for mode in (0, 1, 2, -1):                      # 2 / auto: count and fill of different chunks overlapped on two streams
            bx.lib.check(L.bxg_set_find_mode(mode))
            t = tree_of(bx, s, e)                       # fresh index: first find must grow its hit buffers
            off, hits = t.find_batch(qs, qe)            # pipelined host path, 2 chunks
            assert np.array_equal(off, ooff) and np.array_equal(hits, ohits), mode
            off, hits = t.find_batch(qs[:1000], qe[:1000])
            assert np.array_equal(off, ooff[:1001]) and np.array_equal(hits, ohits[:ooff[1000]]), mode
            total = C.c_int64()
            h = t._index._h
            bx.lib.check(L.bxg_itree_find(h, None, bx.lib.ptr(qs), bx.lib.ptr(qe), nq, bx.lib.HOST, C.byref(total)))
            off2 = np.empty(nq + 1, np.int64); hits2 = np.empty(total.value, np.int32)
            bx.lib.check(L.bxg_itree_fetch(h, bx.lib.ptr(off2), bx.lib.ptr(hits2)))
            assert np.array_equal(off2, ooff) and np.array_equal(hits2, ohits), mode
            # a second, larger batch on the same index outgrows the hit buffer: the speculative fills back off and re-run
            big_s, big_e = np.concatenate([qs, qs[:700_000]]), np.concatenate([qe, qe[:700_000] + 5000])
            bx.lib.check(L.bxg_itree_find(h, None, bx.lib.ptr(big_s), bx.lib.ptr(big_e), len(big_s), bx.lib.HOST, C.byref(total)))
            off3 = np.empty(len(big_s) + 1, np.int64); hits3 = np.empty(total.value, np.int32)
            bx.lib.check(L.bxg_itree_fetch(h, bx.lib.ptr(off3), bx.lib.ptr(hits3)))
            assert np.array_equal(off3[:nq + 1], ooff) and np.array_equal(hits3[:ooff[-1]], ohits), mode
            boff, bhits = orc.OracleIntervalTree(s, e).find(big_s[nq:], big_e[nq:])
            assert np.array_equal(off3[nq:] - off3[nq], boff) and np.array_equal(hits3[ooff[-1]:], bhits), mode
    finally:
        bx.lib.check(L.bxg_set_find_mode(-1))


def test_find_edge_sets_golden(bx):
    g = np.load(os.path.join(G, "find.npz"))
    for k, (s, e, qs, qe) in enumerate(synth.edge_sets()):
        t = tree_of(bx, s, e)
        off, hits = t.find_batch(qs, qe)
        assert np.array_equal(off, g[f"edge{k}_offsets"]), k
        assert np.array_equal(hits, g[f"edge{k}_hits"]), k
        assert np.array_equal(t.order(), g[f"edge{k}_order"]), k


def test_find_scalar_api_reference_answers(bx):
    Interval, IntervalTree = bx.ix.Interval, bx.ix.IntervalTree
    # SURVEY 8(a) addendum 2 (probed on the compiled reference)
    t = IntervalTree()
    for i, (a, b) in enumerate([(10, 20), (15, 12), (30, 30), (-5, 3), (20, 25)]):
        t.insert(a, b, i)
    assert t.find(-100, 100) == [3, 0, 1, 4, 2]
    assert t.find(18, 11) == [0]
    assert t.find(20, 20) == []
    assert t.find(19, 21) == [0, 4]
    assert t.find(30, 30) == []
    assert t.find(29, 31) == [2]
    assert t.find(1.5, 3) == [3]
    with pytest.raises(OverflowError):
        t.find(2**31, 5)
    t2 = IntervalTree()
    for name, (a, b) in zip("a b z0 c z1 d e y".split(), [(5, 10), (5, 7), (5, 5), (5, 20), (5, 5), (3, 6), (5, 6), (4, 4)]):
        t2.insert(a, b, name)
    assert t2.find(0, 100) == "d y z1 z0 a b c e".split()
    # doctest intersection.pyx:341-361
    it = IntervalTree()
    it.insert(0, 10, "food")
    it.insert(3, 7, dict(foo="bar"))
    assert it.find(2, 5) == ["food", {"foo": "bar"}]
    it = IntervalTree()
    for a, b in [(0, 10), (3, 7), (3, 40), (13, 50)]:
        it.insert_interval(Interval(a, b))
    assert repr(it.find(30, 50)) == "[Interval(3, 40), Interval(13, 50)]"
    assert it.find(100, 200) == []
    # intersection_tests.py:158-186
    iv = IntervalTree()
    n = 0
    for i in range(1, 1000, 80):
        iv.insert(i, i + 10, {"value": i * i})
        iv.add(i + 20, i + 30, {"astr": str(i * i)})
        iv.insert_interval(Interval(i + 40, i + 50, value={"astr": str(i * i)}))
        iv.add_interval(Interval(i + 60, i + 70, value={"astr": str(i * i)}))
        n += 4
    assert len(iv.find(100, 200)) == 5
    seen = []
    iv.traverse(seen.append)
    assert len(seen) == n and all(node.interval for node in seen)
    # lazily rebuilt after a mutation
    iv.insert(150, 160, "late")
    r = iv.find(100, 200)
    assert len(r) == 6 and sum(1 for v in r if isinstance(v, str) and v == "late") == 1


def test_find_lotsa_reference_case(bx):
    # intersection_tests.py:104-155: 100k zero-length intervals + 600 duplicates of (0,1)
    mx = 1000000
    s = [1] + list(range(0, mx, 10)) + [0] * 600
    e = [2] + list(range(0, mx, 10)) + [1] * 600
    t = tree_of(bx, s, e)
    rng = np.random.default_rng(5)
    qs = rng.integers(0, mx - 10000, 25).astype(np.int32)
    qe = (qs + rng.integers(100, 10000, 25)).astype(np.int32)
    off, hits = t.find_batch(qs, qe)
    S, E = np.array(s), np.array(e)
    for q in range(25):
        h = hits[off[q]:off[q + 1]]
        assert len(h) > 0
        assert np.all(((E[h] >= qs[q]) & (E[h] <= qe[q])) | ((S[h] <= qe[q]) & (S[h] >= qs[q])))
    off2, hits2 = t.find_batch([0], [1])
    assert off2[1] == 600   # the 600 (0,1) duplicates; the zero-length (0,0) item does not overlap [0,1)


@pytest.mark.parametrize("seed", range(6))
def test_find_random_vs_oracle(bx, orc, seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(1, 200000))
    nq = 50000
    G_ = int(rng.choice([50, 5000, 1_000_000, 200_000_000]))
    if seed % 2:
        s = rng.integers(-3, G_, n)
        e = s + rng.integers(-2, 6, n)
        qs = rng.integers(-5, G_ + 5, nq)
        qe = qs + rng.integers(-3, 12, nq)
    else:
        s, e = synth.uniform_intervals(rng, n, G_ + 3000, 2000)
        qs, qe = synth.uniform_intervals(rng, nq, G_ + 3000, 2000)
    s, e, qs, qe = (np.asarray(a, np.int32) for a in (s, e, qs, qe))
    off, hits = tree_of(bx, s, e).find_batch(qs, qe)
    ooff, ohits = orc.OracleIntervalTree(s, e).find(qs, qe)
    assert np.array_equal(off, ooff) and np.array_equal(hits, ohits)


def test_find_adversarial_long_intervals(bx, orc):
    """A few chromosome-long intervals at the front defeat the prefix-max bound; the max hierarchy must skip."""
    rng = np.random.default_rng(77)
    n, nq, G_ = 300000, 20000, 50_000_000
    s, e = synth.uniform_intervals(rng, n, G_, 500)
    s[:3] = [0, 5, 10]
    e[:3] = [G_, G_ - 7, G_ // 2]
    qs, qe = synth.uniform_intervals(rng, nq, G_, 500)
    off, hits = tree_of(bx, s, e).find_batch(qs, qe)
    ooff, ohits = orc.OracleIntervalTree(s, e).find(qs, qe)
    assert np.array_equal(off, ooff) and np.array_equal(hits, ohits)


def test_forest_vs_per_tree_oracle(bx, orc):
    """hg38-shaped: one index holding 24 trees, queries carry their chromosome id."""
    n, nq = 400000, 400000
    db = synth.genome_intervals(n, 2001)
    qq = synth.genome_intervals(nq, 2002)
    tid = np.concatenate([np.full(len(s), c, np.int32) for c, (s, e) in enumerate(db)])
    S = np.concatenate([s for s, e in db]); E = np.concatenate([e for s, e in db])
    qt = np.concatenate([np.full(len(s), c, np.int32) for c, (s, e) in enumerate(qq)])
    QS = np.concatenate([s for s, e in qq]); QE = np.concatenate([e for s, e in qq])
    # shuffle items and queries so neither arrives grouped by chromosome
    rng = np.random.default_rng(9)
    p = rng.permutation(len(S)); tid, S, E = tid[p], S[p], E[p]
    pq = rng.permutation(len(QS)); qt, QS, QE = qt[pq], QS[pq], QE[pq]
    forest = bx.ix.IntervalForest(24).build(tid, S, E)
    off, hits = forest.find_batch(qt, QS, QE)
    cnt = forest.count_batch(qt, QS, QE)
    assert np.array_equal(cnt, np.diff(off).astype(np.int32))
    for c in range(24):
        items = np.nonzero(tid == c)[0]
        o = orc.OracleIntervalTree(S[items], E[items])
        qsel = np.nonzero(qt == c)[0]
        ooff, ohits = o.find(QS[qsel], QE[qsel])
        ohits = items[ohits].astype(np.int32)          # local -> global item ids
        for j in range(0, len(qsel), max(1, len(qsel) // 300)):
            q = qsel[j]
            assert np.array_equal(hits[off[q]:off[q + 1]], ohits[ooff[j]:ooff[j + 1]])
        assert np.array_equal(np.diff(off)[qsel], np.diff(ooff))
    # out-of-range chromosome ids yield no hits
    off2, hits2 = forest.find_batch([24, -1], [0, 0], [10**9, 10**9])
    assert off2.tolist() == [0, 0, 0]


def test_find_large_properties(bx):
    """Size-independent checks at 2M x 2M: every hit satisfies the predicate, per-query hits are in index order,
    counts equal an independent rank formula (valid because every synthetic interval is proper)."""
    n = nq = 2_000_000
    s, e = synth.uniform_intervals(np.random.default_rng(11), n, 248956422)
    qs, qe = synth.uniform_intervals(np.random.default_rng(12), nq, 248956422)
    t = tree_of(bx, s, e)
    off, hits = t.find_batch(qs, qe)
    cnt = np.diff(off)
    q_of_hit = np.repeat(np.arange(nq), cnt)
    assert np.all(e[hits] > qs[q_of_hit]) and np.all(s[hits] < qe[q_of_hit])
    expected = np.searchsorted(np.sort(s), qe, "left") - np.searchsorted(np.sort(e), qs, "right")
    assert np.array_equal(cnt, expected)
    order = t.order()
    rank = np.empty(n, np.int64); rank[order] = np.arange(n)
    r = rank[hits]
    inc = np.ones(len(hits), bool); inc[1:] = r[1:] > r[:-1]
    inc[off[:-1][cnt > 0]] = True
    assert inc.all()


def test_find_int32_offsets(bx):
    """find_batch(offsets32=True): same CSR with int32 offsets (bxg_itree_find_host32), single tree and forest, several
    pipeline chunks."""
    rng = np.random.default_rng(31)
    s, e = synth.uniform_intervals(rng, 300_000, 5_000_000)
    qs, qe = synth.uniform_intervals(rng, 2_500_000, 5_000_000)
    t = tree_of(bx, s, e)
    off64, hits64 = t.find_batch(qs, qe)
    off32, hits32 = t.find_batch(qs, qe, offsets32=True)
    assert off32.dtype == np.int32 and off64.dtype == np.int64
    assert np.array_equal(off32, off64) and np.array_equal(hits32, hits64)
    tid = (np.arange(len(s)) % 5).astype(np.int32)
    qt = (np.arange(len(qs)) % 6).astype(np.int32)
    f = bx.ix.IntervalForest(5).build(tid, s, e)
    a = f.find_batch(qt, qs, qe)
    b = f.find_batch(qt, qs, qe, offsets32=True)
    assert b[0].dtype == np.int32 and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    e0 = bx.ix.IntervalTree().find_batch([1, 2], [3, 4], offsets32=True)
    assert e0[0].tolist() == [0, 0, 0] and len(e0[1]) == 0


def test_find_small_path(bx, orc):
    """bxg_itree_find_small (the scalar IntervalTree.find route): up to 32 queries answered by one warp into mapped
    pinned memory; identical CSR to the batched path, incl. the fallback when more than 65536 hits come back."""
    import ctypes as C
    L = bx.lib.lib()
    rng = np.random.default_rng(21)
    s, e = synth.uniform_intervals(rng, 50_000, 2_000_000)
    tid = (np.arange(len(s)) % 3).astype(np.int32)
    forest = bx.ix.IntervalForest(3).build(tid, s, e)
    for nq in (0, 1, 2, 31, 32):
        qs, qe = synth.uniform_intervals(rng, max(nq, 1), 2_000_000)
        qs, qe = qs[:nq].copy(), qe[:nq].copy()
        qt = rng.integers(-1, 4, nq).astype(np.int32)                 # -1 and 3 are not trees: no hits
        p_off, p_hits, total = C.c_void_p(), C.c_void_p(), C.c_int64()
        bx.lib.check(L.bxg_itree_find_small(forest.handle, bx.lib.ptr(qt), bx.lib.ptr(qs), bx.lib.ptr(qe), nq,
                                            C.byref(p_off), C.byref(p_hits), C.byref(total)))
        off = np.frombuffer((C.c_int64 * (nq + 1)).from_address(p_off.value), np.int64).copy()
        hits = np.frombuffer((C.c_int32 * max(total.value, 1)).from_address(p_hits.value), np.int32)[:total.value].copy()
        eoff, ehits = forest.find_batch(qt, qs, qe)
        assert np.array_equal(off, eoff) and np.array_equal(hits, ehits)
    # scalar API on a single tree against the oracle, incl. an answer too large for the mapped buffer
    t = tree_of(bx, s, e)
    o = orc.OracleIntervalTree(s, e)
    for a, b in ((0, 2_000_000), (1000, 1001), (5, 5), (1_999_999, 3_000_000), (-10, 0)):
        assert t.find(a, b) == o.find([a], [b])[1].tolist()
    big = bx.ix.IntervalTree()
    big.insert_many(np.zeros(70_000, np.int32), np.full(70_000, 10, np.int32))
    assert big.find(0, 5) == list(range(70_000))                     # > 65536 hits: falls back to the general path
    assert big.find(10, 20) == []


def test_find_c2_full_size_vs_oracle(bx, orc):
    """BASELINE configs[1] at full size -- 10 M hg38-shaped intervals vs 10 M queries (bench.py's workload, seeds
    2001 / 2002), one forest of 24 chromosomes, queries in shuffled (file) order: offsets and ordered hit lists of
    EVERY query bit-identical to the oracle, chromosome by chromosome."""
    n = nq = 10_000_000
    db, qq = synth.genome_intervals(n, 2001), synth.genome_intervals(nq, 2002)
    tid = np.concatenate([np.full(len(d[0]), c, np.int32) for c, d in enumerate(db)])
    s, e = np.concatenate([d[0] for d in db]), np.concatenate([d[1] for d in db])
    qt = np.concatenate([np.full(len(q[0]), c, np.int32) for c, q in enumerate(qq)])
    qs, qe = np.concatenate([q[0] for q in qq]), np.concatenate([q[1] for q in qq])
    perm = np.random.default_rng(7).permutation(nq)
    forest = bx.ix.IntervalForest(24).build(tid, s, e)
    off, hits = forest.find_batch(qt[perm], qs[perm], qe[perm], copy=False)
    cnt = np.diff(off)
    base = np.concatenate([[0], np.cumsum([len(d[0]) for d in db])])
    qbase = np.concatenate([[0], np.cumsum([len(q[0]) for q in qq])])
    inv = np.empty(nq, np.int64)
    inv[perm] = np.arange(nq)                          # position of original query j in the shuffled batch
    total = 0
    for c in range(24):
        ooff, ohits = orc.OracleIntervalTree(*db[c]).find(*qq[c])
        where = inv[qbase[c]:qbase[c + 1]]
        assert np.array_equal(cnt[where], np.diff(ooff)), f"counts differ on chromosome {c}"
        # gather this chromosome's hit lists back into original query order and compare them wholesale
        starts = off[where]
        lens = np.diff(ooff)
        idx = np.repeat(starts - ooff[:-1], lens) + np.arange(len(ohits))
        assert np.array_equal(hits[idx] - base[c], ohits), f"hit lists differ on chromosome {c}"
        total += len(ohits)
    assert total == off[-1] == len(hits)


# ---------------------------------------------------------------------------------------------------------------
# neighbours: before / after / *_interval / upstream / downstream  (intersection.pyx:192-260, 408-477)
# ---------------------------------------------------------------------------------------------------------------
def test_neighbors_golden(bx):
    for case in json.load(open(os.path.join(G, "neighbors.json"))):
        s, e, queries = synth.neighbor_case(case["seed"])
        t = tree_of(bx, s, e)
        for (pos, k, md), (b, a) in zip(queries, case["results"]):
            assert t.before(pos, k, md) == b, (case["seed"], pos, k, md)
            assert t.after(pos, k, md) == a, (case["seed"], pos, k, md)


def test_neighbors_reference_unit_tests(bx):
    Interval, IntervalTree = bx.ix.Interval, bx.ix.IntervalTree
    # doctests intersection.pyx:363-378
    it = IntervalTree()
    for a, b in [(0, 10), (3, 7), (3, 40), (13, 50)]:
        it.insert_interval(Interval(a, b))
    assert repr(it.before_interval(Interval(10, 20))) == "[Interval(3, 7)]"
    assert it.before_interval(Interval(5, 20)) == []
    assert repr(it.upstream_of_interval(Interval(11, 12))) == "[Interval(0, 10)]"
    assert repr(it.upstream_of_interval(Interval(11, 12, strand="-"))) == "[Interval(13, 50)]"
    assert repr(it.upstream_of_interval(Interval(1, 2, strand="-"), num_intervals=3)) == \
        "[Interval(3, 7), Interval(3, 40), Interval(13, 50)]"
    # intersection_tests.py:57-101 UpDownStreamTestCase
    iv = IntervalTree()
    iv.add_interval(Interval(50, 59))
    for i in range(0, 110, 10):
        if i != 50:
            iv.add_interval(Interval(i, i + 9))
    assert all(u.end < 59 for u in iv.upstream_of_interval(Interval(59, 60), num_intervals=200))
    assert all(u.start > 70 for u in iv.upstream_of_interval(Interval(60, 70, strand=-1), num_intervals=200))
    assert all(u.start > 59 for u in iv.upstream_of_interval(Interval(58, 58, strand=-1), num_intervals=200))
    assert all(d.start > 60 for d in iv.downstream_of_interval(Interval(59, 60), num_intervals=200))
    assert all(d.start < 59 for d in iv.downstream_of_interval(Interval(59, 60, strand=-1), num_intervals=200))
    for i in range(0, 90, 10):
        r = iv.after(i, max_dist=20, num_intervals=2)
        assert (r[0].start, r[1].start) == (i + 10, i + 20)
        r = iv.after_interval(Interval(i, i), max_dist=20, num_intervals=2)
        assert (r[0].start, r[1].start) == (i + 10, i + 20)
    # intersection_tests.py:17-54 NeighborTestCase (left == before, right == after)
    assert str(iv.before(60, 2)) == str([Interval(50, 59), Interval(40, 49)])
    for i in range(10, 100, 10):
        assert iv.before(i, 1, 10)[0].end == i - 1
    assert len(iv.before(60, 200)) == 6
    for i in range(10, 100, 10):
        assert iv.after(i + 1, 1)[0].start == i + 10
    for i in range(0, 100, 10):
        assert iv.after(i - 1, 1, 10)[0].start == i
    # empty tree (:188-201)
    e = IntervalTree()
    assert e.after(100) == e.before(100) == e.after_interval(100) == e.before_interval(100) == []
    assert e.upstream_of_interval(100) == e.downstream_of_interval(100) == []


def test_interval_node_root_handle(bx):
    """intersection_tests.py:17-54 NeighborTestCase, written against IntervalNode exactly as the reference test is."""
    Interval, IntervalNode = bx.ix.Interval, bx.ix.IntervalNode
    iv = IntervalNode(50, 59, Interval(50, 59))
    for i in range(0, 110, 10):
        if i == 50:
            continue
        f = Interval(i, i + 9)
        iv = iv.insert(f.start, f.end, f)
    assert str(iv.left(60, n=2)) == str([Interval(50, 59), Interval(40, 49)])
    for i in range(10, 100, 10):
        assert iv.left(i, max_dist=10, n=1)[0].end == i - 1
    assert len(iv.left(60, n=200)) == 6
    for i in range(10, 100, 10):
        r = iv.right(i + 1, n=1)
        assert len(r) == 1 and r[0].start == i + 10
    for i in range(0, 100, 10):
        assert iv.right(i - 1, max_dist=10, n=1)[0].start == i
    assert [x.start for x in iv.find(45, 72)] == [40, 50, 60, 70]
    seen = []
    iv.traverse(lambda node: seen.append((node.start, node.end)))
    assert seen == [(i, i + 9) for i in range(0, 110, 10)]


def test_neighbors_lotsa(bx):
    # intersection_tests.py:104-140 LotsaTestCase.test_count / test_max_dist
    mx = 1000000
    s = [1] + list(range(0, mx, 10)) + [0] * 600
    e = [2] + list(range(0, mx, 10)) + [1] * 600
    t = tree_of(bx, s, e)
    assert len(t.after(1, 33)) == 33
    assert len(t.before(1, 33)) == 1
    assert len(t.after(1, 9999)) == 250
    assert len(t.after(1, 9999, 99999)) == 9999
    assert len(t.after(1, 10, 0)) == 0
    for n, d in enumerate(range(10, 1000, 10)):
        assert len(t.after(1, 10000, d)) == n + 1


def test_neighbors_random_vs_oracle(bx, orc):
    rng = np.random.default_rng(808)
    for trial in range(6):
        n = int(rng.integers(1, 50000))
        G_ = int(rng.integers(50, 2_000_000))
        s = rng.integers(0, G_, n).astype(np.int32)
        e = (s + rng.integers(0, 300, n)).astype(np.int32)
        t = tree_of(bx, s, e)
        o = orc.OracleIntervalTree(s, e)
        for _ in range(60):
            pos = int(rng.integers(-5, G_ + 300)); k = int(rng.integers(1, 8)); md = int(rng.choice([0, 1, 30, 2500, 10**6]))
            assert t.before(pos, k, md) == o.before(pos, k, md).tolist()
            assert t.after(pos, k, md) == o.after(pos, k, md).tolist()


# ---------------------------------------------------------------------------------------------------------------
# bitsets
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cls", ["flat", "binned"])
def test_bitset_reference_unit_tests(bx, cls):
    # lib/bx/bitset_tests.py:10-119
    def new(size=100):
        return bx.bitset.BitSet(size) if cls == "flat" else bx.bitset.BinnedBitSet(size, size % 11)

    def bits(b):
        return [b[i] for i in range(b.size)]
    with pytest.raises(ValueError):
        new(4000000000)
    b = new()
    with pytest.raises(IndexError):
        b.set(-5)
    with pytest.raises(IndexError):
        b.set(110)
    l = [0] * 100
    assert bits(b) == l
    for pos in (11, 14, 70, 16):
        b.set(pos); l[pos] = 1
    for pos in (14, 80, 16):
        b.clear(pos); l[pos] = 0
    assert bits(b) == l
    b = new(); l = [0] * 100
    for s, e in ((11, 14), (20, 75), (90, 99)):
        b.set_range(s, e - s)
        for p in range(s, e):
            l[p] = 1
    assert bits(b) == l
    b = new()
    for s, e in ((11, 14), (20, 75), (90, 100)):
        b.set_range(s, e - s)
    assert [b.count_range(0, 0), b.count_range(0, 20), b.count_range(25, 25), b.count_range(80, 20),
            b.count_range(0, 100)] == [0, 3, 25, 10, 68]
    assert [b.next_set(0), b.next_set(13), b.next_set(15)] == [11, 13, 20]
    assert [b.next_clear(0), b.next_clear(11), b.next_clear(20), b.next_clear(92)] == [0, 14, 75, 100]
    b1, b2 = new(), new(); b1.set_range(20, 40); b2.set_range(50, 25); b1.iand(b2)
    assert bits(b1) == [1 if 50 <= i < 60 else 0 for i in range(100)]
    b1, b2 = new(), new(); b1.set_range(20, 40); b2.set_range(50, 25); b1.ior(b2)
    assert bits(b1) == [1 if 20 <= i < 75 else 0 for i in range(100)]
    b1 = new(); b1.set_range(20, 40); b1.invert()
    assert bits(b1) == [0 if 20 <= i < 60 else 1 for i in range(100)]


def test_bitset_survey_probes(bx):
    B = bx.bitset.BinnedBitSet
    geo = {(2**29, 1024): 524288, (250000000, 1024): 244141, (248956422, 1024): 243122, (100, 1): 100,
           (100, 3): 34, (100, 1024): 1, (1000, 10): 100, (1000, 20): 50}
    for (size, gran), bin_size in geo.items():
        assert B(size, gran).bin_size == bin_size
    b = B(1000, 10)
    for args, msg in [((-1, 5), r"BitSet index \(-1\) must be non-negative."),
                      ((1000, 0), r"1000 is larger than the size of this BitSet \(1000\)."),
                      ((999, 2), r"End \(1001\) is larger than the size of this BinnedBitSet \(1000\)."),
                      ((5, -1), r"Count \(-1\) must be non-negative.")]:
        with pytest.raises(IndexError, match=msg):
            b.set_range(*args)
    b.set_range(5, 0)
    assert b.count_range(0, 0) == 0
    with pytest.raises(IndexError):
        b.count_range(1000, 0)
    with pytest.raises(IndexError):
        b.next_set(1000)
    with pytest.raises(ValueError, match="BitSets must have the same size"):
        b.iand(B(999, 10))
    with pytest.raises(TypeError):
        b.iand(bx.bitset.BitSet(1000))
    b = B(10000, 10); b.set_range(0, 10); b.invert()
    assert [b.count_range(1500, 100), b.count_range(1000, 1000), b.count_range(1100, 1900),
            b.count_range(0, 10000)] == [-400, 1000, 1800, 9990]
    assert b.count_ranges([1500], [100], strict=False).tolist() == [100]
    b = B(95, 10); b.set_range(90, 5); b.invert()
    assert [b[i] for i in range(88, 95)] == [1, 1, 0, 0, 0, 0, 0] and b.next_set(94) == 95


def test_bitset_golden_sequences(bx):
    for case in json.load(open(os.path.join(G, "bitset.json"))):
        size, gran, ops, probes = synth.bitset_case(case["seed"])
        b = [bx.bitset.BinnedBitSet(size, gran), bx.bitset.BinnedBitSet(size, gran)]
        for op in ops:
            synth.apply_bitset_op(b, op)
        ps = np.array([s for s, _ in probes], np.int32); pc = np.array([c for _, c in probes], np.int32)
        for k in (0, 1):
            r = case["results"][k]
            assert b[k].bin_size == r["bin_size"]
            assert b[k].count_ranges(ps, pc).tolist() == r["count"], case["seed"]
            assert b[k].get_many(ps).tolist() == r["get"]
            assert [b[k].next_set(s) for s in ps[:12].tolist()] == r["next_set"][:12]
            assert [b[k].next_clear(s) for s in ps[:12].tolist()] == r["next_clear"][:12]


def test_bitset_c3_golden(bx):
    g = np.load(os.path.join(G, "bitset_c3.npz"))
    for tag, nr in (("dense", 4000), ("sparse", 200)):
        size = 2_500_000
        a, b = bx.bitset.BinnedBitSet(size), bx.bitset.BinnedBitSet(size)
        (sa, ca), (sb, cb), (qs, qc) = synth.c3_case(size, nr, 31)
        a.set_ranges(sa, ca); b.set_ranges(sb, cb)
        assert a.count_range(0, size) == int(g[f"{tag}_count_a"])
        assert a.and_count(b) == int(g[f"{tag}_count_and"])
        assert a.count_range(0, size) == int(g[f"{tag}_count_and"])
        assert np.array_equal(a.count_ranges(qs, qc), g[f"{tag}_counts"])
        rs, re = a.runs()
        assert np.array_equal(np.stack([rs, re], 1), g[f"{tag}_runs"])
        a.invert()
        assert np.array_equal(a.count_ranges(qs, qc), g[f"{tag}_inv_counts"])
        assert a.count_range(0, size) == int(g[f"{tag}_inv_total"])


@pytest.mark.parametrize("seed", range(30))
def test_bitset_random_vs_oracle(bx, orc, seed):
    rng = np.random.default_rng(400 + seed)
    size = int(rng.integers(1, 200000))
    gran = int(rng.choice([1, 3, 10, 64, 1024]))
    b = [bx.bitset.BinnedBitSet(size, gran) for _ in range(2)]
    o = [orc.OracleBinnedBitSet(size, gran) for _ in range(2)]
    for _ in range(int(rng.integers(1, 12))):
        k = int(rng.integers(0, 2)); op = int(rng.integers(0, 6))
        if op == 0:
            m = int(rng.integers(1, 200))
            s = rng.integers(0, size, m); c = rng.integers(0, np.minimum(size - s, max(1, size // 50)) + 1)
            b[k].set_ranges(s, c); o[k].set_ranges(s, c)
        elif op == 1:
            p = int(rng.integers(0, size)); b[k].set(p); o[k].set(p)
        elif op == 2:
            p = int(rng.integers(0, size)); b[k].clear(p); o[k].clear(p)
        elif op == 3:
            b[k].invert(); o[k].invert()
        elif op == 4:
            b[k].iand(b[1 - k]); o[k].iand(o[1 - k])
        else:
            b[k].ior(b[1 - k]); o[k].ior(o[1 - k])
    for x, y in zip(b, o):
        assert (x.size, x.bin_size) == (y.size, y.bin_size)
        assert np.array_equal(x.bin_states(), y.states())
        assert np.array_equal(x.to_words(), y.words())
        s = rng.integers(0, size, 500); c = rng.integers(0, size - s + 1)
        assert np.array_equal(x.count_ranges(s, c), y.count_ranges(s, c))
        assert np.array_equal(x.get_many(s), y.read(s))
        rs, re = x.runs(); ors, ore = y.runs()
        assert np.array_equal(rs, ors) and np.array_equal(re, ore)
        for p in s[:10].tolist():
            assert x.next_set(p) == y.next_set(p) and x.next_clear(p) == y.next_clear(p)
        assert x.count_all() == int(np.unpackbits(y.words().view(np.uint8)).sum())


def test_bitset_genome_batch_vs_oracle(bx, orc):
    """iand_many / ior_many / and_count_many over 24 differently sized bitsets == per-pair reference ops."""
    rng = np.random.default_rng(600)
    sizes = [int(x) for x in rng.integers(1, 3_000_000, 24)]
    sizes[3], sizes[7] = 1, 64 * 2048          # degenerate and exactly chunk-aligned
    A, B, OA, OB = [], [], [], []
    for size in sizes:
        gran = int(rng.choice([1, 10, 1024]))
        a, b = bx.bitset.BinnedBitSet(size, gran), bx.bitset.BinnedBitSet(size, gran)
        oa, ob = orc.OracleBinnedBitSet(size, gran), orc.OracleBinnedBitSet(size, gran)
        for x, ox in ((a, oa), (b, ob)):
            m = int(rng.integers(0, 60))
            s = rng.integers(0, size, m); c = rng.integers(0, np.minimum(size - s, max(1, size // 20)) + 1)
            x.set_ranges(s, c); ox.set_ranges(s, c)
            if rng.random() < 0.3:
                x.invert(); ox.invert()
        A.append(a); B.append(b); OA.append(oa); OB.append(ob)
    counts = bx.bitset.and_count_many(A, B)
    for oa, ob in zip(OA, OB):
        oa.iand(ob)
    for k, (a, oa) in enumerate(zip(A, OA)):
        assert np.array_equal(a.to_words(), oa.words()), k
        assert np.array_equal(a.bin_states(), oa.states()), k
        assert counts[k] == int(np.unpackbits(oa.words().view(np.uint8)).sum()), k
    bx.bitset.ior_many(A, B)
    for oa, ob in zip(OA, OB):
        oa.ior(ob)
    for k, (a, oa) in enumerate(zip(A, OA)):
        assert np.array_equal(a.to_words(), oa.words()) and np.array_equal(a.bin_states(), oa.states()), k
    bx.bitset.iand_many(B, A)
    for oa, ob in zip(OA, OB):
        ob.iand(oa)
    for k, (b, ob) in enumerate(zip(B, OB)):
        assert np.array_equal(b.to_words(), ob.words()) and np.array_equal(b.bin_states(), ob.states()), k
    with pytest.raises(ValueError):
        bx.bitset.iand_many([A[0]], [A[1]])


def test_set_ranges_long_and_short_mixed(bx, orc):
    """Chromosome-arm-long ranges (handed to the grid-wide fill) mixed with short ones, odd/even word alignments,
    overlapping each other; scalar set_range of the whole bitmap."""
    size = 60_000_000
    rng = np.random.default_rng(602)
    s = np.concatenate([[3, 100, 64 * 30001, 64 * 30001 + 1, 1_999_999], rng.integers(0, size - 3000, 500)])
    c = np.concatenate([[40_000_000, 2_000_013, 64 * 20000, 64 * 20001 + 7, 30_000_001], rng.integers(0, 3000, 500)])
    b, o = bx.bitset.BinnedBitSet(size), orc.OracleBinnedBitSet(size)
    b.set_ranges(s, c); o.set_ranges(s, c)
    assert np.array_equal(b.to_words(), o.words())
    assert np.array_equal(b.bin_states(), o.states())
    b2, o2 = bx.bitset.BinnedBitSet(size, 10), orc.OracleBinnedBitSet(size, 10)
    b2.set_range(0, size); o2.set_range(0, size)
    assert b2.count_range(0, size) == size == o2.count_range(0, size)
    b2.invert()
    assert b2.count_all() == 0 and b2.next_set(0) == size


def test_set_ranges_many_vs_oracle(bx, orc):
    """bitset.set_ranges_many: a shuffled multi-chromosome batch in one launch == per-set oracle, incl. a very long range,
    binned and flat sets mixed, and the scalar call's IndexError for an offending entry."""
    rng = np.random.default_rng(41)
    sizes = [1_000_003, 250_000, 64, 5_000_000]
    sets = [bx.bitset.BinnedBitSet(sizes[0]), bx.bitset.BinnedBitSet(sizes[1], 7), bx.bitset.BitSet(sizes[2]),
            bx.bitset.BinnedBitSet(sizes[3])]
    oracles = [orc.OracleBinnedBitSet(sizes[0]), orc.OracleBinnedBitSet(sizes[1], 7), orc.OracleBitSet(sizes[2]),
               orc.OracleBinnedBitSet(sizes[3])]
    n = 40_000
    which = rng.integers(0, 4, n).astype(np.int32)
    size_of = np.asarray(sizes)[which]
    start = (rng.random(n) * (size_of - 1)).astype(np.int64)
    count = np.minimum(rng.integers(0, 3000, n), size_of - start)
    start[0], count[0], which[0] = 1000, 4_000_000, 3          # one range far above the warp-sweep limit
    bx.bitset.set_ranges_many(sets, which, start, count)
    for k, (b, o) in enumerate(zip(sets, oracles)):
        sel = which == k
        for s_, c_ in zip(start[sel].tolist(), count[sel].tolist()):
            o.set_range(s_, c_)
        words = o.words() if hasattr(o, "words") else None
        if words is not None:
            assert np.array_equal(b.to_words(), words), f"set {k}"
        else:
            pos = rng.integers(0, sizes[k], 64)
            assert b.get_many(pos).tolist() == [o[int(p)] for p in pos]
        assert b.count_range(0, sizes[k]) == o.count_range(0, sizes[k])
    with pytest.raises(IndexError):
        bx.bitset.set_ranges_many(sets, [2], [60], [10])
    with pytest.raises(IndexError):
        bx.bitset.set_ranges_many(sets, [4], [0], [1])
    bx.bitset.set_ranges_many(sets, [], [], [])


def test_count_ranges_many_vs_oracle(bx, orc):
    """bed_intersect's per-line `bitsets[chrom].count_range(...)` as one launch over several bit sets (strict mode incl.
    inverted sets), with lines naming a chromosome that has no bitset."""
    rng = np.random.default_rng(601)
    sets, osets = [], []
    for k in range(5):
        size = int(rng.integers(1000, 400_000)); gran = int(rng.choice([10, 1024]))
        b, o = bx.bitset.BinnedBitSet(size, gran), orc.OracleBinnedBitSet(size, gran)
        s = rng.integers(0, size, 80); c = rng.integers(0, np.minimum(size - s, 500) + 1)
        b.set_ranges(s, c); o.set_ranges(s, c)
        if k % 2:
            b.invert(); o.invert()
        sets.append(b); osets.append(o)
    n = 20000
    which = rng.integers(-1, 6, n).astype(np.int32)
    start = np.zeros(n, np.int32); count = np.zeros(n, np.int32)
    for k in range(5):
        sel = which == k
        m = int(sel.sum())
        start[sel] = rng.integers(0, sets[k].size, m)
        count[sel] = rng.integers(0, np.minimum(sets[k].size - start[sel], 3000) + 1)
    for strict in (True, False):
        got = bx.bitset.count_ranges_many(sets, which, start, count, strict=strict)
        for k in range(5):
            sel = which == k
            exp = osets[k].count_ranges(start[sel], count[sel]) if strict else sets[k].count_ranges(start[sel], count[sel], strict=False)
            assert np.array_equal(got[sel], exp), (k, strict)
        assert np.all(got[(which < 0) | (which > 4)] == 0)
    with pytest.raises(IndexError):
        bx.bitset.count_ranges_many(sets, [0], [sets[0].size - 1], [5])


@pytest.mark.parametrize("seed", range(6))
def test_flat_bitset_vs_oracle(bx, orc, seed):
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(1, 50000))
    b = [bx.bitset.BitSet(n) for _ in range(2)]
    o = [orc.OracleBitSet(n) for _ in range(2)]
    for _ in range(20):
        k = int(rng.integers(0, 2)); op = int(rng.integers(0, 7))
        if op == 0:
            s = int(rng.integers(0, n)); c = int(rng.integers(0, n - s + 1)); b[k].set_range(s, c); o[k].set_range(s, c)
        elif op == 1:
            p = int(rng.integers(0, n)); b[k].set(p); o[k].set(p)
        elif op == 2:
            p = int(rng.integers(0, n)); b[k].clear(p); o[k].clear(p)
        elif op == 3:
            b[k].invert(); o[k].invert()
        elif op == 4:
            b[k] &= b[1 - k]; o[k].iand(o[1 - k])
        elif op == 5:
            b[k] |= b[1 - k]; o[k].ior(o[1 - k])
        else:
            b[k].ixor(b[1 - k]); o[k].ixor(o[1 - k])
    for x, y in zip(b, o):
        for _ in range(40):
            s = int(rng.integers(0, n)); e = int(rng.integers(s, n + 1))
            assert x.count_range(s, e - s) == y.count_range(s, e - s)
            assert x.next_set(s, e) == y.next_set(s, e)
            assert x.next_clear(s, e) == y.next_clear(s, e)
            assert x[s] == y[s]
    c = b[0].clone()
    assert np.array_equal(c.to_words(), b[0].to_words())


def test_bitset_chromosome_scale_properties(bx, orc):
    """C3 shape at full size for one chromosome pair: De Morgan, popcount identities, run/coverage consistency."""
    size = 250_000_000
    (sa, ca), (sb, cb), (qs, qc) = synth.c3_case(size, 400_000, 1, nq=100_000)
    a, b = bx.bitset.BinnedBitSet(size), bx.bitset.BinnedBitSet(size)
    a.set_ranges(sa, ca); b.set_ranges(sb, cb)
    na, nb = a.count_all(), b.count_all()
    oa = orc.OracleBinnedBitSet(size); oa.set_ranges(sa, ca)
    assert na == oa.count_range(0, size)         # full-size bitmap against the CPU restatement
    u = bx.bitset.BinnedBitSet(size); u.ior(a); u.ior(b)
    nand = a.and_count(b)                        # a := a & b
    assert u.count_all() == na + nb - nand       # inclusion-exclusion
    rs, re = a.runs()
    assert int(np.sum(re.astype(np.int64) - rs)) == nand and np.all(rs[1:] > re[:-1])
    cnt = a.count_ranges(qs, qc)
    assert cnt.min() >= 0 and np.all(cnt <= qc)
    assert a.count_range(0, size) == nand
    a.invert()
    assert a.count_all() == size - nand
    assert np.array_equal(a.count_ranges(qs, qc, strict=False), qc - cnt)


# ---------------------------------------------------------------------------------------------------------------
# callers: bitset_builders / bitset_utils drop-ins against the reference's own outputs
# ---------------------------------------------------------------------------------------------------------------
def test_builders_and_utils_golden(bx):
    from bx_python_b200 import bitset_builders as bb
    from bx_python_b200 import bitset_utils as bu

    def runs(d):
        return {c: [list(r) for r in bu.bits2list(b)] for c, b in d.items()}
    for case in json.load(open(os.path.join(G, "builders.json"))):
        lines, lens = synth.bed_lines(case["seed"])
        g = case["out"]
        noblank = [ln for ln in lines if not ln.isspace()]
        assert runs(bb.binned_bitsets_from_file(lines, lens=lens)) == g["file"]
        assert runs(bb.binned_bitsets_from_file(lines, lens=lens, upstream_pad=25)) == g["file_pad"]
        chr1 = [ln for ln in lines if ln.startswith("chr1\t")]
        assert runs(bb.binned_bitsets_from_file(chr1, lens=lens, upstream_pad=10, downstream_pad=40)) == g["file_pad1"]
        assert runs(bb.binned_bitsets_from_bed_file(lines, lens=lens)) == g["bed"]
        assert runs(bb.binned_bitsets_proximity(noblank, upstream=30, downstream=10)) == g["prox"]
        lst = [ln.split()[:3] for ln in noblank if not ln.startswith("#")]
        assert runs(bb.binned_bitsets_from_list(lst)) == g["list"]
        assert [list(r) for r in bu.bits2list(bb.binned_bitsets_by_chrom(noblank, "chr2"))] == g["by_chrom"]
        ex1, ex2 = synth.exon_lists(case["seed"])
        assert [list(r) for r in bu.bitset_intersect(ex1, ex2)] == g["intersect"]
        assert [list(r) for r in bu.bitset_subtract(ex1, ex2)] == g["subtract"]
        assert [list(r) for r in bu.bitset_complement(ex1)] == g["complement"]
        assert [list(r) for r in bu.bitset_union(ex1 + ex2)] == g["union"]
        bits = bu.list2bits(ex1)
        got = [[list(r) for r in bu.bitset_interval_intersect(bits, a, b)] for a, b in ((0, 5000), (100, 900), (2500, 2600))]
        assert got == g["interval_intersect"]
    # the reference raises on the offending line (here: end < start -> negative count), before touching the device
    with pytest.raises(IndexError, match="must be non-negative"):
        bb.binned_bitsets_from_file(["chr1\t50\t10\n"], lens={"chr1": 1000})
    chroms = np.array([0, 1, 0, 1]); s = np.array([5, 10, 100, 0]); e = np.array([50, 20, 130, 5])
    out = bb.binned_bitsets_from_arrays(chroms, s, e, [1000, 500])
    assert bu.bits2list(out[0]) == [(5, 50), (100, 130)] and bu.bits2list(out[1]) == [(0, 5), (10, 20)]


# ---------------------------------------------------------------------------------------------------------------
# aggregate
# ---------------------------------------------------------------------------------------------------------------
def test_aggregate_golden(bx):
    for case in json.load(open(os.path.join(G, "aggregate.json"))):
        origin, v, ws, we, mask_runs = synth.aggregate_case(case["seed"])
        track = bx.aggregate.ScoreTrack(v, origin)
        mask = None
        if mask_runs is not None:
            mask = bx.bitset.BinnedBitSet()
            for a, b in mask_runs:
                mask.set_range(a, b - a)
        res = track.aggregate(ws, we, mask)
        got = [bx.aggregate.format_line(res, w) for w in range(len(ws))]
        assert got == case["lines"], case["seed"]


def test_aggregate_survey_probe(bx):
    v = np.array([0.1, 0.2, 0.0, 0.3, np.nan, 1e-3, 16777216, 1, 1], np.float32)
    res = bx.aggregate.ScoreTrack(v, 0).aggregate([0, 0, 2], [4, 9, 3])
    f = bx.aggregate.format_line
    assert f(res, 0) == ["0.2", "0.1", "0.3"]
    assert f(res, 1) == ["2.3967452e+06", "0.001", "1.6777216e+07"]
    assert f(res, 2) == ["nan", "nan", "nan"]


def test_aggregate_genome_vs_oracle(bx, orc):
    """One launch over several chromosomes (track ids per window) == per-chromosome oracle, incl. a masked track,
    an unmasked one, and windows naming a chromosome that has no scores."""
    rng = np.random.default_rng(56)
    tracks, masks, dense, mwords = [], [], [], []
    for t in range(4):
        n = int(rng.integers(10_000, 200_000)); origin = int(rng.integers(0, 5000))
        v = synth.aggregate_scores(rng, n)
        tracks.append(bx.aggregate.ScoreTrack(v, origin))
        d = np.full(origin + n, np.nan, np.float32); d[origin:] = v
        dense.append(d)
        if t % 2 == 0:
            m = bx.bitset.BinnedBitSet(origin + n + 100)
            ms = rng.integers(origin, origin + n, 300); mc = rng.integers(1, 40, 300)
            m.set_ranges(ms, mc)
            masks.append(m); mwords.append(m.to_words())
        else:
            masks.append(None); mwords.append(None)
    nw = 50_000
    wt = rng.integers(-1, 5, nw).astype(np.int32)            # -1 and 4 name tracks that do not exist
    ws = np.empty(nw, np.int32); we = np.empty(nw, np.int32)
    for t in range(-1, 5):
        sel = wt == t
        hi = len(dense[t]) if 0 <= t < 4 else 1000
        ws[sel] = rng.integers(-20, hi + 20, int(sel.sum()))
        we[sel] = ws[sel] + rng.integers(0, 70, int(sel.sum()))
    res = bx.aggregate.aggregate_genome(tracks, wt, ws, we, masks)
    for t in range(4):
        sel = np.nonzero(wt == t)[0]
        o = orc.aggregate(dense[t], ws[sel], we[sel], mwords[t])
        for k in ("sum", "avg", "min", "max"):
            assert np.array_equal(res[k][sel].view(np.uint32), o[k].view(np.uint32)), (t, k)
        assert np.array_equal(res["count"][sel], o["count"])
    none = (wt < 0) | (wt > 3)
    assert np.all(res["count"][none] == 0) and np.all(np.isnan(res["avg"][none]))


def test_aggregate_random_vs_oracle(bx, orc):
    rng = np.random.default_rng(55)
    n, nw = 2_000_000, 200_000
    origin = 12345
    v = synth.aggregate_scores(rng, n)
    ws = rng.integers(origin - 50, origin + n + 20, nw).astype(np.int32)
    we = (ws + rng.integers(0, 80, nw)).astype(np.int32)
    mask = bx.bitset.BinnedBitSet(origin + n + 200)
    ms = rng.integers(origin, origin + n, 5000); mc = rng.integers(1, 30, 5000)
    mask.set_ranges(ms, mc)
    res = bx.aggregate.ScoreTrack(v, origin).aggregate(ws, we, mask)
    dense = np.full(origin + n, np.nan, np.float32); dense[origin:] = v
    ores = orc.aggregate(dense, ws, we, mask.to_words())
    for k in ("sum", "avg", "min", "max"):
        assert np.array_equal(res[k].view(np.uint32), ores[k].view(np.uint32)), k
    assert np.array_equal(res["count"], ores["count"])
