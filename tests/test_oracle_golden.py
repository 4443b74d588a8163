"""
The CPU restatement (oracle/bx_oracle.c) against the committed golden vectors (tests/golden/, produced from the
compiled unmodified reference by tests/golden/make_golden.py), the reference's own unit-test known answers
(lib/bx/bitset_tests.py:51-108, lib/bx/intervals/intersection_tests.py:158-201, doctests intersection.pyx:341-376)
and the edge probes recorded in SURVEY.md 8(a) addendum 2.  CPU only.
"""
import json
import os

import numpy as np
import pytest

from bx_python_b200 import synth
from oracle import oracle as orc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_find_c1_and_edges():
    g = np.load(os.path.join(G, "find.npz"))
    s, e, qs, qe = synth.c1_intervals()
    t = orc.OracleIntervalTree(s, e)
    off, hits = t.find(qs, qe)
    assert np.array_equal(off, g["c1_offsets"]) and np.array_equal(hits, g["c1_hits"])
    assert np.array_equal(t.order(), g["c1_order"])
    for k, (s, e, qs, qe) in enumerate(synth.edge_sets()):
        t = orc.OracleIntervalTree(s, e)
        off, hits = t.find(qs, qe)
        assert np.array_equal(off, g[f"edge{k}_offsets"]) and np.array_equal(hits, g[f"edge{k}_hits"])
        assert np.array_equal(t.order(), g[f"edge{k}_order"])


def test_find_survey_probes():
    # SURVEY 8(a) addendum 2, probed on the compiled reference
    data = [(10, 20), (15, 12), (30, 30), (-5, 3), (20, 25)]
    t = orc.OracleIntervalTree([a for a, _ in data], [b for _, b in data])

    def f(a, b):
        off, h = t.find([a], [b])
        return h.tolist()
    assert f(-100, 100) == [3, 0, 1, 4, 2]
    assert f(18, 11) == [0]
    assert f(20, 20) == []
    assert f(19, 21) == [0, 4]
    assert f(30, 30) == []
    assert f(29, 31) == [2]
    # equal-start ordering probe: a b z0 c z1 d e y -> d y z1 z0 a b c e
    names = "a b z0 c z1 d e y".split()
    iv = [(5, 10), (5, 7), (5, 5), (5, 20), (5, 5), (3, 6), (5, 6), (4, 4)]
    t2 = orc.OracleIntervalTree([a for a, _ in iv], [b for _, b in iv])
    _, h = t2.find([0], [100])
    assert [names[i] for i in h] == "d y z1 z0 a b c e".split()


def test_find_reference_doctest_and_unit_answers():
    # intersection.pyx:355-361 doctest: find(30,50) -> [Interval(3,40), Interval(13,50)] ; find(100,200) -> []
    iv = [(0, 10), (3, 7), (3, 40), (13, 50)]
    t = orc.OracleIntervalTree([a for a, _ in iv], [b for _, b in iv])
    assert t.find([30], [50])[1].tolist() == [2, 3]
    assert t.find([100], [200])[1].tolist() == []
    # before_interval(Interval(10,20)) -> [Interval(3,7)] ; before_interval(Interval(5,20)) -> []   (:365-368)
    assert t.before(10).tolist() == [1]
    assert t.before(5).tolist() == []
    # upstream_of_interval(Interval(11,12)) -> [Interval(0,10)] ; strand "-" -> after(12) -> [Interval(13,50)] (:372-375)
    assert t.before(11).tolist() == [0]
    assert t.after(12).tolist() == [3]
    # upstream_of_interval(Interval(1,2,strand="-"), num_intervals=3) -> [(3,7),(3,40),(13,50)]   (:377-378)
    assert t.after(2, 3).tolist() == [1, 2, 3]
    # intersection_tests.py:158-177 IntervalTreeTest.test_find: find(100,200) has 5 hits
    s, e = [], []
    for i in range(1, 1000, 80):
        for d in (0, 20, 40, 60):
            s.append(i + d); e.append(i + d + 10)
    assert len(orc.OracleIntervalTree(s, e).find([100], [200])[1]) == 5


def test_neighbors_golden():
    for case in json.load(open(os.path.join(G, "neighbors.json"))):
        s, e, queries = synth.neighbor_case(case["seed"])
        t = orc.OracleIntervalTree(s, e)
        for (pos, k, md), (b, a) in zip(queries, case["results"]):
            assert t.before(pos, k, md).tolist() == b
            assert t.after(pos, k, md).tolist() == a


@pytest.mark.parametrize("cls", ["flat", "binned"])
def test_bitset_reference_unit_answers(cls):
    # lib/bx/bitset_tests.py:51-77 (granularity = size % 11 -> 1 for size 100, :116-119)
    def new():
        return orc.OracleBitSet(100) if cls == "flat" else orc.OracleBinnedBitSet(100, 100 % 11)
    b = new()
    for s, e in ((11, 14), (20, 75), (90, 100)):
        b.set_range(s, e - s)
    assert [b.count_range(0, 0), b.count_range(0, 20), b.count_range(25, 25), b.count_range(80, 20),
            b.count_range(0, 100)] == [0, 3, 25, 10, 68]
    assert [b.next_set(0), b.next_set(13), b.next_set(15)] == [11, 13, 20]
    assert [b.next_clear(0), b.next_clear(11), b.next_clear(20), b.next_clear(92)] == [0, 14, 75, 100]
    # :79-108 and / or / not
    b1, b2 = new(), new()
    b1.set_range(20, 40); b2.set_range(50, 25); b1.iand(b2)
    assert [b1[i] for i in range(100)] == [1 if 50 <= i < 60 else 0 for i in range(100)]
    b1, b2 = new(), new()
    b1.set_range(20, 40); b2.set_range(50, 25); b1.ior(b2)
    assert [b1[i] for i in range(100)] == [1 if 20 <= i < 75 else 0 for i in range(100)]
    b1 = new(); b1.set_range(20, 40); b1.invert()
    assert [b1[i] for i in range(100)] == [0 if 20 <= i < 60 else 1 for i in range(100)]


def test_bitset_survey_probes():
    geo = {(2**29, 1024): 524288, (250000000, 1024): 244141, (248956422, 1024): 243122, (100, 1): 100,
           (100, 3): 34, (100, 1024): 1, (1000, 10): 100, (1000, 20): 50}
    for (size, gran), bin_size in geo.items():
        assert orc.OracleBinnedBitSet(size, gran).bin_size == bin_size
    b = orc.OracleBinnedBitSet(10000, 10)
    b.set_range(0, 10); b.invert()
    assert [b.count_range(1500, 100), b.count_range(1000, 1000), b.count_range(1100, 1900),
            b.count_range(0, 10000)] == [-400, 1000, 1800, 9990]
    b = orc.OracleBinnedBitSet(95, 10)
    b.set_range(90, 5); b.invert()
    assert [b[i] for i in range(88, 95)] == [1, 1, 0, 0, 0, 0, 0] and b.next_set(94) == 95


def test_bitset_golden_sequences():
    for case in json.load(open(os.path.join(G, "bitset.json"))):
        size, gran, ops, probes = synth.bitset_case(case["seed"])
        b = [orc.OracleBinnedBitSet(size, gran), orc.OracleBinnedBitSet(size, gran)]
        for op in ops:
            synth.apply_bitset_op(b, op)
        for k in (0, 1):
            r = case["results"][k]
            assert b[k].bin_size == r["bin_size"]
            assert [b[k].count_range(s, c) for s, c in probes] == r["count"]
            assert [b[k].next_set(s) for s, _ in probes] == r["next_set"]
            assert [b[k].next_clear(s) for s, _ in probes] == r["next_clear"]
            assert [b[k][s] for s, _ in probes] == r["get"]


def test_bitset_c3_golden():
    g = np.load(os.path.join(G, "bitset_c3.npz"))
    for tag, nr in (("dense", 4000), ("sparse", 200)):
        size = 2_500_000
        a, b = orc.OracleBinnedBitSet(size), orc.OracleBinnedBitSet(size)
        (sa, ca), (sb, cb), (qs, qc) = synth.c3_case(size, nr, 31)
        a.set_ranges(sa, ca); b.set_ranges(sb, cb)
        assert a.count_range(0, size) == int(g[f"{tag}_count_a"])
        a.iand(b)
        assert a.count_range(0, size) == int(g[f"{tag}_count_and"])
        assert np.array_equal(a.count_ranges(qs, qc), g[f"{tag}_counts"])
        rs, re = a.runs()
        assert np.array_equal(np.stack([rs, re], 1), g[f"{tag}_runs"])
        a.invert()
        assert np.array_equal(a.count_ranges(qs, qc), g[f"{tag}_inv_counts"])
        assert a.count_range(0, size) == int(g[f"{tag}_inv_total"])


def fmt_aggregate(res, w):
    """Format one window the way scripts/aggregate_scores_in_intervals.py:126-134 prints it."""
    if res["count"][w] == 0:
        return ["nan", "nan", "nan"]
    return [str(np.float32(res[k][w])) for k in ("avg", "min", "max")]


def aggregate_inputs(seed):
    origin, v, ws, we, mask_runs = synth.aggregate_case(seed)
    n = origin + len(v)
    dense = np.full(n, np.nan, np.float32)
    dense[origin:] = v
    mask = None
    if mask_runs is not None:
        m = orc.OracleBinnedBitSet(max(n, int(we.max()) + 1), 1)
        for a, b in mask_runs:
            m.set_range(a, b - a)
        mask = m.words()
    return dense, ws, we, mask


def test_aggregate_golden():
    for case in json.load(open(os.path.join(G, "aggregate.json"))):
        dense, ws, we, mask = aggregate_inputs(case["seed"])
        res = orc.aggregate(dense, ws, we, mask)
        got = [fmt_aggregate(res, w) for w in range(len(ws))]
        assert got == case["lines"], case["seed"]


def test_aggregate_survey_probe():
    # SURVEY 8(a) addendum 2: hand wiggle at chr1:0-9
    v = np.array([0.1, 0.2, 0.0, 0.3, np.nan, 1e-3, 16777216, 1, 1], np.float32)
    res = orc.aggregate(v, [0, 0, 2], [4, 9, 3])
    assert fmt_aggregate(res, 0) == ["0.2", "0.1", "0.3"]
    assert fmt_aggregate(res, 1) == ["2.3967452e+06", "0.001", "1.6777216e+07"]
    assert fmt_aggregate(res, 2) == ["nan", "nan", "nan"]
