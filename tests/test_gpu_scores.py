"""
GPU parity for the SURVEY 8f-4 rows: device BinnedArray / FileBinnedArray, the wiggle loader, the script-level
aggregate, bigWig summaries and the overlap join -- through the Python shim -> ctypes -> C ABI of libbxb200.so --
against the golden vectors of the unmodified reference (scores.npz, summarize.npz, join.json, aggregate.json) and
against the CPU restatement on seeded random inputs.  Bit-exact: float32 / float64 results are compared as bit
patterns.
"""
import hashlib
import io
import json
import os
import sys

import numpy as np
import pytest

from bx_python_b200 import synth

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, G)
from make_golden import canonical_join_rows  # noqa: E402

KEYS = ("valid_count", "min_val", "max_val", "sum_data", "sum_squares")


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="module", autouse=True)
def device():
    from bx_python_b200 import _lib
    _lib.lib()                     # raises without a CUDA device: these tests never pass on a fallback


# ---------------------------------------------------------------------------------------------------------------
# BinnedArray / wiggle
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(8))
def test_wiggle_load_golden(seed):
    from bx_python_b200 import wiggle
    from bx_python_b200.binned_array import BinnedArray
    g = np.load(os.path.join(G, "scores.npz"))
    text = synth.wiggle_text(seed)
    # (a) the batched loader
    spans = wiggle.read_spans(io.StringIO(text))
    for chrom in g[f"s{seed}_chroms"].tolist():
        ba = BinnedArray(bin_size=1024, max_size=8192)
        ba.set_spans(*spans[chrom])
        assert np.array_equal(bits(ba.get_range(0, 8192)), bits(g[f"s{seed}_{chrom}_dense"]))
        buf = io.BytesIO()
        ba.to_file(buf)
        assert hashlib.sha256(buf.getvalue()).digest() == g[f"s{seed}_{chrom}_file_sha256"].tobytes()
    # (b) the reference's own per-base loop through the scalar API (queued sets, flushed in order)
    arrays = {}
    for chrom, pos, val in wiggle.Reader(io.StringIO(text)):
        if chrom not in arrays:
            arrays[chrom] = BinnedArray(bin_size=1024, max_size=8192)
        arrays[chrom][pos] = val
    assert list(arrays) == g[f"s{seed}_chroms"].tolist()
    for chrom, ba in arrays.items():
        assert np.array_equal(bits(ba[0:8192]), bits(g[f"s{seed}_{chrom}_dense"]))
    # (c) load_scores_wiggle with the default geometry
    loaded = wiggle.load_scores_wiggle(io.StringIO(text))
    assert list(loaded) == g[f"s{seed}_chroms"].tolist()
    for chrom, ba in loaded.items():
        assert np.array_equal(bits(ba.get_range(0, 8192)), bits(g[f"s{seed}_{chrom}_dense"]))


def test_binned_array_reference_unit_tests():
    """The shape of lib/bx/binned_array_tests.py:27-63 (per-index and slice reads of an array filled through
    __setitem__, bin size 128), plus the type and index quirks of BinnedArray.get (binned_array.py:89-94)."""
    from bx_python_b200.binned_array import BinnedArray
    rng = np.random.default_rng(11)
    CHUNK, SIZE = 1000, 10000
    source = np.full(SIZE, np.nan, np.float32)
    target = BinnedArray(128, np.nan, SIZE)
    for i in range(0, SIZE, CHUNK * 2):
        vals = rng.random(CHUNK).astype(np.float32)
        source[i:i + CHUNK] = vals
        for j in range(CHUNK):
            target[i + j] = vals[j]
    for i in range(0, SIZE, 97):
        a, b = source[i], target[i]
        assert (np.isnan(a) and np.isnan(b)) or a == b
    for start in range(0, SIZE - 51, 331):
        assert np.array_equal(bits(source[start:start + 51]), bits(target[start:start + 51]))
    assert target.nbins == 79 and target.get_bin_offset(300) == (2, 44)
    untouched = BinnedArray(128, -1.0, SIZE)
    assert untouched[5] == -1.0 and isinstance(untouched[5], float)          # default object for a missing bin
    untouched[130] = 2.5
    assert isinstance(untouched[129], np.float32) and untouched[129] == np.float32(-1.0)
    assert untouched[-1] == -1.0                                              # negative index wraps to the last bin
    with pytest.raises(IndexError):
        untouched[128 * 79]
    with pytest.raises(AssertionError):
        untouched[0:10:2]


def test_binned_array_file_roundtrip():
    from bx_python_b200.binned_array import BinnedArray, FileBinnedArray
    rng = np.random.default_rng(12)
    ba = BinnedArray(bin_size=512, max_size=20000)
    pos = rng.integers(0, 20000, 3000)
    val = rng.normal(size=3000).astype(np.float32)
    ba.set_many(pos, val)
    dense = np.full(20480, np.nan, np.float32)
    for p, v in zip(pos.tolist(), val.tolist()):        # sequential semantics: later duplicates win
        dense[p] = v
    assert np.array_equal(bits(ba.get_range(0, 20000)), bits(dense[:20000]))
    for comp in ("zlib", "none"):
        buf = io.BytesIO()
        ba.to_file(buf, comp_type=comp)
        buf.seek(0)
        fba = FileBinnedArray(buf)
        assert (fba.max_size, fba.bin_size, fba.nbins) == (20000, 512, 40)
        assert np.array_equal(bits(fba.get_range(0, 20000)), bits(dense[:20000]))
        assert np.array_equal(bits(fba.get_many(pos[:100])), bits(dense[pos[:100]]))
        assert np.isnan(fba.default)


@pytest.mark.parametrize("seed", range(6))
def test_set_spans_random_vs_oracle(orc, seed):
    """Unsorted, overlapping, empty and long spans: the device owner/apply path against the sequential loop."""
    from bx_python_b200.binned_array import BinnedArray
    rng = np.random.default_rng(300 + seed)
    n = 200_000
    k = 20_000
    s = rng.integers(0, n - 3000, k)
    ln = rng.integers(-2, 40, k)
    ln[rng.integers(0, k, 20)] = rng.integers(500, 3000, 20)           # a few long ones (warp-cooperative sweep)
    if seed % 2 == 0:                                                 # sorted + disjoint: the direct path
        ln = np.abs(ln) + 1
        s = np.cumsum(ln + rng.integers(0, 3, k)) - ln
        n = int(s[-1] + ln[-1]) + 10
    e = s + ln
    v = rng.normal(size=k).astype(np.float32)
    ba = BinnedArray(bin_size=4096, max_size=1 << 22)
    ba.set_spans(s, e, v)
    track = np.full(n, np.nan, np.float32)
    orc.scores_set_spans(track, 0, s, e, v)
    assert np.array_equal(bits(ba.get_range(0, n)), bits(track))
    # a second batch on top of the first
    s2 = rng.integers(0, n - 50, 5000)
    e2 = s2 + rng.integers(0, 50, 5000)
    v2 = rng.normal(size=5000).astype(np.float32)
    ba.set_spans(s2, e2, v2)
    orc.scores_set_spans(track, 0, s2, e2, v2)
    assert np.array_equal(bits(ba.get_range(0, n)), bits(track))
    q = rng.integers(0, n, 1000)
    assert np.array_equal(bits(ba.get_many(q)), bits(track[q]))


def test_set_spans_errors():
    from bx_python_b200.binned_array import BinnedArray
    ba = BinnedArray(bin_size=100, max_size=1000)
    with pytest.raises(IndexError):
        ba.set_spans([990], [1001], [1.0])
    with pytest.raises(IndexError):
        ba.set_many([-1], [1.0])
    ba.set_spans([5, 7], [5, 3], [1.0, 2.0])                          # empty spans assign nothing
    assert not ba._allocated.any()


# ---------------------------------------------------------------------------------------------------------------
# aggregate script through the new sources
# ---------------------------------------------------------------------------------------------------------------
def test_aggregate_script_golden():
    """aggregate.json holds the lines printed by the reference script's main(); reproduce them from the same
    wiggle / BED / mask text through load_scores_wiggle + binned_bitsets_from_file + one aggregate launch."""
    from bx_python_b200 import aggregate, wiggle
    from bx_python_b200.bitset_builders import binned_bitsets_from_file
    for case in json.load(open(os.path.join(G, "aggregate.json"))):
        origin, scores, ws, we, mask_runs = synth.aggregate_case(case["seed"])
        wig = [f"fixedStep chrom=chr1 start={origin + 1} step=1\n"]
        wig += ["nan\n" if v != v else repr(float(v)) + "\n" for v in scores]
        bed = [f"chr1\t{a}\t{b}\n" for a, b in zip(ws.tolist(), we.tolist())]
        bed.append("chrUn\t5\t50\n")                                  # a chromosome without scores
        masks = None
        if mask_runs is not None:
            masks = binned_bitsets_from_file([f"chr1\t{a}\t{b}\n" for a, b in mask_runs])
        out = io.StringIO()
        aggregate.aggregate_scores_in_intervals(wiggle.load_scores_wiggle(io.StringIO("".join(wig))), bed, out, masks)
        lines = [ln.split("\t") for ln in out.getvalue().splitlines()]
        assert [ln[3:] for ln in lines[:-1]] == case["lines"]
        assert [ln[:3] for ln in lines[:-1]] == [["chr1", str(a), str(b)] for a, b in zip(ws.tolist(), we.tolist())]
        assert lines[-1] == ["chrUn", "5", "50", "nan", "nan", "nan"]


def test_aggregate_nonnan_default_counts_outside_the_track():
    """A BinnedArray whose default is neither NaN nor 0 contributes that default for positions nobody set -- also
    beyond the highest written bin (binned_array.py:89-94 feeding the script's loop, :113-124)."""
    from bx_python_b200 import aggregate
    from bx_python_b200.binned_array import BinnedArray
    ba = BinnedArray(bin_size=16, default=2.0, max_size=256)
    ba.set_many([3, 4], [5.0, 0.0])
    res = aggregate.aggregate_genome([ba], [0, 0], [0, 100], [8, 104])
    # window [0,8): 2,2,2,5,(0 skipped),2,2,2 ; window [100,104): four defaults beyond the device track
    assert res["count"].tolist() == [7, 4]
    assert res["sum"].tolist() == [17.0, 8.0] and res["min"].tolist() == [2.0, 2.0] and res["max"].tolist() == [5.0, 2.0]
    nan_default = BinnedArray(bin_size=16, max_size=256)
    nan_default.set_many([3], [5.0])
    res = aggregate.aggregate_genome([nan_default], [0, 0], [0, 100], [8, 104])
    assert res["count"].tolist() == [1, 0]


# ---------------------------------------------------------------------------------------------------------------
# bigWig summary
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(24))
def test_summarize_golden(seed):
    from bx_python_b200.bbi import SummarizedData
    g = np.load(os.path.join(G, "summarize.npz"))
    s, e, v, rs, re_, size = synth.summarize_case(seed)
    sd = SummarizedData(rs, re_, size)
    if seed % 4 < 2:
        sd.min_val[:] = np.inf
        sd.max_val[:] = -np.inf
    half = len(s) // 2                                                # two calls continue one summary exactly
    sd.accumulate_intervals(s[:half], e[:half], v[:half])
    sd.accumulate_intervals(s[half:], e[half:], v[half:])
    assert np.array_equal(bits(np.stack([getattr(sd, k) for k in KEYS])), bits(g[f"c{seed}"]))


def test_summarize_bigwig_file_golden():
    from bx_python_b200 import bbi
    g = np.load(os.path.join(G, "summarize.npz"))
    for k, (a, b, size) in enumerate(g["bw_regions"].tolist()):
        sd = bbi.summarize_from_full(g["bw_start"], g["bw_end"], g["bw_val"], a, b, size)
        assert np.array_equal(bits(np.stack([getattr(sd, key) for key in KEYS])), bits(g[f"bw{k}"]))
    # lib/bx/bbi/bigwig_tests.py:73-92 known answers (test_get_leaf: chr1:11000-11005 is answered from the full data;
    # the 10-bin test above it goes through a zoom level of the file, which is container code and out of scope)
    q = bbi.query(bbi.summarize_from_full(g["bw_start"], g["bw_end"], g["bw_val"], 11000, 11005, 5), 11000, 11005, 5)
    assert np.allclose([float(x["mean"]) for x in q],
                       [0.050842501223087311, -2.4589500427246094, 0.050842501223087311, 0.050842501223087311,
                        0.050842501223087311])
    one = bbi.query(bbi.summarize_from_full(g["bw_start"], g["bw_end"], g["bw_val"], 11000, 11005, 1), 11000, 11005, 1)
    assert [float(x["max"]) for x in one] == [0.050842501223087311]
    assert [float(x["min"]) for x in one] == [-2.4589500427246094]
    whole = bbi.summarize_from_full(g["bw_start"], g["bw_end"], g["bw_val"], 10000, 20000, 1)    # :62-67 min / max
    assert whole.max_val.tolist() == [0.289000004529953] and whole.min_val.tolist() == [-3.9100000858306885]
    assert bbi.summarize_from_full([], [], [], 5, 5, 3) is None


def test_summarize_large_vs_oracle(orc):
    from bx_python_b200.bbi import SummarizedData
    rng = np.random.default_rng(77)
    n = 200_000
    ln = rng.integers(1, 30, n)
    s = np.cumsum(ln + rng.integers(0, 5, n)) - ln
    e = s + ln
    v = rng.normal(size=n).astype(np.float32)
    rs, re_, size = 1000, int(e[-1]) - 500, 1000
    sd = SummarizedData(rs, re_, size).accumulate_intervals(s, e, v)
    o = orc.summarize(s, e, v, rs, re_, size, 0.0, 0.0)
    assert all(np.array_equal(bits(getattr(sd, k)), bits(o[k])) for k in KEYS)


# ---------------------------------------------------------------------------------------------------------------
# join
# ---------------------------------------------------------------------------------------------------------------
def test_join_golden():
    from bx_python_b200.intervals.operations import join as J
    for case in json.load(open(os.path.join(G, "join.json"))):
        left, right, mincols = synth.join_case(case["seed"])
        rows = list(J.join(J.BedRows(left), J.BedRows(right), mincols=mincols, leftfill=case["leftfill"],
                           rightfill=case["rightfill"]))
        assert canonical_join_rows(rows, 4) == case["rows"]


@pytest.mark.parametrize("seed", range(4))
def test_join_arrays_vs_oracle(orc, seed):
    from bx_python_b200.intervals.operations.join import join_arrays
    from bx_python_b200.intervals.operations.quicksect import IntervalTree
    rng = np.random.default_rng(500 + seed)
    n, nq, G_ = 3000, 4000, 20000
    chroms = [f"c{int(x)}" for x in rng.integers(0, 4, n)]
    s = rng.integers(0, G_, n)
    e = s + rng.integers(0, 300, n)
    t = IntervalTree()
    t.insert_many(chroms, s, e)
    qc = [f"c{int(x)}" for x in rng.integers(0, 5, nq)]              # c4 is unknown to the tree
    qs = rng.integers(0, G_, nq)
    qe = qs + rng.integers(0, 300, nq)
    for mincols in (1, 25, 400):
        off, items, vis = join_arrays(t, qc, qs, qe, mincols)
        tid = np.asarray([t.chroms[c] for c in chroms], np.int32)
        qt = np.asarray([t.chroms.get(c, -1) for c in qc], np.int32)
        ooff, oitems, ovis = orc.join(tid, s, e, qt, qs, qe, mincols)
        assert np.array_equal(off, ooff) and np.array_equal(vis, ovis)
        for q in range(nq):                                          # same set per left interval (order: index vs id)
            assert sorted(items[off[q]:off[q + 1]].tolist()) == oitems[ooff[q]:ooff[q + 1]].tolist()


def test_quicksect_intersect_and_empty():
    from bx_python_b200.intervals.operations.quicksect import IntervalTree

    class IV:
        def __init__(self, c, s, e):
            self.chrom, self.start, self.end = c, s, e
    t = IntervalTree()
    got = []
    t.intersect(IV("chr1", 0, 10), got.append)
    assert got == []
    for k, (s, e) in enumerate([(0, 10), (5, 15), (20, 30), (10, 20)]):
        t.insert(IV("chr1", s, e), linenum=k, other=[k])
    t.intersect(IV("chr1", 9, 11), got.append)
    assert sorted(n.linenum for n in got) == [0, 1, 3]
    got.clear()
    t.intersect(IV("chr2", 9, 11), got.append)
    assert got == []
    t.insert(IV("chr2", 9, 11), linenum=9)                            # insert after a query rebuilds the index
    t.intersect(IV("chr2", 10, 12), got.append)
    assert [n.linenum for n in got] == [9]
