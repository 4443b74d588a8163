"""
SURVEY 8f-4 rows on the CPU: the restatements in oracle/bx_oracle.c (orc_scores_set_spans, orc_summarize, orc_join)
and the host-side text logic (bx_python_b200.wiggle, the row assembly of operations.join) against the golden vectors
produced from the unmodified reference by tests/golden/make_golden.py (scores.npz: wiggle.Reader -> BinnedArray
per-base loop; summarize.npz: SummarizedData.accumulate_interval_value and BigWigFile.summarize_from_full on the
reference's test.bw; join.json: operations.join.join through the reference's own readers).  No device is touched:
where the product would call the CUDA library the test substitutes the oracle, to check the host logic around it.
"""
import io
import json
import os
import sys

import numpy as np
import pytest

from bx_python_b200 import synth, wiggle
from oracle import oracle as orc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, G)
from make_golden import canonical_join_rows  # noqa: E402  (pure helper; importing does not touch the reference)

KEYS = ("valid_count", "min_val", "max_val", "sum_data", "sum_squares")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32 if a.dtype == np.float32 else np.uint64)


@pytest.mark.parametrize("seed", range(8))
def test_wiggle_reader_and_span_loop_golden(seed):
    g = np.load(os.path.join(G, "scores.npz"))
    text = synth.wiggle_text(seed)
    recs = list(wiggle.IntervalReader(io.StringIO(text)))
    assert len(recs) == int(g[f"s{seed}_nrecords"])
    assert sum((s * 31 + e * 17) % 1000003 for _, s, e, _, _ in recs) == int(g[f"s{seed}_rec_checksum"])
    # Reader = one tuple per covered base, in file order
    per_base = list(wiggle.Reader(io.StringIO(text)))
    assert len(per_base) == sum(max(e - s, 0) for _, s, e, _, _ in recs)
    spans = wiggle.read_spans(io.StringIO(text))
    chroms = [c for c, (s, e, _) in spans.items() if np.any(e > s)]
    assert chroms == g[f"s{seed}_chroms"].tolist()
    for c in chroms:
        s, e, v = spans[c]
        track = np.full(8192, np.nan, np.float32)
        orc.scores_set_spans(track, 0, s, e, v)
        assert np.array_equal(bits(track), bits(g[f"s{seed}_{c}_dense"]))


def test_wiggle_reader_modes():
    text = "chr1\t10\t20\t1.5\tx\t-\nchr1\t1\t2\n\nvariableStep chrom=chr2 span=3\n5 2.5\nfixedStep chrom=chr3 start=11 step=4\n1\n2\n"
    assert list(wiggle.IntervalReader(io.StringIO(text))) == [
        ("chr1", 10, 20, "-", 1.5), ("chr2", 4, 7, "+", 2.5), ("chr3", 10, 11, "+", 1.0), ("chr3", 14, 15, "+", 2.0)]
    assert list(wiggle.Reader(io.StringIO("variableStep chrom=c span=2\n3 7\n"))) == [("c", 2, 7.0), ("c", 3, 7.0)]


@pytest.mark.parametrize("seed", range(24))
def test_summarize_oracle_golden(seed):
    g = np.load(os.path.join(G, "summarize.npz"))
    s, e, v, rs, re_, size = synth.summarize_case(seed)
    init = (np.inf, -np.inf) if seed % 4 < 2 else (0.0, 0.0)
    o = orc.summarize(s, e, v, rs, re_, size, *init)
    assert np.array_equal(bits(np.stack([o[k] for k in KEYS])), bits(g[f"c{seed}"]))


def test_summarize_oracle_bigwig_file_golden():
    g = np.load(os.path.join(G, "summarize.npz"))
    for k, (a, b, size) in enumerate(g["bw_regions"].tolist()):
        o = orc.summarize(g["bw_start"], g["bw_end"], g["bw_val"], a, b, size)
        o["valid_count"] = np.round(o["valid_count"])               # bigwig_file.pyx:182-184
        assert np.array_equal(bits(np.stack([o[key] for key in KEYS])), bits(g[f"bw{k}"]))


def oracle_join_arrays(tree, chroms, starts, ends, mincols=1):
    """Drop-in for operations.join.join_arrays that asks the oracle instead of the device."""
    tid, s, e = tree._arrays()
    qt = np.asarray([tree.chroms.get(c, -1) for c in chroms], np.int32)
    return orc.join(tid, s, e, qt, starts, ends, mincols)


def test_join_host_logic_golden(monkeypatch):
    from bx_python_b200.intervals.operations import join as J
    monkeypatch.setattr(J, "join_arrays", oracle_join_arrays)
    for case in json.load(open(os.path.join(G, "join.json"))):
        left, right, mincols = synth.join_case(case["seed"])
        assert mincols == case["mincols"]
        rows = list(J.join(J.BedRows(left), J.BedRows(right), mincols=mincols, leftfill=case["leftfill"],
                           rightfill=case["rightfill"]))
        assert canonical_join_rows(rows, 4) == case["rows"]


def test_quicksect_traverse_order():
    from bx_python_b200.intervals.operations.quicksect import IntervalTree

    class IV:
        def __init__(self, c, s, e):
            self.chrom, self.start, self.end = c, s, e
    t = IntervalTree()
    for k, (c, s, e) in enumerate([("b", 5, 9), ("a", 7, 8), ("b", 5, 6), ("b", 1, 2), ("a", 7, 9), ("b", 5, 7)]):
        t.insert(IV(c, s, e), linenum=k)
    seen = []
    t.traverse(lambda n: seen.append(n.linenum))
    # chromosomes in order of first insert; by start; among equal starts the later insert first (quicksect.py:52-70)
    assert seen == [3, 5, 2, 0, 4, 1]


def test_wiggle_reference_known_answers():
    """The known answers of lib/bx/wiggle_tests.py:44-92 (UCSC's three-format example: a bedGraph section, a
    variableStep block with span=4, a fixedStep block with step=300 span=3, browser / track / comment lines between)."""
    text = "\n".join([
        "browser position chr19:59302001-59311000", "browser hide all", "#\tcomment",
        'track type=wiggle_0 name="Bed Format" priority=20',
        "chr19 59302000 59302005 -1.0", "chr19 59302300 59302305 -0.75",
        "#\tcomment", 'track type=wiggle_0 name="variableStep" priority=10',
        "variableStep chrom=chr19 span=4", "59304701 10.0", "59304901 12.5",
        'track type=wiggle_0 name="fixedStep" priority=30',
        "fixedStep chrom=chr19 start=59307401 step=300 span=3", "1000", " 900", " 800", ""])
    got = [",".join(map(str, v)) for v in wiggle.IntervalReader(io.StringIO(text))]
    assert got == ["chr19,59302000,59302005,+,-1.0", "chr19,59302300,59302305,+,-0.75", "chr19,59304700,59304704,+,10.0",
                   "chr19,59304900,59304904,+,12.5", "chr19,59307400,59307403,+,1000.0", "chr19,59307700,59307703,+,900.0",
                   "chr19,59308000,59308003,+,800.0"]
    pos = [",".join(map(str, v)) for v in wiggle.Reader(io.StringIO(text))]
    assert len(pos) == 27 and pos[0] == "chr19,59302000,-1.0" and pos[9] == "chr19,59302304,-0.75"
    assert pos[10] == "chr19,59304700,10.0" and pos[18:21] == ["chr19,59307400,1000.0", "chr19,59307401,1000.0",
                                                              "chr19,59307402,1000.0"] and pos[-1] == "chr19,59308002,800.0"
    spans = wiggle.read_spans(io.StringIO(text))
    assert list(spans) == ["chr19"] and spans["chr19"][0].tolist()[:3] == [59302000, 59302300, 59304700]
