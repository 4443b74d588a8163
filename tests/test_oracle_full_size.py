"""
The CPU restatement (oracle/bx_oracle.c) against the COMPILED REFERENCE at BASELINE.json's full sizes: the digests in
tests/golden/full_size.json (tests/golden/make_golden_full.py) are recomputed from the restatement's answers on the
regenerated inputs.  This closes the chain  reference == restatement == CUDA path  at full size; a sample of
chromosomes keeps the CPU suite short (the GPU tests cover all 24).
"""
import hashlib
import json
import os

import numpy as np
import pytest

from bx_python_b200 import synth
from oracle import oracle as orc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CHROMS = (0, 2, 20)          # chr1 (largest), chr3 (the reference's float32 bin geometry falls 7 bits short), chr21


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(G, "full_size.json")))


def test_golden_file_shape(gold):
    assert sorted(gold) == ["c2", "c3", "c4", "c5"] and all(len(gold[k]) == 24 for k in gold)
    assert sum(g["intervals"] for g in gold["c2"]) == 10_000_000 and sum(g["queries"] for g in gold["c2"]) == 10_000_000
    assert sum(g["lines"] for g in gold["c4"]) == 50_000_000 and sum(g["ranges"] for g in gold["c4"]) == 50_000_000
    assert sum(g["scores"] for g in gold["c5"]) == 100_000_000 and sum(g["windows"] for g in gold["c5"]) == 5_000_000


def test_c2_restatement_vs_compiled_reference(gold):
    db = synth.genome_intervals(10_000_000, 2001)
    qq = synth.genome_intervals(10_000_000, 2002)
    for c in CHROMS:
        off, hits = orc.OracleIntervalTree(*db[c]).find(*qq[c])
        assert int(off[-1]) == gold["c2"][c]["hits"]
        assert sha(off, hits) == gold["c2"][c]["sha256"], c


@pytest.mark.parametrize("tag,nranges,seed0", [("dense", 400_000, 0), ("sparse", 20_000, 100)])
def test_c3_restatement_vs_compiled_reference(gold, tag, nranges, seed0):
    size = 250_000_000
    c = 5
    g = gold["c3"][c][tag]
    (sa, ca), (sb, cb), (ps, pc) = synth.c3_case(size, nranges, seed0 + c, nq=41_667)
    a, b = orc.OracleBinnedBitSet(size), orc.OracleBinnedBitSet(size)
    a.set_ranges(sa, ca)
    b.set_ranges(sb, cb)
    assert (a.count_range(0, size), b.count_range(0, size)) == (g["count_a"], g["count_b"])
    a.iand(b)
    assert a.count_range(0, size) == g["count_and"]
    assert sha(a.count_ranges(ps, pc)) == g["counts_sha256"]
    rs, re = a.runs()
    assert len(rs) == g["nruns"] and sha(np.stack([rs, re], axis=1).astype(np.int32)) == g["runs_sha256"]
    a.invert()
    assert a.count_range(0, size) == g["inv_total"]
    assert sha(a.count_ranges(ps, pc)) == g["inv_counts_sha256"]


def test_c4_restatement_vs_compiled_reference(gold):
    f2 = synth.genome_intervals(50_000_000, 4002)
    f1 = synth.genome_intervals(50_000_000, 4001)
    for c in CHROMS:
        g = gold["c4"][c]
        size = int(synth.HG38_LENS[c])
        b = orc.OracleBinnedBitSet(size)
        b.set_ranges(f2[c][0], f2[c][1] - f2[c][0])
        counts = b.count_ranges(f1[c][0], f1[c][1] - f1[c][0])
        assert sha(counts) == g["sha256"], c
        assert (int((counts >= 1).sum()), int(counts.astype(np.int64).sum())) == (g["overlapping"], g["sum_counts"])
        assert b.count_range(0, g["addressable"]) == g["covered"]
    # hg38 chr3: binBitsAlloc's float32 geometry gives 1024 bins x 193 648 bits = 7 positions short of the chromosome
    assert gold["c4"][2]["addressable"] == int(synth.HG38_LENS[2]) - 7
    assert all(gold["c4"][c]["addressable"] == int(synth.HG38_LENS[c]) for c in range(24) if c != 2)


def test_c5_restatement_vs_reference_script(gold):
    tracks = synth.genome_scores(100_000_000, 5_000_000, 5001)
    for c in CHROMS:
        origin, v, ws, we = tracks[c]
        dense = np.full(origin + len(v), np.nan, np.float32)
        dense[origin:] = v
        r = orc.aggregate(dense, ws, we)
        g = gold["c5"][c]
        assert int((r["count"] == 0).sum()) == g["nan_lines"]
        assert sha(r["avg"], r["min"], r["max"]) == g["sha256"], c
