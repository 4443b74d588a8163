"""
Pins the CPU restatement (oracle/bx_oracle.c) against the UNMODIFIED reference compiled into oracle/_ref
(bx.intervals.intersection / bx.bitset built from /root/reference by oracle/Makefile).  Skipped where
oracle/_ref was not built; tests/test_oracle_golden.py covers the same ground from committed fixtures.
"""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.skipif(not orc.ref_available(), reason="oracle/_ref not built")


def _ref():
    return orc.ref_modules()


def _rand_tree(rng, n, G, weird):
    if weird:
        s = rng.integers(-3, G, n)
        e = s + rng.integers(-2, 6, n)
    else:
        s = rng.integers(0, G, n)
        e = s + rng.integers(1, max(2, G // 4), n)
    return s.astype(np.int32), e.astype(np.int32)


@pytest.mark.parametrize("seed", range(40))
def test_find_order_and_hits(seed):
    _, ix = _ref()
    rng = np.random.default_rng(seed)
    n = int(rng.integers(0, 300))
    G = int(rng.integers(5, 60)) if seed % 2 else int(rng.integers(100, 100000))
    s, e = _rand_tree(rng, n, G, weird=bool(seed % 2))
    t = ix.IntervalTree()
    for i in range(n):
        t.insert(int(s[i]), int(e[i]), i)
    o = orc.OracleIntervalTree(s, e)
    order = []
    t.traverse(lambda node: order.append(node.interval))
    assert order == o.order().tolist()
    nq = 60
    qs = rng.integers(-5, G + 5, nq).astype(np.int32)
    qe = (qs + rng.integers(-3, max(12, G // 3), nq)).astype(np.int32)
    off, hits = o.find(qs, qe)
    for q in range(nq):
        assert t.find(int(qs[q]), int(qe[q])) == hits[off[q]:off[q + 1]].tolist()


class _V:
    """value object with .start/.end so the reference's attrgetter post-sort works"""
    __slots__ = ("start", "end", "i")

    def __init__(self, s, e, i):
        self.start, self.end, self.i = s, e, i


@pytest.mark.parametrize("seed", range(20))
def test_before_after(seed):
    _, ix = _ref()
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 200))
    G = int(rng.integers(20, 3000))
    s = rng.integers(0, G, n).astype(np.int32)
    e = (s + rng.integers(0, 30, n)).astype(np.int32)
    t = ix.IntervalTree()
    for i in range(n):
        t.insert(int(s[i]), int(e[i]), _V(int(s[i]), int(e[i]), i))
    o = orc.OracleIntervalTree(s, e)
    for _ in range(60):
        pos = int(rng.integers(-5, G + 40))
        k = int(rng.integers(1, 6))
        md = int(rng.choice([0, 1, 5, 50, 2500]))
        assert [v.i for v in t.before(pos, k, md)] == o.before(pos, k, md).tolist()
        assert [v.i for v in t.after(pos, k, md)] == o.after(pos, k, md).tolist()


def _random_ops(rng, size, gran, nops, ref_cls):
    rb = [ref_cls(size, gran) for _ in range(2)]
    ob = [orc.OracleBinnedBitSet(size, gran) for _ in range(2)]
    for _ in range(nops):
        k = int(rng.integers(0, 2))
        op = int(rng.integers(0, 7))
        if op == 0:
            s = int(rng.integers(0, size)); c = int(rng.integers(0, min(size - s, max(1, size // 3)) + 1))
            rb[k].set_range(s, c); ob[k].set_range(s, c)
        elif op == 1:
            p = int(rng.integers(0, size)); rb[k].set(p); ob[k].set(p)
        elif op == 2:
            p = int(rng.integers(0, size)); rb[k].clear(p); ob[k].clear(p)
        elif op == 3:
            rb[k].invert(); ob[k].invert()
        elif op == 4:
            rb[k].iand(rb[1 - k]); ob[k].iand(ob[1 - k])
        elif op == 5:
            rb[k].ior(rb[1 - k]); ob[k].ior(ob[1 - k])
    return rb, ob


@pytest.mark.parametrize("seed", range(60))
def test_binned_bitset_state_machine(seed):
    bs, _ = _ref()
    rng = np.random.default_rng(2000 + seed)
    size = int(rng.integers(1, 3000))
    gran = int(rng.choice([1, 2, 3, 7, 10, 64, 1024]))
    rb, ob = _random_ops(rng, size, gran, int(rng.integers(1, 25)), bs.BinnedBitSet)
    for r, o in zip(rb, ob):
        assert (r.size, r.bin_size) == (o.size, o.bin_size)
        for _ in range(80):
            s = int(rng.integers(0, size)); c = int(rng.integers(0, size - s + 1))
            assert r.count_range(s, c) == o.count_range(s, c)
            assert r.next_set(s) == o.next_set(s)
            assert r.next_clear(s) == o.next_clear(s)
            assert r[s] == o[s]


def test_geometry():
    bs, _ = _ref()
    for size, gran in [(2**29, 1024), (250000000, 1024), (248956422, 1024), (100, 1), (100, 3), (100, 1024),
                       (1000, 10), (1000, 20), (2**31 - 1, 1024), (57227415, 1024), (7, 3)]:
        assert bs.BinnedBitSet(size, gran).bin_size == orc.OracleBinnedBitSet(size, gran).bin_size


@pytest.mark.parametrize("seed", range(10))
def test_flat_bitset(seed):
    bs, _ = _ref()
    rng = np.random.default_rng(3000 + seed)
    n = int(rng.integers(1, 2000))
    r = [bs.BitSet(n) for _ in range(2)]
    o = [orc.OracleBitSet(n) for _ in range(2)]
    for _ in range(30):
        k = int(rng.integers(0, 2)); op = int(rng.integers(0, 7))
        if op == 0:
            s = int(rng.integers(0, n)); c = int(rng.integers(0, n - s + 1)); r[k].set_range(s, c); o[k].set_range(s, c)
        elif op == 1:
            p = int(rng.integers(0, n)); r[k].set(p); o[k].set(p)
        elif op == 2:
            p = int(rng.integers(0, n)); r[k].clear(p); o[k].clear(p)
        elif op == 3:
            r[k].invert(); o[k].invert()
        elif op == 4:
            r[k].iand(r[1 - k]); o[k].iand(o[1 - k])
        elif op == 5:
            r[k].ior(r[1 - k]); o[k].ior(o[1 - k])
        else:
            r[k].ixor(r[1 - k]); o[k].ixor(o[1 - k])
    for a, b in zip(r, o):
        for _ in range(100):
            s = int(rng.integers(0, n)); e = int(rng.integers(s, n + 1))
            assert a.count_range(s, e - s) == b.count_range(s, e - s)
            assert a.next_set(s, e) == b.next_set(s, e)
            assert a.next_clear(s, e) == b.next_clear(s, e)
            assert a[s] == b[s]


def test_runs_idiom():
    bs, _ = _ref()
    rng = np.random.default_rng(7)
    for size, gran in [(1000, 10), (95, 10), (4096, 1024), (777, 3)]:
        r = bs.BinnedBitSet(size, gran); o = orc.OracleBinnedBitSet(size, gran)
        for _ in range(20):
            s = int(rng.integers(0, size)); c = int(rng.integers(0, min(40, size - s) + 1))
            r.set_range(s, c); o.set_range(s, c)
        if size == 95:
            r.set_range(90, 5); o.set_range(90, 5)
        runs = []
        end = 0
        while end < size:
            st = r.next_set(end)
            if st == size:
                break
            end = r.next_clear(st)
            runs.append((st, end))
        rs, re = o.runs()
        assert runs == list(zip(rs.tolist(), re.tolist()))


# ---------------------------------------------------------------------------------------------------------------
# score sources, bigWig summary, join -- against the live reference (needs /root/reference for its
# pure-Python modules, so these also skip on a box that only carries the compiled oracle/_ref)
# ---------------------------------------------------------------------------------------------------------------
import io  # noqa: E402
import os  # noqa: E402
import sys  # noqa: E402

REFERENCE = "/root/reference"
needs_reference = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="/root/reference not present")


def _ref_pure_python():
    orc.ref_modules()
    import bx
    import bx.intervals
    import bx.bbi
    for pkg, sub in ((bx, ""), (bx.intervals, "intervals"), (bx.bbi, "bbi")):
        d = os.path.join(REFERENCE, "lib", "bx", sub)
        if d not in pkg.__path__:
            pkg.__path__.append(d)
    if os.path.join(REFERENCE, "lib") not in sys.path:
        sys.path.append(os.path.join(REFERENCE, "lib"))


@needs_reference
@pytest.mark.parametrize("seed", range(100, 140))
def test_summarize_vs_compiled_accumulate(seed):
    from bx_python_b200 import synth
    _ref_pure_python()                                 # bx.bbi.bbi_file imports the pure-Python bx.misc.binary_file
    try:
        import bbi_shim
        from bx.bbi.bbi_file import SummarizedData
    except ImportError:
        pytest.skip("bbi modules not built in oracle/_ref")
    s, e, v, rs, re_, size = synth.summarize_case(seed)
    sd = SummarizedData(rs, re_, size)
    bbi_shim.accumulate(sd, s.tolist(), e.tolist(), v.tolist())
    o = orc.summarize(s, e, v, rs, re_, size, 0.0, 0.0)
    for k in o:
        assert np.array_equal(o[k].view(np.uint64), getattr(sd, k).view(np.uint64)), k


@needs_reference
@pytest.mark.parametrize("seed", range(100, 112))
def test_span_loop_vs_reference_binned_array(seed):
    from bx_python_b200 import synth, wiggle
    _ref_pure_python()
    import bx.wiggle
    from bx.binned_array import BinnedArray
    text = synth.wiggle_text(seed)
    arrays = {}
    for chrom, pos, val in bx.wiggle.Reader(io.StringIO(text)):
        arrays.setdefault(chrom, BinnedArray(bin_size=1024, max_size=8192))[pos] = val
    assert repr(list(wiggle.IntervalReader(io.StringIO(text)))) == repr(list(bx.wiggle.IntervalReader(io.StringIO(text))))
    spans = wiggle.read_spans(io.StringIO(text))
    for chrom, ba in arrays.items():
        t = np.full(8192, np.nan, np.float32)
        orc.scores_set_spans(t, 0, *spans[chrom])
        assert np.array_equal(t.view(np.uint32), ba.get_range(0, 8192).view(np.uint32))


@needs_reference
@pytest.mark.parametrize("seed", range(100, 106))
def test_join_vs_reference(seed):
    from bx_python_b200 import synth
    _ref_pure_python()
    from bx.intervals.io import NiceReaderWrapper
    from bx.intervals.operations.join import join
    left, right, mincols = synth.join_case(seed, nl=60, nr=70)

    def rd(lines):
        return NiceReaderWrapper(io.StringIO("\n".join(lines) + "\n"), chrom_col=0, start_col=1, end_col=2, fix_strand=True)
    rows = [r for r in join(rd(left), rd(right), mincols=mincols, leftfill=False, rightfill=False) if isinstance(r, list)]
    lf = [ln.split("\t") for ln in left if not ln.startswith("#")]
    rf = [ln.split("\t") for ln in right]
    cid = {}
    it = [cid.setdefault(r[0], len(cid)) for r in rf]
    qt = [cid.get(r[0], -1) for r in lf]
    off, items, _ = orc.join(it, [int(r[1]) for r in rf], [int(r[2]) for r in rf], qt, [int(r[1]) for r in lf],
                             [int(r[2]) for r in lf], mincols)
    mine = sorted(lf[q] + rf[i] for q in range(len(lf)) for i in items[off[q]:off[q + 1]].tolist())
    assert mine == sorted(rows)
