"""
make_golden.py -- generate the committed golden vectors from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

It imports the compiled reference extension modules from oracle/_ref (bx.bitset, bx.intervals.intersection;
built from /root/reference/lib/bx/{bitset.pyx,intervals/intersection.pyx} + src/binBits.c + src/kent/*.c) and,
for the aggregate vectors, the reference's pure-Python callers straight from /root/reference/lib and
/root/reference/scripts/aggregate_scores_in_intervals.py (its main() is executed on small generated files).
Outputs: tests/golden/*.npz / *.json -- inputs are regenerated from the recorded seeds by bx_python_b200/synth.py.
"""
import contextlib
import importlib.util
import io
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from bx_python_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

REFERENCE = "/root/reference"


def ref_find_csr(ix, s, e, qs, qe):
    t = ix.IntervalTree()
    for i, (a, b) in enumerate(zip(s.tolist(), e.tolist())):
        t.insert(a, b, i)
    off = [0]
    hits = []
    for a, b in zip(qs.tolist(), qe.tolist()):
        hits.extend(t.find(a, b))
        off.append(len(hits))
    order = []
    t.traverse(lambda node: order.append(node.interval))
    return np.array(off, np.int64), np.array(hits, np.int32), np.array(order, np.int32)


def golden_find(ix):
    out = {}
    s, e, qs, qe = synth.c1_intervals()
    off, hits, order = ref_find_csr(ix, s, e, qs, qe)
    out["c1_offsets"], out["c1_hits"], out["c1_order"] = off, hits, order
    for k, (s, e, qs, qe) in enumerate(synth.edge_sets()):
        off, hits, order = ref_find_csr(ix, s, e, qs, qe)
        out[f"edge{k}_offsets"], out[f"edge{k}_hits"], out[f"edge{k}_order"] = off, hits, order
    np.savez_compressed(os.path.join(HERE, "find.npz"), **out)
    print("find.npz:", {k: v.shape for k, v in out.items() if k.startswith("c1")})


class _V:
    __slots__ = ("start", "end", "i")

    def __init__(self, s, e, i):
        self.start, self.end, self.i = s, e, i


def golden_neighbors(ix):
    cases = []
    for seed in range(12):
        s, e, queries = synth.neighbor_case(seed)
        t = ix.IntervalTree()
        for i, (a, b) in enumerate(zip(s.tolist(), e.tolist())):
            t.insert(a, b, _V(a, b, i))
        res = []
        for pos, k, md in queries:
            res.append([[v.i for v in t.before(pos, k, md)], [v.i for v in t.after(pos, k, md)]])
        cases.append({"seed": seed, "results": res})
    json.dump(cases, open(os.path.join(HERE, "neighbors.json"), "w"))
    print("neighbors.json:", len(cases), "cases")


def golden_bitset(bs):
    cases = []
    for seed in range(48):
        size, gran, ops, probes = synth.bitset_case(seed)
        b = [bs.BinnedBitSet(size, gran), bs.BinnedBitSet(size, gran)]
        for op in ops:
            synth.apply_bitset_op(b, op)
        res = []
        for k in (0, 1):
            r = {"bin_size": b[k].bin_size, "count": [], "next_set": [], "next_clear": [], "get": []}
            for s, c in probes:
                r["count"].append(b[k].count_range(s, c))
                r["next_set"].append(b[k].next_set(s))
                r["next_clear"].append(b[k].next_clear(s))
                r["get"].append(b[k][s])
            res.append(r)
        cases.append({"seed": seed, "results": res})
    json.dump(cases, open(os.path.join(HERE, "bitset.json"), "w"))
    print("bitset.json:", len(cases), "cases")

    # C3-shaped (scaled 1/100): two 2.5 Mbp bitmaps, dense and sparse fills, and + count + runs + invert bug probe
    out = {}
    for tag, nr in (("dense", 4000), ("sparse", 200)):
        size = 2_500_000
        a, b2 = bs.BinnedBitSet(size), bs.BinnedBitSet(size)
        (sa, ca), (sb, cb), (qs, qc) = synth.c3_case(size, nr, 31)
        for s, c in zip(sa.tolist(), ca.tolist()):
            a.set_range(s, c)
        for s, c in zip(sb.tolist(), cb.tolist()):
            b2.set_range(s, c)
        out[f"{tag}_count_a"] = np.int64(a.count_range(0, size))
        a.iand(b2)
        out[f"{tag}_count_and"] = np.int64(a.count_range(0, size))
        out[f"{tag}_counts"] = np.array([a.count_range(s, c) for s, c in zip(qs.tolist(), qc.tolist())], np.int32)
        runs = []
        end = 0
        while end < size:
            st = a.next_set(end)
            if st == size:
                break
            end = a.next_clear(st)
            runs.append((st, end))
        out[f"{tag}_runs"] = np.array(runs, np.int32).reshape(-1, 2)
        a.invert()
        out[f"{tag}_inv_counts"] = np.array([a.count_range(s, c) for s, c in zip(qs.tolist(), qc.tolist())], np.int32)
        out[f"{tag}_inv_total"] = np.int64(a.count_range(0, size))
    np.savez_compressed(os.path.join(HERE, "bitset_c3.npz"), **out)
    print("bitset_c3.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


def golden_aggregate():
    """Execute the reference script's main() on generated wiggle/BED/mask files and record its printed lines."""
    import bx
    import bx.intervals
    bx.__path__.append(os.path.join(REFERENCE, "lib", "bx"))
    bx.intervals.__path__.append(os.path.join(REFERENCE, "lib", "bx", "intervals"))
    sys.path.append(os.path.join(REFERENCE, "lib"))
    spec = importlib.util.spec_from_file_location(
        "ref_aggregate_script", os.path.join(REFERENCE, "scripts", "aggregate_scores_in_intervals.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases = []
    for seed in range(6):
        origin, scores, ws, we, mask_runs = synth.aggregate_case(seed)
        with tempfile.TemporaryDirectory() as d:
            wig, bed, out, msk = (os.path.join(d, n) for n in ("s.wig", "w.bed", "o.txt", "m.bed"))
            with open(wig, "w") as f:
                f.write(f"fixedStep chrom=chr1 start={origin + 1} step=1\n")
                for v in scores:
                    f.write("nan\n" if v != v else repr(float(v)) + "\n")
            with open(bed, "w") as f:
                for a, b in zip(ws.tolist(), we.tolist()):
                    f.write(f"chr1\t{a}\t{b}\n")
            argv = [mod.__file__, wig, bed, out]
            if mask_runs is not None:
                with open(msk, "w") as f:
                    for a, b in mask_runs:
                        f.write(f"chr1\t{a}\t{b}\n")
                argv = [mod.__file__, "--mask", msk, wig, bed, out]
            old = sys.argv
            sys.argv = argv
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    mod.main()
            finally:
                sys.argv = old
            lines = [ln.rstrip("\n").split("\t")[3:] for ln in open(out)]
        cases.append({"seed": seed, "lines": lines})
    json.dump(cases, open(os.path.join(HERE, "aggregate.json"), "w"))
    print("aggregate.json:", len(cases), "cases;", cases[0]["lines"][:3])


def golden_builders():
    """Reference lib/bx/bitset_builders.py + lib/bx/bitset_utils.py (pure Python over the compiled bx.bitset)."""
    import bx
    if os.path.join(REFERENCE, "lib", "bx") not in bx.__path__:
        bx.__path__.append(os.path.join(REFERENCE, "lib", "bx"))
    from bx import bitset_builders as bb
    from bx import bitset_utils as bu
    cases = []
    for seed in range(4):
        lines, lens = synth.bed_lines(seed)
        out = {}
        for name, fn, kw in (("file", bb.binned_bitsets_from_file, {"lens": lens}),
                             ("file_pad", bb.binned_bitsets_from_file, {"lens": lens, "upstream_pad": 25}),
                             ("bed", bb.binned_bitsets_from_bed_file, {"lens": lens}),
                             ("prox", bb.binned_bitsets_proximity, {"upstream": 30, "downstream": 10})):
            d = fn(lines if name != "prox" else [ln for ln in lines if not ln.isspace()], **kw)   # prox does not skip blanks
            out[name] = {c: bu.bits2list(b) for c, b in d.items()}
        chr1 = [ln for ln in lines if ln.startswith("chr1\t")]
        out["file_pad1"] = {c: bu.bits2list(b) for c, b in
                            bb.binned_bitsets_from_file(chr1, lens=lens, upstream_pad=10, downstream_pad=40).items()}
        lst = [ln.split()[:3] for ln in lines if not ln.startswith("#") and not ln.isspace()]
        out["list"] = {c: bu.bits2list(b) for c, b in bb.binned_bitsets_from_list(lst).items()}
        noblank = [ln for ln in lines if not ln.isspace()]            # by_chrom does not skip blank lines either
        out["by_chrom"] = bu.bits2list(bb.binned_bitsets_by_chrom(noblank, "chr2"))
        ex1, ex2 = synth.exon_lists(seed)
        out["intersect"] = bu.bitset_intersect(ex1, ex2)
        out["subtract"] = bu.bitset_subtract(ex1, ex2)
        out["complement"] = bu.bitset_complement(ex1)
        out["union"] = bu.bitset_union(ex1 + ex2)
        bits = bu.list2bits(ex1)
        out["interval_intersect"] = [bu.bitset_interval_intersect(bits, a, b) for a, b in ((0, 5000), (100, 900), (2500, 2600))]
        cases.append({"seed": seed, "out": out})
    json.dump(cases, open(os.path.join(HERE, "builders.json"), "w"))
    print("builders.json:", len(cases), "cases")


def _ref_pure_python():
    """Make the reference's pure-Python modules importable next to the compiled ones in oracle/_ref."""
    import bx
    import bx.bbi
    import bx.intervals
    for pkg, sub in ((bx, ""), (bx.intervals, "intervals"), (bx.bbi, "bbi")):
        d = os.path.join(REFERENCE, "lib", "bx", sub)
        if d not in pkg.__path__:
            pkg.__path__.append(d)
    if os.path.join(REFERENCE, "lib") not in sys.path:
        sys.path.append(os.path.join(REFERENCE, "lib"))           # bx_extras


def golden_scores():
    """Reference wiggle.Reader -> BinnedArray per-base loop (load_scores_wiggle), BinnedArray.to_file bytes, and the
    aggregate script run on the same wiggle text."""
    _ref_pure_python()
    import hashlib
    import bx.wiggle
    from bx.binned_array import BinnedArray
    out = {}
    for seed in range(8):
        text = synth.wiggle_text(seed)
        recs = [(c, s, e, st, float(v)) for c, s, e, st, v in bx.wiggle.IntervalReader(io.StringIO(text))]
        arrays = {}
        for chrom, pos, val in bx.wiggle.Reader(io.StringIO(text)):
            if chrom not in arrays:
                arrays[chrom] = BinnedArray(bin_size=1024, max_size=8192)
            arrays[chrom][pos] = val
        for chrom, ba in arrays.items():
            out[f"s{seed}_{chrom}_dense"] = ba.get_range(0, 8192)
            buf = io.BytesIO()
            ba.to_file(buf)
            out[f"s{seed}_{chrom}_file_sha256"] = np.frombuffer(hashlib.sha256(buf.getvalue()).digest(), np.uint8)
        out[f"s{seed}_chroms"] = np.array(list(arrays.keys()))
        out[f"s{seed}_nrecords"] = np.int64(len(recs))
        out[f"s{seed}_rec_checksum"] = np.int64(sum((s * 31 + e * 17) % 1000003 for _, s, e, _, _ in recs))
    np.savez_compressed(os.path.join(HERE, "scores.npz"), **out)
    print("scores.npz:", len(out), "arrays")


def golden_summarize():
    """SummarizedData.accumulate_interval_value of the compiled reference (through oracle/bbi_shim) on synthetic
    batches, and BigWigFile.get / summarize_from_full on the reference's own test.bw."""
    _ref_pure_python()
    import bbi_shim
    from bx.bbi.bbi_file import SummarizedData
    from bx.bbi.bigwig_file import BigWigFile
    out = {}
    keys = ("valid_count", "min_val", "max_val", "sum_data", "sum_squares")
    for seed in range(24):
        s, e, v, rs, re_, size = synth.summarize_case(seed)
        sd = SummarizedData(rs, re_, size)
        if seed % 4 < 2:
            sd.min_val[:] = np.inf
            sd.max_val[:] = -np.inf
        bbi_shim.accumulate(sd, s.tolist(), e.tolist(), v.tolist())
        out[f"c{seed}"] = np.stack([getattr(sd, k) for k in keys])
    bw = BigWigFile(file=open(os.path.join(REFERENCE, "test_data", "bbi_tests", "test.bw"), "rb"))
    iv = bw.get(b"chr1", 10000, 20000)
    out["bw_start"] = np.array([x[0] for x in iv], np.int32)
    out["bw_end"] = np.array([x[1] for x in iv], np.int32)
    out["bw_val"] = np.array([x[2] for x in iv], np.float32)
    regions = [(10000, 20000, 10), (10000, 20000, 7), (10917, 11500, 33), (12345, 19999, 100), (10000, 10920, 4)]
    out["bw_regions"] = np.array(regions, np.int64)
    for k, (a, b, size) in enumerate(regions):
        sd = bw.summarize_from_full(b"chr1", a, b, size)
        out[f"bw{k}"] = np.stack([getattr(sd, key) for key in keys])      # valid_count is rounded by the caller (:182-184)
    np.savez_compressed(os.path.join(HERE, "summarize.npz"), **out)
    print("summarize.npz:", len(out), "arrays;", len(iv), "bigWig intervals")


def canonical_join_rows(rows, leftlen):
    """Rows of one left interval are contiguous but in treap pre-order (random): sort inside each such group."""
    out, group, key = [], [], None
    for r in rows:
        if not isinstance(r, list):
            r = ["<passthrough>"]
        k = tuple(r[:leftlen])
        if k != key or all(x == "." for x in k):
            out.extend(sorted(group))
            group, key = [], k
        if all(x == "." for x in k):
            out.append(r)                              # left-fill tail: order is deterministic, keep it
        else:
            group.append(r)
    out.extend(sorted(group))
    return out


def golden_join():
    _ref_pure_python()
    from bx.intervals.io import NiceReaderWrapper
    from bx.intervals.operations.join import join
    cases = []
    for seed in range(6):
        left, right, mincols = synth.join_case(seed)

        def rd(lines):
            return NiceReaderWrapper(io.StringIO("\n".join(lines) + "\n"), chrom_col=0, start_col=1, end_col=2,
                                     fix_strand=True)
        for lf, rf in ((True, True), (False, True), (True, False)):
            rows = list(join(rd(left), rd(right), mincols=mincols, leftfill=lf, rightfill=rf))
            cases.append({"seed": seed, "mincols": mincols, "leftfill": lf, "rightfill": rf,
                          "rows": canonical_join_rows(rows, 4)})
    json.dump(cases, open(os.path.join(HERE, "join.json"), "w"))
    print("join.json:", len(cases), "cases;", sum(len(c["rows"]) for c in cases), "rows")


def golden_operations():
    """The reference's own interval operations (lib/bx/intervals/operations/*.py over NiceReaderWrapper readers) on
    synth.ops_case inputs -- what bx_python_b200.intervals.operations.arrays must reproduce row for row."""
    _ref_pure_python()
    from bx.intervals.io import GenomicInterval, NiceReaderWrapper
    from bx.intervals.operations.base_coverage import base_coverage
    from bx.intervals.operations.complement import complement
    from bx.intervals.operations.coverage import coverage
    from bx.intervals.operations.intersect import intersect
    from bx.intervals.operations.merge import merge
    from bx.intervals.operations.subtract import subtract

    def rd(lines):
        return NiceReaderWrapper(iter([ln + "\n" for ln in lines]), chrom_col=0, start_col=1, end_col=2, strand_col=5,
                                 fix_strand=True)

    def rows(gen):
        out = []
        for r in gen:
            if isinstance(r, GenomicInterval):
                out.append([str(f).rstrip("\n") for f in r.fields])
            elif isinstance(r, list):
                out.append([str(f) for f in r])
        return out
    cases = []
    for seed in range(4):
        p, s2, s3, lens = synth.ops_case(seed)
        c = {"seed": seed}
        for pieces in (True, False):
            for mincols in (1, 40):
                key = f"pieces{int(pieces)}_min{mincols}"
                c["intersect_" + key] = rows(intersect([rd(p), rd(s2)], mincols=mincols, pieces=pieces, lens=lens))
                c["subtract_" + key] = rows(subtract([rd(p), rd(s2)], mincols=mincols, pieces=pieces, lens=lens))
        c["intersect3"] = rows(intersect([rd(p), rd(s2), rd(s3)], lens=lens))
        c["subtract3"] = rows(subtract([rd(p), rd(s2), rd(s3)], lens=lens))
        c["merge"] = rows(merge(rd(p)))
        c["complement"] = rows(complement(rd(s2), lens))
        c["coverage"] = rows(coverage([rd(p), rd(s2)]))
        c["coverage3"] = rows(coverage([rd(p), rd(s2), rd(s3)]))
        c["base_coverage"] = int(base_coverage(rd(p)))
        cases.append(c)
    json.dump(cases, open(os.path.join(HERE, "operations.json"), "w"))
    print("operations.json:", {k: (len(v) if isinstance(v, list) else v) for k, v in cases[0].items()})


def golden_scripts():
    """stdout of the reference's scripts (staged unmodified in oracle/_ref/scripts) run on the compiled reference over
    the inputs of tests/dropin.py:make_inputs -- what the drop-in tests compare the shadowed runs with."""
    import subprocess
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dropin
    out = {"inputs_seed": 0, "runs": []}
    with tempfile.TemporaryDirectory() as d:
        dropin.make_inputs(d, 0)
        for name, argv, _ in dropin.SCRIPT_RUNS:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin.py"), "--impl", "reference", "script",
                                name, d] + argv, capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, (name, argv, r.stderr[-2000:])
            out["runs"].append({"script": name, "argv": argv, "stdout": r.stdout.splitlines()})
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin.py"), "--impl", "reference", "unittests"],
                           capture_output=True, text=True, timeout=600)
        res = json.loads(r.stdout.strip().splitlines()[-1])
        assert res["rc"] == 0 and res["failed"] == 0
        out["unittests_passed"] = res["passed"]
    json.dump(out, open(os.path.join(HERE, "scripts.json"), "w"), indent=0)
    print("scripts.json:", [(r["script"], len(r["stdout"])) for r in out["runs"]], "unit tests:", out["unittests_passed"])


if __name__ == "__main__":
    orc.build_ref(REFERENCE)
    if sys.argv[1:] == ["scripts"]:
        golden_scripts()
        sys.exit(0)
    if sys.argv[1:] == ["operations"]:
        orc.ref_modules()
        golden_operations()
        sys.exit(0)
    bs, ix = orc.ref_modules()
    golden_find(ix)
    golden_neighbors(ix)
    golden_bitset(bs)
    golden_aggregate()
    golden_builders()
    golden_scores()
    golden_summarize()
    golden_join()
    golden_scripts()
    golden_operations()
