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


if __name__ == "__main__":
    orc.build_ref(REFERENCE)
    bs, ix = orc.ref_modules()
    golden_find(ix)
    golden_neighbors(ix)
    golden_bitset(bs)
    golden_aggregate()
    golden_builders()
