"""
make_golden_full.py -- FULL-SIZE golden answers of the UNMODIFIED reference for BASELINE configs[1..4].

Run in the build container only (needs /root/reference and `make -C oracle ref`; ~10 CPU-minutes on 8 cores):

    python tests/golden/make_golden_full.py [c2] [c3] [c4] [c5]

The inputs are far too large to commit, and so are the outputs, so what is committed (tests/golden/full_size.json) is, per
chromosome, a SHA-256 of the reference's answer plus a few integers; the GPU tests regenerate the inputs from the seeds
(bx_python_b200/synth.py), run the CUDA path at full size and compare digests.  Definitions (little-endian, C order):

  c2[c]  IntervalTree of genome_intervals(10 M, 2001)[c] (value = insertion index), find() for every query of
         genome_intervals(10 M, 2002)[c] in generated order:
             sha256( offsets int64[nq+1] || hits int32[total] ), "hits": total
  c3[c]  a, b = BinnedBitSet(250 000 000) filled from synth.c3_case(size, R, seed, nq=41667) with R = 400 000, seed = c
         ("dense") or R = 20 000, seed = 100 + c ("sparse"):
             count_a, count_b; a.iand(b); count_and; sha256(int32 count_range(probes));
             runs of a by the next_set / next_clear idiom: nruns, sha256(int32[nruns][2]);
             a.invert(): inv_total = count_range(0, size), sha256(int32 count_range(probes))   (strict ALL_ONE arithmetic)
  c4[c]  bits = BinnedBitSet(hg38 length) with set_range(s, e - s) for genome_intervals(50 M, 4002)[c];
         counts = count_range(s, e - s) for genome_intervals(50 M, 4001)[c]   (scripts/bed_intersect.py:42-53):
             lines, overlapping = #(counts >= 1), sum_counts, covered = count_range(0, addressable), sha256(int32 counts)
         (addressable = min(size, nbins * bin_size): for hg38 chr3 the reference's float32 bin geometry leaves the last 7
         positions outside its bins array -- undefined behaviour there; no generated interval touches them)
  c5[c]  scripts/aggregate_scores_in_intervals.py main() on a fixedStep wiggle + BED written from
         synth.genome_scores(100 M, 5 M, 5001)[c]; its printed avg / min / max columns parsed back to float32:
             windows, nan_lines, sha256( avg f32[] || min f32[] || max f32[] )
"""
import contextlib
import hashlib
import importlib.util
import io
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from bx_python_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

REFERENCE = "/root/reference"
OUT = os.path.join(HERE, "full_size.json")
C3_SIZE = 250_000_000
C3_PROBES = 41_667

_G = {}          # inputs inherited by the forked workers


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


# ---- C2 ---------------------------------------------------------------------------------------------------------------
def _c2_worker(c):
    _, ix = orc.ref_modules()
    (s, e), (qs, qe) = _G["db"][c], _G["qq"][c]
    t = ix.IntervalTree()
    ins = t.insert
    for i, (a, b) in enumerate(zip(s.tolist(), e.tolist())):
        ins(a, b, i)
    off = np.empty(len(qs) + 1, np.int64)
    off[0] = 0
    hits = []
    ext = hits.extend
    f = t.find
    for k, (a, b) in enumerate(zip(qs.tolist(), qe.tolist())):
        ext(f(a, b))
        off[k + 1] = len(hits)
    hits = np.array(hits, np.int32)
    return c, {"intervals": len(s), "queries": len(qs), "hits": int(off[-1]), "sha256": sha(off, hits)}


def golden_c2(pool_size):
    _G["db"] = synth.genome_intervals(10_000_000, 2001)
    _G["qq"] = synth.genome_intervals(10_000_000, 2002)
    order = sorted(range(24), key=lambda c: -len(_G["db"][c][0]))
    with mp.get_context("fork").Pool(pool_size) as pool:
        res = dict(pool.imap_unordered(_c2_worker, order))
    return [res[c] for c in range(24)]


# ---- C3 ---------------------------------------------------------------------------------------------------------------
def c3_one(bs, nranges, seed):
    size = C3_SIZE
    (sa, ca), (sb, cb), (ps, pc) = synth.c3_case(size, nranges, seed, nq=C3_PROBES)
    a, b = bs.BinnedBitSet(size), bs.BinnedBitSet(size)
    for s, c in zip(sa.tolist(), ca.tolist()):
        a.set_range(s, c)
    for s, c in zip(sb.tolist(), cb.tolist()):
        b.set_range(s, c)
    out = {"ranges": nranges, "seed": seed, "bin_size": a.bin_size,
           "count_a": a.count_range(0, size), "count_b": b.count_range(0, size)}
    a.iand(b)
    out["count_and"] = a.count_range(0, size)
    probes = list(zip(ps.tolist(), pc.tolist()))
    out["counts_sha256"] = sha(np.array([a.count_range(s, c) for s, c in probes], np.int32))
    runs = []
    end = 0
    while end < size:
        st = a.next_set(end)
        if st == size:
            break
        end = a.next_clear(st)
        runs.append((st, end))
    out["nruns"] = len(runs)
    out["runs_sha256"] = sha(np.array(runs, np.int32).reshape(-1, 2))
    a.invert()
    out["inv_total"] = a.count_range(0, size)
    out["inv_counts_sha256"] = sha(np.array([a.count_range(s, c) for s, c in probes], np.int32))
    return out


def _c3_worker(c):
    bs, _ = orc.ref_modules()
    return c, {"dense": c3_one(bs, 400_000, c), "sparse": c3_one(bs, 20_000, 100 + c)}


def golden_c3(pool_size):
    with mp.get_context("fork").Pool(pool_size) as pool:
        res = dict(pool.imap_unordered(_c3_worker, range(24)))
    return [res[c] for c in range(24)]


# ---- C4 ---------------------------------------------------------------------------------------------------------------
def _c4_worker(c):
    bs, _ = orc.ref_modules()
    size = int(synth.HG38_LENS[c])
    (s2, e2), (s1, e1) = _G["f2"][c], _G["f1"][c]
    bits = bs.BinnedBitSet(size)
    sr = bits.set_range
    for a, n in zip(s2.tolist(), (e2 - s2).tolist()):
        sr(a, n)
    cr = bits.count_range
    counts = np.array([cr(a, n) for a, n in zip(s1.tolist(), (e1 - s1).tolist())], np.int32)
    # binBitsAlloc's float32 geometry (binBits.c:8-17) can leave nbins * bin_size < size: hg38 chr3 (198 295 559) gets
    # 1024 bins of 193 648 bits = 198 295 552, so its last 7 positions index bins[1024] -- out of bounds, undefined
    # behaviour in the reference (this script segfaulted on count_range(0, size) there).  No generated interval reaches
    # those positions, so "covered" is counted over the part the reference can address.
    nbins = int(np.ceil(np.float32(size) / np.float32(bits.bin_size)))
    addressable = min(size, nbins * bits.bin_size)
    return c, {"lines": len(s1), "ranges": len(s2), "overlapping": int((counts >= 1).sum()),
               "sum_counts": int(counts.astype(np.int64).sum()), "covered": bits.count_range(0, addressable),
               "addressable": addressable, "sha256": sha(counts)}


def golden_c4(pool_size):
    _G["f2"] = synth.genome_intervals(50_000_000, 4002)
    _G["f1"] = synth.genome_intervals(50_000_000, 4001)
    order = sorted(range(24), key=lambda c: -len(_G["f2"][c][0]))
    with mp.get_context("fork").Pool(pool_size) as pool:
        res = dict(pool.imap_unordered(_c4_worker, order))
    return [res[c] for c in range(24)]


# ---- C5 ---------------------------------------------------------------------------------------------------------------
def _ref_script():
    import bx
    import bx.intervals
    for pkg, sub in ((bx, ""), (bx.intervals, "intervals")):
        d = os.path.join(REFERENCE, "lib", "bx", sub)
        if d not in pkg.__path__:
            pkg.__path__.append(d)
    if os.path.join(REFERENCE, "lib") not in sys.path:
        sys.path.append(os.path.join(REFERENCE, "lib"))
    spec = importlib.util.spec_from_file_location(
        "ref_aggregate_script", os.path.join(REFERENCE, "scripts", "aggregate_scores_in_intervals.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _c5_worker(c):
    orc.ref_modules()
    mod = _ref_script()
    origin, v, ws, we = _G["tracks"][c]
    name = synth.HG38_NAMES[c]
    with tempfile.TemporaryDirectory(dir="/tmp") as d:
        wig, bed, out = (os.path.join(d, n) for n in ("s.wig", "w.bed", "o.txt"))
        with open(wig, "w") as f:
            f.write(f"fixedStep chrom={name} start={origin + 1} step=1\n")
            f.write("\n".join(map(repr, v.astype(np.float64).tolist())))
            f.write("\n")
        with open(bed, "w") as f:
            f.write("".join(f"{name}\t{a}\t{b}\n" for a, b in zip(ws.tolist(), we.tolist())))
        old = sys.argv
        sys.argv = [mod.__file__, wig, bed, out]
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                mod.main()
        finally:
            sys.argv = old
        cols = [ln.rstrip("\n").split("\t")[3:] for ln in open(out)]
    avg = np.array([np.float32(r[0]) for r in cols], np.float32)
    mn = np.array([np.float32(r[1]) for r in cols], np.float32)
    mx = np.array([np.float32(r[2]) for r in cols], np.float32)
    return c, {"scores": len(v), "windows": len(cols), "origin": origin,
               "nan_lines": int(sum(1 for r in cols if r[0] == "nan")), "sha256": sha(avg, mn, mx)}


def golden_c5(pool_size):
    _G["tracks"] = synth.genome_scores(100_000_000, 5_000_000, 5001)
    order = sorted(range(24), key=lambda c: -len(_G["tracks"][c][1]))
    with mp.get_context("fork").Pool(pool_size) as pool:
        res = dict(pool.imap_unordered(_c5_worker, order))
    return [res[c] for c in range(24)]


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if a in ("c2", "c3", "c4", "c5")] or ["c2", "c3", "c4", "c5"]
    orc.build_ref(REFERENCE)
    nproc = min(8, os.cpu_count() or 1)
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in which:
        t0 = time.time()
        out[name] = {"c2": golden_c2, "c3": golden_c3, "c4": golden_c4, "c5": golden_c5}[name](nproc)
        print(f"{name}: {time.time() - t0:.0f} s", flush=True)
        json.dump(out, open(OUT, "w"), indent=1)
    print("wrote", OUT)
