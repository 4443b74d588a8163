"""
Seeded synthetic inputs shared by bench.py, the parity tests and tests/golden/make_golden.py.

All generators use numpy.random.default_rng(seed) (PCG64) and int32/float32 arrays, so the CPU checker and the
CUDA path see byte-identical inputs.  Pure NumPy; no device code.
"""
from __future__ import annotations

import numpy as np

# UCSC hg38.chrom.sizes, primary assembly (chr1..22, X, Y)
HG38 = [
    ("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
    ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422),
    ("chr11", 135086622), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189),
    ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
    ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415),
]
HG38_NAMES = [c for c, _ in HG38]
HG38_LENS = np.array([n for _, n in HG38], np.int64)


def uniform_intervals(rng, n, G, max_len=2000):
    """start ~ U[0, G-max_len), len ~ U{1..max_len}; returns int32 (start, end)."""
    s = rng.integers(0, max(1, G - max_len), n, dtype=np.int64)
    ln = rng.integers(1, max_len + 1, n, dtype=np.int64)
    return s.astype(np.int32), (s + ln).astype(np.int32)


# ---- C1: 10k vs 10k on one chromosome ---------------------------------------------------------------------------
def c1_intervals(n=10_000, nq=10_000, G=1_000_000):
    s, e = uniform_intervals(np.random.default_rng(1001), n, G)
    qs, qe = uniform_intervals(np.random.default_rng(1002), nq, G)
    return s, e, qs, qe


def edge_sets(k=8, n=400, nq=300):
    """Zero-length, inverted, duplicate and negative-coordinate intervals; reversed and empty queries."""
    rng = np.random.default_rng(1003)
    out = []
    for _ in range(k):
        G = int(rng.integers(5, 61))
        s = rng.integers(-3, G + 1, n)
        e = s + rng.integers(-2, 6, n)
        qs = rng.integers(-5, G + 6, nq)
        qe = qs + rng.integers(-3, 12, nq)
        out.append(tuple(a.astype(np.int32) for a in (s, e, qs, qe)))
    return out


def neighbor_case(seed):
    rng = np.random.default_rng(1100 + seed)
    n = int(rng.integers(1, 300))
    G = int(rng.integers(20, 5000))
    s = rng.integers(0, G, n)
    e = s + rng.integers(0, 40, n)
    queries = [(int(rng.integers(-5, G + 50)), int(rng.integers(1, 6)), int(rng.choice([0, 1, 7, 60, 2500])))
               for _ in range(40)]
    return s.astype(np.int32), e.astype(np.int32), queries


# ---- C2 / C4: hg38-shaped genome-wide interval sets ---------------------------------------------------------------
def genome_intervals(n, seed, max_len=2000, chrom_lens=HG38_LENS):
    """-> list over chromosomes of (start int32[], end int32[]); chromosome ~ multinomial(length)."""
    rng = np.random.default_rng(seed)
    counts = rng.multinomial(n, chrom_lens / chrom_lens.sum())
    out = []
    for L, c in zip(chrom_lens.tolist(), counts.tolist()):
        out.append(uniform_intervals(rng, c, L, max_len))
    return out


# ---- bitsets ------------------------------------------------------------------------------------------------------
def bitset_case(seed):
    """-> size, granularity, ops, probes.  ops are tuples understood by apply_bitset_op."""
    rng = np.random.default_rng(1200 + seed)
    size = int(rng.integers(1, 5000))
    gran = int(rng.choice([1, 2, 3, 7, 10, 64, 1024]))
    ops = []
    for _ in range(int(rng.integers(1, 30))):
        k = int(rng.integers(0, 2))
        op = int(rng.integers(0, 6))
        if op == 0:
            s = int(rng.integers(0, size))
            c = int(rng.integers(0, min(size - s, max(1, size // 3)) + 1))
            ops.append(("set_range", k, s, c))
        elif op == 1:
            ops.append(("set", k, int(rng.integers(0, size))))
        elif op == 2:
            ops.append(("clear", k, int(rng.integers(0, size))))
        elif op == 3:
            ops.append(("invert", k))
        elif op == 4:
            ops.append(("iand", k))
        else:
            ops.append(("ior", k))
    probes = []
    for _ in range(60):
        s = int(rng.integers(0, size))
        probes.append((s, int(rng.integers(0, size - s + 1))))
    return size, gran, ops, probes


def apply_bitset_op(b, op):
    """b = [bitset0, bitset1] of any class exposing the BinnedBitSet API."""
    name, k = op[0], op[1]
    if name == "set_range":
        b[k].set_range(op[2], op[3])
    elif name == "set":
        b[k].set(op[2])
    elif name == "clear":
        b[k].clear(op[2])
    elif name == "invert":
        b[k].invert()
    elif name == "iand":
        b[k].iand(b[1 - k])
    elif name == "ior":
        b[k].ior(b[1 - k])
    else:
        raise ValueError(name)


def c3_case(size, nranges, seed, nq=2000, max_len=2000):
    """Two range lists (start,count) for operands A and B plus nq count_range probes (start,count)."""
    ra = np.random.default_rng(3000 + seed)
    rb = np.random.default_rng(3100 + seed)
    rq = np.random.default_rng(3200 + seed)

    def ranges(r, n):
        s = r.integers(0, size - max_len, n)
        c = r.integers(1, max_len + 1, n)
        return s.astype(np.int32), c.astype(np.int32)
    return ranges(ra, nranges), ranges(rb, nranges), ranges(rq, nq)


# ---- aggregate_scores_in_intervals ---------------------------------------------------------------------------------
def aggregate_scores(rng, n):
    """float32 N(0,1) with 1 % exact zeros and 1 % NaN."""
    v = rng.normal(0.0, 1.0, n).astype(np.float32)
    u = rng.random(n)
    v[u < 0.01] = 0.0
    v[(u >= 0.01) & (u < 0.02)] = np.nan
    return v


def aggregate_case(seed, n=3000, nw=400):
    """-> origin, scores float32[n] (position origin+i), ws, we (absolute coords), mask_runs or None."""
    rng = np.random.default_rng(5000 + seed)
    origin = int(rng.integers(0, 1000))
    v = aggregate_scores(rng, n)
    if seed % 3 == 1:                       # large magnitudes: exercises float32 rounding order
        v[rng.integers(0, n, 20)] = np.float32(16777216.0)
    if seed % 3 == 2:
        v *= np.float32(1e-3)
    ws = rng.integers(max(0, origin - 30), origin + n + 10, nw)
    we = ws + rng.integers(0, 60, nw)
    mask_runs = None
    if seed % 2 == 1:
        ms = np.sort(rng.integers(origin, origin + n - 20, 40))
        mask_runs = [(int(a), int(a + rng.integers(1, 20))) for a in ms]
    return origin, v, ws.astype(np.int32), we.astype(np.int32), mask_runs


def genome_scores(n_total, nw_total, seed, chrom_lens=HG38_LENS, mask_density=0.0):
    """C5: per chromosome (origin, scores float32[], ws int32[], we int32[]); windows len ~ U{1..40}."""
    rng = np.random.default_rng(seed)
    frac = chrom_lens / chrom_lens.sum()
    ns = rng.multinomial(n_total, frac)
    nws = rng.multinomial(nw_total, frac)
    out = []
    for L, n, nw in zip(chrom_lens.tolist(), ns.tolist(), nws.tolist()):
        origin = int(rng.integers(0, max(1, L - n)))
        v = aggregate_scores(rng, n)
        ws = rng.integers(origin, max(origin + 1, origin + n - 40), nw, dtype=np.int64)
        we = ws + rng.integers(1, 41, nw, dtype=np.int64)
        out.append((origin, v, ws.astype(np.int32), we.astype(np.int32)))
    return out


# ---- text inputs for the bitset_builders / bitset_utils drop-ins ------------------------------------------------------
def bed_lines(seed, n=400):
    """-> (lines, lens): BED-like text incl. comments, blank lines, a strand column, unsorted chromosomes."""
    rng = np.random.default_rng(6000 + seed)
    lens = {"chr1": 50000, "chr2": 20000, "chrX": 7777}
    names = list(lens) + ["chrUn"]            # chrUn has no length -> MAX-sized bitset
    lines = ["# header", "\n"]
    for _ in range(n):
        c = names[int(rng.integers(0, len(names)))]
        L = lens.get(c, 100000)
        s = int(rng.integers(0, L - 300))
        e = s + int(rng.integers(0, 300))
        strand = "+-"[int(rng.integers(0, 2))]
        lines.append(f"{c}\t{s}\t{e}\tname\t0\t{strand}\n")
    return lines, lens


def exon_lists(seed):
    rng = np.random.default_rng(6100 + seed)

    def one():
        s = np.sort(rng.integers(0, 5000, int(rng.integers(1, 30))))
        return [(int(a), int(a + rng.integers(1, 120))) for a in s]
    return one(), one()


def ops_case(seed, n=300):
    """Three BED line lists + a lens dict for the interval operations (intersect / subtract / merge / complement /
    coverage / base_coverage): overlapping intervals, zero-length ones, a few with end < start (dropped by the readers), a
    few beyond the chromosome length (dropped by BitsetSafeReaderWrapper), a chromosome without a length (MAX-sized bit
    set) and one that only the primary file has."""
    rng = np.random.default_rng(6500 + seed)
    lens = {"chr1": 30000, "chr2": 12000}

    def bed(m, names, tag, at_end=False):
        out = ["# a comment line"]
        for i in range(m):
            c = names[int(rng.integers(0, len(names)))]
            L = lens.get(c, 20000)
            s = int(rng.integers(0, L - 200))
            e = s + int(rng.integers(0, 200))
            r = rng.random()
            if r < 0.03:
                s, e = e + 5, s                       # end < start
            elif r < 0.06:
                e = L + int(rng.integers(1, 50))      # runs past the chromosome
            elif r < 0.08 and at_end:
                s = e = L                             # empty interval exactly at the end: fine as a primary line, but as a
                                                      # bit-set line the reference dies with set_range's IndexError
            out.append(f"{c}\t{s}\t{e}\t{tag}{i}\t0\t{'+-'[int(rng.integers(0, 2))]}")
        return out
    primary = bed(n, ["chr1", "chr2", "chrX", "chrOnlyPrimary"], "p", at_end=True)
    second = bed(n, ["chr1", "chr2", "chrX"], "s")
    third = bed(n // 2, ["chr1", "chrX", "chrOnlyThird"], "t")
    return primary, second, third, lens


# ---- score sources / summary / join -------------------------------------------------------------------
def wiggle_text(seed, n=6000, chroms=("chr1", "chr2")):
    """A wiggle file mixing fixedStep / variableStep blocks (with and without span), a leading bedGraph-style
    section, comments and blank lines; blocks overlap on purpose so that file order decides the final score."""
    rng = np.random.default_rng(7000 + seed)
    lines = ["track type=wiggle_0 name=synthetic", "# comment"]
    if seed % 2:                                     # bed mode comes first (before any declaration)
        for _ in range(int(rng.integers(1, 30))):
            s = int(rng.integers(0, n - 50))
            lines.append(f"{chroms[int(rng.integers(0, len(chroms)))]}\t{s}\t{s + int(rng.integers(0, 40))}\t{rng.normal():.4f}")
        lines.append("chr1\t5\t9")                   # too few fields: ignored by the reader
    for _ in range(int(rng.integers(2, 8))):
        chrom = chroms[int(rng.integers(0, len(chroms)))]
        kind = int(rng.integers(0, 3))
        span = int(rng.integers(1, 7))
        if kind == 0:
            step = int(rng.integers(1, 6))
            cnt = int(rng.integers(0, 150))
            start = int(rng.integers(1, n - cnt * step - span - 1))
            lines.append(f"fixedStep chrom={chrom} start={start} step={step}" + (f" span={span}" if span > 1 else ""))
            for _ in range(cnt):
                r = rng.random()
                lines.append("nan" if r < 0.03 else "0" if r < 0.06 else repr(float(np.float32(rng.normal()))))
            if rng.random() < 0.5:
                lines.append("")
        else:
            lines.append(f"variableStep chrom={chrom}" + (f" span={span}" if kind == 1 else ""))
            for p in rng.integers(1, n - 10, int(rng.integers(0, 150))).tolist():
                lines.append(f"{p} {rng.normal():.5f}")
    return "\n".join(lines) + "\n"


def summarize_case(seed):
    """(start, end, val, region start, region end, size): sorted disjoint batches (what a bigWig holds) for even
    seeds, arbitrary overlapping / unsorted / empty / clipped ones for odd seeds."""
    rng = np.random.default_rng(8000 + seed)
    n = int(rng.integers(0, 400))
    rs = int(rng.integers(0, 1000))
    re_ = rs + int(rng.integers(1, 5000))
    size = int(rng.integers(1, 64))
    if seed % 2 == 0:
        gaps, lens = rng.integers(0, 40, n), rng.integers(1, 90, n)
        s = np.maximum(rs - 150 + np.cumsum(gaps + np.concatenate([[0], lens[:-1]])), 0)
        e = s + lens
    else:
        s = rng.integers(max(rs - 200, 0), re_ + 200, n)
        e = np.maximum(s + rng.integers(-2, 150, n), 0)
    return s.astype(np.int32), e.astype(np.int32), rng.normal(0, 3, n).astype(np.float32), rs, re_, size


def join_case(seed, nl=120, nr=100):
    """Two BED line lists (4 columns) over three chromosomes, with zero-length intervals, duplicates and a comment."""
    rng = np.random.default_rng(9000 + seed)

    def bed(n, tag):
        out = []
        for i in range(n):
            s = int(rng.integers(0, 800))
            out.append(f"chr{int(rng.integers(1, 4))}\t{s}\t{s + int(rng.integers(0, 70))}\t{tag}{i}")
        return out
    left, right = bed(nl, "L"), bed(nr, "R")
    left.insert(3, "# a comment in the left file")
    right.append("chr9\t10\t20\tRlonely")
    left.append("chr8\t10\t20\tLlonely")
    return left, right, int(rng.integers(1, 12))
