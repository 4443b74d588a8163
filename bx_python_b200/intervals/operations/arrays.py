"""
Array-in / array-out forms of the bit-set operations of ``bx.intervals.operations``
(``/root/reference/lib/bx/intervals/operations/{intersect,subtract,merge,complement,coverage,base_coverage}.py``).

The reference functions are generators over reader objects that call ``count_range`` / ``next_set`` / ``next_clear`` once
per interval (they keep working unmodified on the shadowed ``bx.bitset``, see bx_python_b200.shadow).  Here one call
handles a whole file: the secondary files become bit sets with one launch each (``IntervalTable.binned_bitsets``, the
semantics of ``GenomicIntervalReader.binned_bitsets``), the set algebra is one launch per genome (``iand_many`` /
``ior_many``), the per-interval ``count_range`` one launch (``count_ranges_many``) and the per-interval pieces one ranged
run extraction per chromosome (``BinnedBitSet.runs_in_ranges``).  Results are column arrays in the reference's output
order (the order of the primary file; chromosomes in first-seen order for merge / complement):

    src    index of the primary line an output row derives from (the row keeps that line's other fields)
    start, end   the row's coordinates

``skipped`` lists the primary lines the reference would count in ``primary.skipped`` *before* emitting anything for them
(start > end, or an IndexError from ``count_range``); the reference additionally counts a line as skipped -- after having
emitted its pieces -- when its piece generator runs off the end of the bit set (operations/__init__.py:10-33).
"""
from __future__ import annotations

import numpy as np

from ...bitset import MAX, iand_many, ior_many
from ...bitset import count_ranges_many
from . import MAX_END


def _and_into(bitsets, other):
    keys = [c for c in bitsets if c in other]
    for c in keys:
        bitsets[c]._check_same(other[c])
    iand_many([bitsets[c] for c in keys], [other[c] for c in keys])


def _or_into(bitsets, other, adopt):
    keys = [c for c in other if c in bitsets]
    ior_many([bitsets[c] for c in keys], [other[c] for c in keys])
    if adopt:                                          # subtract.py:33-37: a chromosome only the later file has is adopted
        for c in other:
            if c not in bitsets:
                bitsets[c] = other[c]


def _secondary_bitsets(others, lens, combine):
    """First secondary file through the BitsetSafeReaderWrapper filter, the rest through the plain builder (as the
    reference does, e.g. intersect.py:28-37), combined into the first."""
    bitsets, _ = others[0].binned_bitsets(lens=lens, safe=True)
    for t in others[1:]:
        combine(bitsets, t.binned_bitsets(lens=lens))
    return bitsets


def _primary_counts(primary, bitsets):
    """-> (set index per line or -1, usable mask, counts) with the reference's IndexError cases masked out."""
    names = list(bitsets)
    sets = [bitsets[c] for c in names]
    idx = {c: k for k, c in enumerate(names)}
    remap = np.asarray([idx.get(c, -1) for c in primary.names] or [0], np.int32)
    which = remap[primary.chrom] if len(primary) else np.zeros(0, np.int32)
    s, e = primary.start, primary.end
    sizes = np.asarray([b.size for b in sets] or [0], np.int64)
    size_of = np.where(which >= 0, sizes[np.clip(which, 0, max(len(sets) - 1, 0))], 0)
    # count_range(start, end - start) raises IndexError for start < 0, start >= size, end < start, end > size
    bad = (which >= 0) & ((s > e) | (s < 0) | (s >= size_of) | (e > size_of))
    ok = (which >= 0) & ~bad
    counts = np.zeros(len(primary), np.int32)
    sel = np.nonzero(ok)[0]
    if len(sel) and sets:
        counts[sel] = count_ranges_many(sets, which[sel], s[sel], (e - s)[sel])
    return which, ok, bad, counts, sets


def _pieces(sets, which, sel, s, e, val):
    """Runs of bits == val inside the selected lines' ranges -> (line index per run, run starts, run ends), line order."""
    src, rs, re = [], [], []
    for k in np.unique(which[sel]).tolist():
        lines = sel[which[sel] == k]
        off, a, b = sets[k].runs_in_ranges(s[lines], e[lines], val)
        src.append(np.repeat(lines, np.diff(off)))
        rs.append(a)
        re.append(b)
    if not src:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64)
    src, rs, re = np.concatenate(src), np.concatenate(rs).astype(np.int64), np.concatenate(re).astype(np.int64)
    order = np.argsort(src, kind="stable")             # back to the primary file's order (runs stay in position order)
    return src[order], rs[order], re[order]


def _merge_rows(parts):
    src = np.concatenate([p[0] for p in parts])
    order = np.argsort(src, kind="stable")
    return tuple(np.concatenate([p[k] for p in parts])[order] for k in range(3))


def intersect(primary, others, mincols=1, pieces=True, lens=None):
    """operations/intersect.py:19-84.  -> dict(src, start, end, skipped)."""
    bitsets = _secondary_bitsets(list(others), lens or {}, _and_into)
    which, ok, bad, counts, sets = _primary_counts(primary, bitsets)
    hit = np.nonzero(ok & (counts >= mincols))[0]
    if pieces:
        src, rs, re = _pieces(sets, which, hit, primary.start, primary.end, 1)
    else:
        src, rs, re = hit, primary.start[hit], primary.end[hit]
    return {"src": src, "start": rs, "end": re, "skipped": np.nonzero(bad)[0]}


def subtract(primary, others, mincols=1, pieces=True, lens=None):
    """operations/subtract.py:22-79.  -> dict(src, start, end, skipped)."""
    bitsets = _secondary_bitsets(list(others), lens or {}, lambda a, b: _or_into(a, b, True))
    which, ok, bad, counts, sets = _primary_counts(primary, bitsets)
    # chromosome without a bit set: yielded as is (a line with end < start never leaves the reader: ParseError, io.py:66-67)
    untouched = np.nonzero((which < 0) & (primary.start <= primary.end))[0]
    below = np.nonzero(ok & (counts < mincols))[0]                     # not enough overlap: the whole interval
    parts = [(untouched, primary.start[untouched], primary.end[untouched]),
             (below, primary.start[below], primary.end[below])]
    if pieces:
        parts.append(_pieces(sets, which, np.nonzero(ok & (counts >= mincols))[0], primary.start, primary.end, 0))
    src, rs, re = _merge_rows(parts)
    return {"src": src, "start": rs, "end": re, "skipped": np.nonzero(bad | ((which < 0) & (primary.start > primary.end)))[0]}


def merge(table, mincols=1):
    """operations/merge.py:13-38 (mincols is accepted and unused there too).  -> dict(chrom ids into `names`, start, end)."""
    bitsets, _ = table.binned_bitsets(lens={}, safe=True)
    names = list(bitsets)
    cid, rs, re = [], [], []
    for k, c in enumerate(names):
        a, b = bitsets[c].runs()                                       # bits_set_in_range(bitset, 0, MAX_END)
        keep = a < MAX_END
        cid.append(np.full(int(keep.sum()), k, np.int32))
        rs.append(a[keep])
        re.append(np.minimum(b[keep], MAX_END))
    cat = (lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt))
    return {"names": names, "chrom": cat(cid, np.int32), "start": cat(rs, np.int64), "end": cat(re, np.int64)}


def complement(table, lens):
    """operations/complement.py:13-57.  -> dict(names, chrom, start, end)."""
    bitsets, _ = table.binned_bitsets(lens=lens, safe=True)
    names = list(bitsets)
    cid, rs, re = [], [], []
    for k, c in enumerate(names):
        bits = bitsets[c]
        bits.invert()
        limit = min(lens.get(c, MAX), bits.size)
        a, b = bits.runs()
        keep = a < limit
        cid.append(np.full(int(keep.sum()), k, np.int32))
        rs.append(a[keep])
        re.append(np.minimum(b[keep], limit))
    cat = (lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt))
    return {"names": names, "chrom": cat(cid, np.int32), "start": cat(rs, np.int64), "end": cat(re, np.int64)}


def coverage(primary, others):
    """operations/coverage.py:17-78.  -> dict(src, bases_covered, percent, skipped); rows in primary order."""
    def or_common(a, b):
        _or_into(a, b, False)                                          # coverage.py:29-32: only chromosomes both have
    bitsets = _secondary_bitsets(list(others), {}, or_common)
    which, ok, bad, counts, _ = _primary_counts(primary, bitsets)
    inverted = primary.start > primary.end                             # skipped before the chromosome is even looked up
    rows = np.nonzero((ok | (which < 0)) & ~inverted)[0]
    length = (primary.end - primary.start)[rows].astype(np.float64)
    covered = counts[rows].astype(np.int64)
    percent = np.divide(covered, length, out=np.zeros(len(rows)), where=length != 0)
    return {"src": rows, "bases_covered": covered, "percent": percent, "skipped": np.nonzero(bad | inverted)[0]}


def base_coverage(table):
    """operations/base_coverage.py:9-24: number of bases covered by the file's intervals."""
    bitsets, _ = table.binned_bitsets(lens={}, safe=True)
    return int(sum(b.count_range(0, min(MAX_END, b.size)) for b in bitsets.values()))
