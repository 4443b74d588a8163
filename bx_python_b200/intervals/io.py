"""
Array form of the interval readers' bit-set builder: ``GenomicIntervalReader.binned_bitsets``
(``/root/reference/lib/bx/intervals/io.py:190-216``).

The reference walks a reader line by line and calls ``BinnedBitSet.set_range`` per interval.  Its semantics differ from
``bx.bitset_builders`` (which raises on out-of-range lines): the interval is CLAMPED -- ``start = max(start, 0)``,
``end = min(end, bitset.size)`` (:212-213); a line with ``end < start`` never gets that far (the reader raises ParseError,
:66-67; ``NiceReaderWrapper`` skips it, :233-245), and ``BitsetSafeReaderWrapper`` (:262-289) also skips lines with
``end > lens.get(chrom, MAX)``.  Here lines with ``end < start`` are always skipped (there is no parse step to refuse
them); ``safe=True`` adds the BitsetSafeReaderWrapper filter and reports what it skipped.  Text parsing (``read_bed``) is
a plain host loop, as in the reference.
"""
from __future__ import annotations

import numpy as np

from ..bitset import MAX, BinnedBitSet, set_ranges_many


class IntervalTable:
    """Column arrays of an interval file: ``names`` (chromosome names in first-seen order), ``chrom`` (int32 index into
    names per line), ``start`` / ``end`` (int64 per line), ``fields`` (the split lines, for callers that print them)."""

    def __init__(self, names, chrom, start, end, fields=None):
        self.names = list(names)
        self.chrom = np.ascontiguousarray(chrom, np.int32)
        self.start = np.ascontiguousarray(start, np.int64)
        self.end = np.ascontiguousarray(end, np.int64)
        self.fields = fields
        if not (len(self.chrom) == len(self.start) == len(self.end)):
            raise ValueError("column arrays must have the same length")

    def __len__(self):
        return len(self.chrom)

    @classmethod
    def from_arrays(cls, chroms, starts, ends):
        """chroms: sequence of chromosome names (one per interval)."""
        names, index = [], {}
        ids = np.empty(len(chroms), np.int32)
        for i, c in enumerate(chroms):
            k = index.get(c)
            if k is None:
                k = index[c] = len(names)
                names.append(c)
            ids[i] = k
        return cls(names, ids, starts, ends)

    def binned_bitsets(self, upstream_pad=0, downstream_pad=0, lens=None, safe=False):
        """io.py:190-216 for the whole table in one launch -> {chrom: BinnedBitSet} (chromosomes in the order of their
        first interval).  The reference accepts the pad arguments and ignores them (:190); so does this.

        A line with end < start never reaches the bit sets: GenomicInterval refuses it (:66-67) and NiceReaderWrapper --
        the reader the operations are fed with -- skips it (:233-245); such lines are skipped here in both modes.
        safe=False: whatever ``set_range`` refuses after the clamping raises its IndexError.
        safe=True : BitsetSafeReaderWrapper (:262-289) as well -- lines with end > lens.get(chrom, MAX) are skipped too;
        returns (bitsets, skipped line indices)."""
        lens = lens or {}
        s, e = self.start, self.end
        n = len(s)
        limit = np.asarray([lens.get(c, MAX) for c in self.names] or [MAX], np.int64)
        lim_of = limit[self.chrom] if n else np.zeros(0, np.int64)
        keep = ~(e < s)
        if safe:
            keep &= ~(e > lim_of)
        sel = np.nonzero(keep)[0]
        # a chromosome gets its bit set when its first surviving interval is read (:199-211)
        order = []
        seen = set()
        for k in self.chrom[sel].tolist():
            if k not in seen:
                seen.add(k)
                order.append(k)
        sets = {}
        for k in order:
            size = int(limit[k])
            try:
                sets[k] = BinnedBitSet(size)
            except ValueError as err:
                raise Exception(f"Invalid chrom length {str(size)} in 'lens' dictionary. {str(err)}")
        slot = np.full(max(len(self.names), 1), -1, np.int32)
        slot[order] = np.arange(len(order), dtype=np.int32)
        cs = np.maximum(s[sel], 0)                                        # :212-213
        ce = np.minimum(e[sel], lim_of[sel])
        set_ranges_many([sets[k] for k in order], slot[self.chrom[sel]], cs, ce - cs)   # IndexError as set_range raises it
        bitsets = {self.names[k]: sets[k] for k in order}
        return (bitsets, np.nonzero(~keep)[0]) if safe else bitsets


def read_bed(lines, chrom_col=0, start_col=1, end_col=2, comment_prefixes=("#", "track ")):
    """Interval lines -> IntervalTable (comments / blank lines skipped; tab separated, falling back to whitespace)."""
    names, index, ids, starts, ends, rows = [], {}, [], [], [], []
    for line in lines:
        if not line.strip() or line.startswith(comment_prefixes):
            continue
        f = line.rstrip("\r\n").split("\t")
        if len(f) <= max(chrom_col, start_col, end_col):
            f = line.split()
        c = f[chrom_col]
        k = index.get(c)
        if k is None:
            k = index[c] = len(names)
            names.append(c)
        ids.append(k)
        starts.append(int(f[start_col]))
        ends.append(int(f[end_col]))
        rows.append(f)
    return IntervalTable(names, np.asarray(ids, np.int32), np.asarray(starts, np.int64), np.asarray(ends, np.int64), rows)
