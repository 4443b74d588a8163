"""
Drop-in for ``bx.bitset_builders`` (``/root/reference/lib/bx/bitset_builders.py``): dictionaries of ``BinnedBitSet``
built from interval text / lists.

Same function names, arguments, per-line semantics and exceptions as the reference; the difference is that the
``set_range`` calls of a whole file are collected and issued as ONE batched kernel (``bitset.set_ranges_many``; a
single chromosome: ``BinnedBitSet.set_ranges``) instead of one C call per line.  Text parsing stays on the host (it is Python in the reference too).
"""
from __future__ import annotations

import re
from warnings import warn

import numpy as np

from .bitset import MAX, BinnedBitSet, set_ranges_many


class _Collector:
    """Per-chromosome (start, count) lists; validates each range like BinnedBitSet.set_range does, in file order."""

    def __init__(self):
        self.bitsets, self._s, self._c = {}, {}, {}

    def get(self, chrom, size):
        if chrom not in self.bitsets:
            self.bitsets[chrom] = BinnedBitSet(size)
            self._s[chrom], self._c[chrom] = [], []
        return self.bitsets[chrom]

    def add(self, chrom, start, count):
        self.bitsets[chrom]._check_range_count(start, count)      # the reference raises on the offending line
        self._s[chrom].append(start)
        self._c[chrom].append(count)

    def finish(self):
        chroms = [c for c in self.bitsets if self._s[c]]
        if len(chroms) == 1:
            c = chroms[0]
            self.bitsets[c].set_ranges(np.asarray(self._s[c], np.int32), np.asarray(self._c[c], np.int32))
        elif chroms:                                   # the whole file in one launch
            which = np.concatenate([np.full(len(self._s[c]), k, np.int32) for k, c in enumerate(chroms)])
            starts = np.concatenate([np.asarray(self._s[c], np.int32) for c in chroms])
            counts = np.concatenate([np.asarray(self._c[c], np.int32) for c in chroms])
            set_ranges_many([self.bitsets[c] for c in chroms], which, starts, counts)
        return self.bitsets


def binned_bitsets_from_file(f, chrom_col=0, start_col=1, end_col=2, strand_col=5, upstream_pad=0, downstream_pad=0,
                             lens={}):
    """bitset_builders.py:17-54."""
    col = _Collector()
    size = MAX
    last_chrom = None
    for line in f:
        if line.startswith("#") or line.isspace():
            continue
        fields = line.split()
        chrom = fields[chrom_col]
        if chrom != last_chrom:
            if chrom not in col.bitsets:
                size = lens[chrom] if chrom in lens else MAX
                col.get(chrom, size)
            last_chrom = chrom
        start, end = int(fields[start_col]), int(fields[end_col])
        if upstream_pad:
            start = max(0, start - upstream_pad)
        if downstream_pad:
            end = min(size, end + downstream_pad)
        if start > end:
            warn("Interval start after end!")
        col.add(chrom, start, end - start)
    return col.finish()


def binned_bitsets_from_bed_file(f, chrom_col=0, start_col=1, end_col=2, strand_col=5, upstream_pad=0, downstream_pad=0,
                                 lens={}):
    """bitset_builders.py:57-104 (honours `track ... offset=N`, skips `browser` lines)."""
    col = _Collector()
    size = MAX
    last_chrom = None
    offset = 0
    for line in f:
        if line.startswith("#") or line.isspace() or line.startswith("browser"):
            continue
        if line.startswith("track"):
            m = re.search(r"offset=(\d+)", line)
            if m and m.group(1):
                offset = int(m.group(1))
            continue
        fields = line.split()
        chrom = fields[chrom_col]
        if chrom != last_chrom:
            if chrom not in col.bitsets:
                size = lens[chrom] if chrom in lens else MAX
                col.get(chrom, size)
            last_chrom = chrom
        start, end = int(fields[start_col]) + offset, int(fields[end_col]) + offset
        if upstream_pad:
            start = max(0, start - upstream_pad)
        if downstream_pad:
            end = min(size, end + downstream_pad)
        if start > end:
            warn("Interval start after end!")
        col.add(chrom, start, end - start)
    return col.finish()


def binned_bitsets_proximity(f, chrom_col=0, start_col=1, end_col=2, strand_col=5, upstream=0, downstream=0):
    """bitset_builders.py:107-139."""
    col = _Collector()
    for line in f:
        if line.startswith("#"):
            continue
        fields = line.split()
        strand = "+"
        if len(fields) >= strand_col + 1 and fields[strand_col] == "-":
            strand = "-"
        chrom = fields[chrom_col]
        col.get(chrom, MAX)
        start, end = int(fields[start_col]), int(fields[end_col])
        if strand == "+":
            if upstream:
                start = max(0, start - upstream)
            if downstream:
                end = min(MAX, end + downstream)
        if strand == "-":
            if upstream:
                end = min(MAX, end + upstream)
            if downstream:
                start = max(0, start - downstream)
        if end - start > 0:
            col.add(chrom, start, end - start)
    return col.finish()


def binned_bitsets_from_list(list=[]):
    """bitset_builders.py:142-156."""
    col = _Collector()
    for item in list:
        chrom = item[0]
        col.get(chrom, MAX)
        start, end = int(item[1]), int(item[2])
        col.add(chrom, start, end - start)
    return col.finish()


def binned_bitsets_by_chrom(f, chrom, chrom_col=0, start_col=1, end_col=2):
    """bitset_builders.py:159-169."""
    col = _Collector()
    bitset = col.get(chrom, MAX)
    for line in f:
        if line.startswith("#"):
            continue
        fields = line.split()
        if fields[chrom_col] == chrom:
            start, end = int(fields[start_col]), int(fields[end_col])
            col.add(chrom, start, end - start)
    col.finish()
    return bitset


def binned_bitsets_from_arrays(chrom_ids, starts, ends, sizes):
    """Array form (no reference equivalent): chrom_ids index into `sizes`; one batched set_ranges per chromosome.
    -> list of BinnedBitSet, one per entry of sizes."""
    chrom_ids, starts, ends = np.asarray(chrom_ids), np.asarray(starts), np.asarray(ends)
    out = []
    for c, size in enumerate(sizes):
        b = BinnedBitSet(int(size))
        sel = chrom_ids == c
        if sel.any():
            b.set_ranges(starts[sel], ends[sel] - starts[sel])
        out.append(b)
    return out
