"""
``bx.intervals.operations.join.join`` (lib/bx/intervals/operations/join.py:14-75) with the candidate search, the
overlap arithmetic (:35-50), the ``mincols`` filter and the ``visited`` marks on the device (``bxg_itree_join``).

Same generator contract: items of ``leftSet`` that are not intervals pass through in place; every interval yields one
output row per right-hand interval whose overlap is >= ``mincols`` (``list(interval) + item.fields``), or, with
``rightfill``, one row padded with ``"."`` when nothing qualifies; with ``leftfill`` the right-hand intervals that
were never matched follow at the end (in-order traversal of the quicksect trees).  The reference streams
``leftSet``; here it is read completely first so that all of it is answered by one launch -- the rows come out in the
same left order.  Within one left interval the reference's row order is the pre-order of a randomly balanced treap
(not reproducible between its own runs); rows are emitted in index order instead.

"Is an interval" is duck-typed (``chrom``, ``start``, ``end``, ``fields`` attributes) because ``GenomicInterval``
(lib/bx/intervals/io.py) is a text-reader class outside this package; the reference's readers satisfy it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ... import _lib
from ..._lib import as_i32, check, ptr
from .quicksect import IntervalTree


class BedRow(list):
    """Minimal stand-in for ``GenomicInterval`` (lib/bx/intervals/io.py): the list of text fields with ``chrom`` /
    ``start`` / ``end`` / ``fields`` / ``nfields`` attributes -- all ``join`` touches."""

    def __init__(self, fields, chrom_col=0, start_col=1, end_col=2):
        super().__init__(fields)
        self.fields, self.nfields = list(fields), len(fields)
        self.chrom, self.start, self.end = fields[chrom_col], int(fields[start_col]), int(fields[end_col])


class BedRows:
    """Iterate BED-like text lines as ``BedRow`` objects; comment / blank lines pass through as strings.
    ``linenum`` counts the lines consumed so far, like the reference's readers."""

    def __init__(self, lines, **cols):
        self._lines, self._cols, self.linenum = lines, cols, 0

    def __iter__(self):
        for line in self._lines:
            self.linenum += 1
            if line.startswith("#") or not line.strip():
                yield line
            else:
                yield BedRow(line.rstrip("\r\n").split("\t"), **self._cols)


def _is_interval(x):
    return all(hasattr(x, a) for a in ("chrom", "start", "end", "fields"))


def join_arrays(tree, chroms, starts, ends, mincols=1):
    """Array form: -> (pair_offsets int64[nq+1], pair_items int32[total], visited bool[n]) for the left intervals
    (chroms[i], starts[i], ends[i]) against the quicksect ``tree``."""
    nq, n = len(starts), len(tree._nodes)
    if n == 0:
        return np.zeros(nq + 1, np.int64), np.empty(0, np.int32), np.zeros(0, bool)
    forest = tree._ensure()
    qt = np.asarray([tree.chroms.get(c, -1) for c in chroms], np.int32)
    qs, qe = as_i32(starts), as_i32(ends)
    total = C.c_int64()
    check(_lib.lib().bxg_itree_join(forest.handle, ptr(qt), ptr(qs), ptr(qe), nq, ptr(tree._start32), ptr(tree._end32),
                                   int(mincols), _lib.HOST, C.byref(total)))
    poff, items, vis = np.empty(nq + 1, np.int64), np.empty(total.value, np.int32), np.empty(n, np.uint8)
    check(_lib.lib().bxg_itree_join_fetch(ptr(poff), ptr(items), ptr(vis)))
    return poff, items, vis.astype(bool)


def join(leftSet, rightSet, mincols=1, leftfill=True, rightfill=True):
    rightTree = IntervalTree()
    rightlen = 0
    for item in rightSet:
        if _is_interval(item):
            rightTree.insert(item, getattr(rightSet, "linenum", 0), item.fields)
            if rightlen == 0:
                rightlen = item.nfields
    left = list(leftSet)
    idx = [i for i, x in enumerate(left) if _is_interval(x)]
    ivs = [left[i] for i in idx]
    leftlen = ivs[0].nfields if ivs else 0
    poff, items, visited = join_arrays(rightTree, [x.chrom for x in ivs], [x.start for x in ivs], [x.end for x in ivs],
                                       mincols)
    poff = poff.tolist()
    k = 0
    for x in left:
        if not _is_interval(x):
            yield x
            continue
        a, b = poff[k], poff[k + 1]
        k += 1
        for it in items[a:b].tolist():
            out = list(x)
            out.extend(rightTree._nodes[it].other)
            yield out
        if a == b and rightfill:
            yield list(x) + ["."] * rightlen
    if leftfill:
        for i in rightTree._inorder():
            if not visited[i]:
                yield ["."] * leftlen + list(rightTree._nodes[i].other)
