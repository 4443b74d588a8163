"""
``bx.intervals.operations.quicksect`` (lib/bx/intervals/operations/quicksect.py:11-125) on the device index.

The reference is a pure-Python treap per chromosome: ``insert(interval, linenum, other)``, ``intersect(interval,
report_func)`` reporting every node with ``start < node.end and end > node.start`` (:115-121) and ``traverse(func)``
(in-order, :123-129).  Here every chromosome is one tree of a single ``IntervalForest`` (built lazily, rebuilt after
further inserts), ``intersect`` is a device ``find`` and ``intersect_batch`` answers many intervals in one launch.

Order of the reported nodes: the reference walks its randomly balanced treap in PRE-order, so the order differs from
run to run of the reference itself; nodes are reported here in the index's in-order sequence.  ``traverse`` keeps the
reference's deterministic in-order: by start, and among equal starts the LATER insert first (``start > self.start``
goes right, ties go left, :52-70), chromosomes in order of first insert.
"""
from __future__ import annotations

import numpy as np

from ..intersection import IntervalForest


class IntervalNode:
    """What ``report_func`` / ``traverse`` receive: the stored interval (attributes set on it persist, as
    join.py:54 relies on)."""
    __slots__ = ("start", "end", "linenum", "other", "__dict__")

    def __init__(self, start, end, linenum=0, other=None):
        self.start, self.end, self.linenum, self.other = start, end, linenum, other


class IntervalTree:
    def __init__(self):
        self.chroms = {}                # chrom -> tree id (the reference maps chrom -> root node)
        self._nodes = []
        self._tid = []
        self._forest = None

    def insert(self, interval, linenum=0, other=None):
        tid = self.chroms.setdefault(interval.chrom, len(self.chroms))
        self._nodes.append(IntervalNode(interval.start, interval.end, linenum, other))
        self._tid.append(tid)
        self._forest = None

    def insert_many(self, chroms, starts, ends, linenums=None, others=None):
        for i, (c, s, e) in enumerate(zip(chroms, starts, ends)):
            tid = self.chroms.setdefault(c, len(self.chroms))
            self._nodes.append(IntervalNode(int(s), int(e), 0 if linenums is None else linenums[i],
                                            None if others is None else others[i]))
            self._tid.append(tid)
        self._forest = None

    def _arrays(self):
        return (np.asarray(self._tid, np.int32), np.asarray([n.start for n in self._nodes], np.int64),
                np.asarray([n.end for n in self._nodes], np.int64))

    def _ensure(self):
        if self._forest is None:
            tid, s, e = self._arrays()
            self._forest = IntervalForest(max(len(self.chroms), 1)).build(tid, s, e)
            self._start32, self._end32 = s.astype(np.int32), e.astype(np.int32)
        return self._forest

    def intersect_batch(self, chroms, starts, ends):
        """CSR (offsets, node indices) of the nodes each query interval intersects; unknown chromosomes hit nothing."""
        qt = np.asarray([self.chroms.get(c, -1) for c in chroms], np.int32)
        if not self._nodes:
            return np.zeros(len(qt) + 1, np.int64), np.empty(0, np.int32)
        return self._ensure().find_batch(qt, starts, ends)

    def intersect(self, interval, report_func):
        if interval.chrom not in self.chroms:
            return
        _, hits = self.intersect_batch([interval.chrom], [interval.start], [interval.end])
        for i in hits.tolist():
            report_func(self._nodes[i])

    def _inorder(self):
        if not self._nodes:
            return []
        tid, s, _ = self._arrays()
        return np.lexsort((-np.arange(len(s)), s, tid)).tolist()

    def traverse(self, func):
        for i in self._inorder():
            func(self._nodes[i])
