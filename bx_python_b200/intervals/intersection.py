"""
Drop-in for ``bx.intervals.intersection`` (``/root/reference/lib/bx/intervals/intersection.pyx``):
``Interval``, ``IntervalNode``, ``IntervalTree``, ``Intersecter``.

``insert``/``add_interval`` queue (start, end, value) on the host; the first query after a mutation uploads the
int32 arrays and builds the device index (radix sort + implicit max hierarchy, csrc/itree.cu).  ``find`` returns the
stored objects in exactly the reference's order (in-order traversal of its treap, intersection.pyx:180-189).
Bulk callers use ``find_batch`` / ``count_batch``; ``IntervalForest`` holds one tree per chromosome in a single
device index so that millions of (chrom, start, end) queries are answered by one kernel launch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .._lib import as_i32, check, ptr


def _c_int(x):
    """Cython `int` argument coercion (intersection.pyx:388): truncate floats, OverflowError outside int32."""
    v = int(x)
    if v > 0x7FFFFFFF or v < -0x80000000:
        raise OverflowError("value too large to convert to int")
    return v


class Interval:
    """intersection.pyx:274-323."""

    __slots__ = ("start", "end", "value", "chrom", "strand")

    def __init__(self, start, end, value=None, chrom=None, strand=None):
        start, end = _c_int(start), _c_int(end)
        assert start <= end, "start must be less than end"
        self.start, self.end, self.value, self.chrom, self.strand = start, end, value, chrom, strand

    def __repr__(self):
        fstr = "Interval(%d, %d" % (self.start, self.end)
        if self.value is not None:
            fstr += ", value=" + str(self.value)
        return fstr + ")"

    # intersection.pyx:305-323
    def __lt__(self, other): return self.start < other.start or self.end < other.end
    def __eq__(self, other): return self.start == other.start and self.end == other.end
    def __ne__(self, other): return self.start != other.start or self.end != other.end
    def __gt__(self, other): return self.start > other.start or self.end > other.end
    def __le__(self, other): return self == other or self < other
    def __ge__(self, other): return self == other or self > other
    __hash__ = None


class IntervalNode:
    """intersection.pyx:61-268.  Two roles, as in the reference: the object ``traverse`` hands to its callback
    (``start``, ``end``, ``interval``), and -- when used directly -- the root handle of a tree:
    ``root = IntervalNode(s, e, obj); root = root.insert(s2, e2, obj2); root.find(a, b); root.left(p, n, max_dist)``.
    The treap itself does not exist here; a root handle forwards to a device-backed ``IntervalTree``."""

    __slots__ = ("start", "end", "interval", "_tree")

    def __init__(self, start, end, interval, _tree=None):
        self.start, self.end, self.interval = _c_int(start), _c_int(end), interval
        self._tree = _tree

    def __repr__(self):
        return "IntervalNode(%i, %i)" % (self.start, self.end)

    def _root(self):
        if self._tree is None:              # first use as a root: the node's own interval is the first item
            self._tree = IntervalTree()
            self._tree.insert(self.start, self.end, self.interval)
        return self._tree

    def insert(self, start, end, interval):
        """Insert into the tree rooted here; returns the (new) root, like the reference (:103-138)."""
        self._root().insert(start, end, interval)
        return self

    def intersect(self, start, end, sort=True):
        return self._root().find(start, end)

    find = intersect

    def left(self, position, n=1, max_dist=2500):
        return self._root().before(position, n, max_dist)

    def right(self, position, n=1, max_dist=2500):
        return self._root().after(position, n, max_dist)

    def traverse(self, func):
        self._root().traverse(func)


_SMALL_CAP = 1 << 16       # hit capacity of the low-latency path's mapped buffer (csrc/itree.cu SMALL_CAP)


class _DeviceIndex:
    """Owner of one bxg_itree handle (ntrees trees)."""

    def __init__(self):
        self._h = C.c_void_p()
        self._one = None
        check(_lib.lib().bxg_itree_create(C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib._lib is not None:
            _lib._lib.bxg_itree_free(h)
            self._h = None

    def build(self, tree, start, end, ntrees):
        check(_lib.lib().bxg_itree_build(self._h, ptr(tree), ptr(start), ptr(end), len(start), ntrees, _lib.HOST))

    def find(self, qtree, qs, qe, copy=True, offsets32=False):
        """Batched find over host arrays (copies overlapped with the kernels, bxg_itree_find_host).
        copy=False returns views of the index's pinned result buffers, valid until its next find.
        offsets32=True asks for int32 CSR offsets (half the offset bytes over PCIe); falls back to int64 when the
        hits do not fit."""
        L = _lib.lib()
        total = C.c_int64()
        p_off, p_hits = C.c_void_p(), C.c_void_p()
        nq = len(qs)
        if offsets32:
            rc = L.bxg_itree_find_host32(self._h, ptr(qtree), ptr(qs), ptr(qe), nq, C.byref(p_off), C.byref(p_hits),
                                         C.byref(total))
            if rc == 0:
                n = total.value
                off = np.frombuffer((C.c_int32 * (nq + 1)).from_address(p_off.value), np.int32)
                hits = (np.frombuffer((C.c_int32 * n).from_address(p_hits.value), np.int32) if n else np.empty(0, np.int32))
                return (off.copy(), hits.copy()) if copy else (off, hits)
            if rc != _lib.ERR_MISMATCH:
                check(rc)
        check(L.bxg_itree_find_host(self._h, ptr(qtree), ptr(qs), ptr(qe), len(qs), C.byref(p_off), C.byref(p_hits),
                                    C.byref(total)))
        n = total.value
        off = np.frombuffer((C.c_int64 * (nq + 1)).from_address(p_off.value), np.int64)
        hits = (np.frombuffer((C.c_int32 * n).from_address(p_hits.value), np.int32) if n else np.empty(0, np.int32))
        if copy:
            return off.copy(), hits.copy()
        return off, hits

    def find_one(self, start, end, tree=0):
        """One query through the low-latency path (bxg_itree_find_small): -> list of item indices."""
        a = self._one
        if a is None:
            p = C.c_void_p()
            a = self._one = [p, C.byref(p), _lib.lib().bxg_itree_find1, None, None]
        n = a[2](self._h, tree, start, end, a[1])
        if n <= 0:
            if n < 0:
                check(int(n))
            return []
        addr = a[0].value
        if addr != a[3]:                   # the index's mapped result buffer: one ctypes view for all calls
            a[3], a[4] = addr, (C.c_int32 * _SMALL_CAP).from_address(addr)
        return a[4][:n] if n <= _SMALL_CAP else (C.c_int32 * n).from_address(addr)[:]

    def count(self, qtree, qs, qe):
        out = np.empty(len(qs), np.int32)
        check(_lib.lib().bxg_itree_count(self._h, ptr(qtree), ptr(qs), ptr(qe), len(qs), _lib.HOST, ptr(out), None))
        return out

    def order(self, n, ntrees):
        perm = np.empty(n, np.int32)
        toff = np.empty(ntrees + 1, np.int64)
        check(_lib.lib().bxg_itree_order(self._h, ptr(perm), ptr(toff)))
        return perm, toff

    def neighbors(self, qtree, pos, n, max_dist, direction):
        L = _lib.lib()
        total = C.c_int64()
        check(L.bxg_itree_neighbors(self._h, ptr(qtree), ptr(pos), ptr(n), ptr(max_dist), len(pos), direction,
                                    _lib.HOST, C.byref(total)))
        off = np.empty(len(pos) + 1, np.int64)
        hits = np.empty(total.value, np.int32)
        check(L.bxg_itree_fetch(self._h, ptr(off), ptr(hits)))
        return off, hits


class IntervalTree:
    """intersection.pyx:325-485 -- same methods, device-resident index."""

    # Items inserted since the last device build stay in a small host-side tail: the scalar `find` answers them by a
    # direct scan and merges them into the device hits at their in-order position, so the reference idiom of interleaved
    # insert / find (scripts/maf_drop_overlapping-style loops: `if not tree.find(s, e): tree.insert(s, e)`) costs one small
    # launch per find and one index build per TAIL_MAX inserts instead of one build per find.
    TAIL_MAX = 1024

    def __init__(self):
        self._starts, self._ends, self._values = [], [], []
        self._index = None
        self._dirty = False
        self._built = 0                    # items the device index holds (a prefix of the lists)

    # ---- position based interface --------------------------------------------------------------------------
    def insert(self, start, end, value=None):
        """Insert the interval [start,end) associated with value `value`."""
        self._starts.append(_c_int(start))
        self._ends.append(_c_int(end))
        self._values.append(value)
        self._dirty = True

    add = insert

    def insert_many(self, starts, ends, values=None):
        """Array form of insert (no reference equivalent); values defaults to the running item index."""
        s, e = as_i32(starts), as_i32(ends)
        base = len(self._starts)
        self._starts.extend(s.tolist())
        self._ends.extend(e.tolist())
        self._values.extend(values if values is not None else range(base, base + len(s)))
        self._dirty = True

    def _ensure(self, allow_tail=False):
        """The device index over all items -- or, with allow_tail, over all but a tail of at most TAIL_MAX recent ones."""
        if self._index is None:
            self._index = _DeviceIndex()
            self._dirty = True
            self._built = 0
        if self._dirty and not (allow_tail and self._built and len(self._starts) - self._built <= self.TAIL_MAX):
            self._s = np.asarray(self._starts, np.int32)
            self._e = np.asarray(self._ends, np.int32)
            self._index.build(None, self._s, self._e, 1)
            self._built = len(self._starts)
            self._dirty = False
        return self._index

    def _order_key(self, i):
        """Position of item i in the reference treap's in-order sequence (intersection.pyx:110-116): by start, items
        with end <= start first and in reverse insertion order, the others in insertion order."""
        s = self._starts[i]
        return (s, 1, i) if self._ends[i] > s else (s, 0, -i)

    def find(self, start, end):
        """Return a sorted list of all intervals overlapping [start,end)."""
        if not self._starts:
            return []
        start, end = _c_int(start), _c_int(end)
        hits = self._ensure(allow_tail=True).find_one(start, end)
        built, n = self._built, len(self._starts)
        if built < n:                      # recent inserts the device index does not hold yet
            ss, ee = self._starts, self._ends
            tail = [i for i in range(built, n) if ee[i] > start and ss[i] < end]
            if tail:
                key = self._order_key
                hits = sorted(list(hits) + tail, key=key) if hits else sorted(tail, key=key)
        v = self._values
        return [v[i] for i in hits]

    def find_batch(self, starts, ends, offsets32=False):
        """-> (offsets int64[nq+1], hits int32[total]): item indices (insertion order) per query, reference order.
        offsets32=True returns int32 offsets when the hits fit (less PCIe traffic)."""
        qs, qe = as_i32(starts), as_i32(ends)
        if not self._starts:
            return np.zeros(len(qs) + 1, np.int32 if offsets32 else np.int64), np.empty(0, np.int32)
        return self._ensure().find(None, qs, qe, offsets32=offsets32)

    def count_batch(self, starts, ends):
        """len(find(s, e)) for every query -> int32 array."""
        qs, qe = as_i32(starts), as_i32(ends)
        if not self._starts:
            return np.zeros(len(qs), np.int32)
        return self._ensure().count(None, qs, qe)

    # ---- neighbours (intersection.pyx:192-260, 408-477) ------------------------------------------------------
    def _neighbors(self, position, n, max_dist, direction):
        if not self._starts:
            return []
        pos = np.array([_c_int(position)], np.int32)
        _, hits = self._ensure().neighbors(None, pos, np.array([_c_int(n)], np.int32),
                                           np.array([_c_int(max_dist)], np.int32), direction)
        v = self._values
        return [v[i] for i in hits.tolist()]

    def before(self, position, num_intervals=1, max_dist=2500):
        return self._neighbors(position, num_intervals, max_dist, 0)

    def after(self, position, num_intervals=1, max_dist=2500):
        return self._neighbors(position, num_intervals, max_dist, 1)

    # ---- interval-like object based interface ----------------------------------------------------------------
    def insert_interval(self, interval):
        self.insert(interval.start, interval.end, interval)

    add_interval = insert_interval

    def before_interval(self, interval, num_intervals=1, max_dist=2500):
        if not self._starts:
            return []
        return self.before(interval.start, num_intervals, max_dist)

    def after_interval(self, interval, num_intervals=1, max_dist=2500):
        if not self._starts:
            return []
        return self.after(interval.end, num_intervals, max_dist)

    def upstream_of_interval(self, interval, num_intervals=1, max_dist=2500):
        if not self._starts:
            return []
        if interval.strand == -1 or interval.strand == "-":
            return self.after(interval.end, num_intervals, max_dist)
        return self.before(interval.start, num_intervals, max_dist)

    def downstream_of_interval(self, interval, num_intervals=1, max_dist=2500):
        if not self._starts:
            return []
        if interval.strand == -1 or interval.strand == "-":
            return self.before(interval.start, num_intervals, max_dist)
        return self.after(interval.end, num_intervals, max_dist)

    def traverse(self, fn):
        """call fn for each element in the tree (in-order; fn receives an IntervalNode)."""
        if not self._starts:
            return None
        perm, _ = self._ensure().order(len(self._starts), 1)
        for i in perm.tolist():
            fn(IntervalNode(self._starts[i], self._ends[i], self._values[i]))

    def order(self):
        """Item indices in in-order sequence (int32 array)."""
        if not self._starts:
            return np.empty(0, np.int32)
        return self._ensure().order(len(self._starts), 1)[0]

    def __len__(self):
        return len(self._starts)


# For backward compatibility (intersection.pyx:488)
Intersecter = IntervalTree


class IntervalForest:
    """One IntervalTree per chromosome in a single device index -- the batched form of the reference idiom
    ``ranges[chrom].add_interval(...)`` / ``ranges[chrom].find(start, end)``
    (scripts/bed_count_overlapping.py:17-33).  Trees are addressed by integer id (0..ntrees-1)."""

    def __init__(self, ntrees):
        self.ntrees = int(ntrees)
        self._index = _DeviceIndex()
        self.n = 0

    def build(self, tree_ids, starts, ends):
        t, s, e = as_i32(tree_ids), as_i32(starts), as_i32(ends)
        self.n = len(s)
        self._index.build(t, s, e, self.ntrees)
        return self

    def find_batch(self, tree_ids, starts, ends, copy=True, offsets32=False):
        """-> CSR (offsets, hits); hits are positions in the arrays passed to build().
        copy=False returns zero-copy views of pinned buffers that the next find_batch overwrites.
        offsets32=True returns int32 offsets when the hits fit (less PCIe traffic), int64 otherwise."""
        return self._index.find(as_i32(tree_ids), as_i32(starts), as_i32(ends), copy=copy, offsets32=offsets32)

    def count_batch(self, tree_ids, starts, ends):
        return self._index.count(as_i32(tree_ids), as_i32(starts), as_i32(ends))

    def order(self):
        return self._index.order(self.n, self.ntrees)

    @property
    def handle(self):
        return self._index._h
