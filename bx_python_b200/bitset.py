"""
Drop-in for ``bx.bitset`` (``/root/reference/lib/bx/bitset.pyx``): ``BitSet``, ``BinnedBitSet``, ``MAX``.

Same class / method names, argument meaning and exceptions (messages copied in spirit from bitset.pyx:78-100,
177-192); every bit operation runs in CUDA kernels of libbxb200.so (csrc/bits.cu).  Scalar mutators are queued on
the host and flushed as one batched kernel before the next read, because a kernel launch per ``set_range`` call
would cost more than the reference's C call; bulk callers should use the array methods (``set_ranges``,
``count_ranges``, ``runs``, ``and_count`` ...).

``strict=True`` (default) reproduces the reference's ALL_ONE-sentinel ``count_range`` arithmetic
(src/binBits.c:155,161) bit for bit; ``strict=False`` returns the true popcount.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_i32, check, ptr

MAX_INT = 2147483647
MAX = 512 * 1024 * 1024


class _Pending:
    """Host-side queue of scalar mutations, flushed in order (consecutive ops of one kind become one launch)."""

    __slots__ = ("kind", "a", "b")

    def __init__(self):
        self.kind, self.a, self.b = None, [], []


class _DeviceBits:
    _kind = "BitSet"

    def _create(self, size, granularity):
        self._h = C.c_void_p()
        check(_lib.lib().bxg_bits_create(int(size), int(granularity), C.byref(self._h)))
        s, bs, nb = C.c_int32(), C.c_int32(), C.c_int32()
        check(_lib.lib().bxg_bits_geometry(self._h, C.byref(s), C.byref(bs), C.byref(nb)))
        self._size, self._bin_size, self._nbins = s.value, bs.value, nb.value
        self._pend = _Pending()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib._lib is not None:
            _lib._lib.bxg_bits_free(h)
            self._h = None

    # ---- queue -----------------------------------------------------------------------------------------------
    def _queue(self, kind, a, b=None):
        p = self._pend
        if p.kind is not None and p.kind != kind:
            self._flush()
        p.kind = kind
        p.a.append(a)
        if b is not None:
            p.b.append(b)
        if len(p.a) >= 1 << 20:
            self._flush()

    def _flush(self):
        p = self._pend
        if p.kind is None:
            return
        L = _lib.lib()
        a = np.asarray(p.a, np.int32)
        if p.kind == "R":
            b = np.asarray(p.b, np.int32)
            check(L.bxg_bits_set_ranges(self._h, ptr(a), ptr(b), len(a), _lib.HOST))
        else:
            check(L.bxg_bits_set_bits(self._h, ptr(a), len(a), 1 if p.kind == "S" else 0, _lib.HOST))
        _lib.sync()          # the staged host arrays must outlive the async copy
        p.kind, p.a, p.b = None, [], []

    # ---- checks (bitset.pyx:78-100 / 177-192) -----------------------------------------------------------------
    def _check_index(self, index):
        if index < 0:
            raise IndexError("BitSet index (%d) must be non-negative." % index)
        if index >= self._size:
            raise IndexError("%d is larger than the size of this BitSet (%d)." % (index, self._size))

    def _check_range_count(self, start, count):
        self._check_index(start)
        if count < 0:
            raise IndexError("Count (%d) must be non-negative." % count)
        if start + count > self._size:
            if self._kind == "BinnedBitSet":
                raise IndexError("End (%d) is larger than the size of this BinnedBitSet (%d)." % (start + count, self._size))
            raise IndexError("End %d is larger than the size of this BitSet (%d)." % (start + count, self._size))

    def _check_arrays(self, start, count):
        """Vectorised form of _check_range_count for the batched entry points: first offender raises."""
        if len(start) == 0:
            return
        s64, c64 = start.astype(np.int64), count.astype(np.int64)
        bad = (s64 < 0) | (s64 >= self._size) | (c64 < 0) | (s64 + c64 > self._size)
        if bad.any():
            k = int(np.argmax(bad))
            self._check_range_count(int(start[k]), int(count[k]))

    def _check_same(self, other):
        if not isinstance(other, type(self)):
            raise TypeError("Argument 'other' has incorrect type (expected %s, got %s)" % (type(self).__name__, type(other).__name__))
        if self._size != other._size:
            raise ValueError("BitSets must have the same size")

    # ---- shared API -------------------------------------------------------------------------------------------
    @property
    def size(self):
        return self._size

    def set(self, index):
        self._check_index(index)
        self._queue("S", int(index))

    def clear(self, index):
        self._check_index(index)
        self._queue("C", int(index))

    def set_range(self, start, count):
        start, count = int(start), int(count)
        self._check_range_count(start, count)
        self._queue("R", start, count)

    def _get(self, index):
        self._check_index(index)
        self._flush()
        a = self._scalar_io()
        a[0][0] = index
        check(a[3].bxg_bits_read(self._h, a[0], 1, a[4], _lib.HOST))
        return a[4][0]

    def _scalar_io(self):
        """Reusable ctypes cells for the one-element calls (no numpy arrays on the scalar path)."""
        a = getattr(self, "_sio", None)
        if a is None:
            a = self._sio = ((C.c_int32 * 1)(), (C.c_int32 * 1)(), (C.c_int32 * 1)(), _lib.lib(), (C.c_uint8 * 1)())
        return a

    def _count_one(self, start, count, strict):
        self._flush()
        a = self._scalar_io()
        a[0][0], a[1][0] = start, count
        check(a[3].bxg_bits_count_ranges(self._h, a[0], a[1], 1, a[2], 1 if strict else 0, _lib.HOST))
        return a[2][0]

    def __getitem__(self, index):
        return self._get(index)

    def invert(self):
        self._flush()
        check(_lib.lib().bxg_bits_not(self._h))

    def _binop(self, other, fn):
        self._check_same(other)
        self._flush()
        other._flush()
        check(fn(self._h, other._h))

    def iand(self, other):
        self._binop(other, _lib.lib().bxg_bits_and)

    def ior(self, other):
        self._binop(other, _lib.lib().bxg_bits_or)

    # ---- batched API (no reference equivalent; these are what bulk callers should use) -------------------------
    def set_ranges(self, starts, counts):
        """set_range for arrays of (start, count)."""
        s, c = as_i32(starts), as_i32(counts)
        self._check_arrays(s, c)
        self._flush()
        check(_lib.lib().bxg_bits_set_ranges(self._h, ptr(s), ptr(c), len(s), _lib.HOST))
        _lib.sync()

    def count_ranges(self, starts, counts, strict=True):
        """count_range for arrays of (start, count) -> int32 array."""
        s, c = as_i32(starts), as_i32(counts)
        self._check_arrays(s, c)
        self._flush()
        out = np.empty(len(s), np.int32)
        check(_lib.lib().bxg_bits_count_ranges(self._h, ptr(s), ptr(c), len(s), ptr(out), 1 if strict else 0, _lib.HOST))
        return out

    def get_many(self, positions):
        p = as_i32(positions)
        self._flush()
        out = np.empty(len(p), np.uint8)
        check(_lib.lib().bxg_bits_read(self._h, ptr(p), len(p), ptr(out), _lib.HOST))
        return out

    def count_all(self):
        """True popcount of the whole bitmap."""
        self._flush()
        n = C.c_int64()
        check(_lib.lib().bxg_bits_count_all(self._h, C.byref(n)))
        return n.value

    def and_count(self, other):
        """self &= other and return the number of set bits of the result (one fused kernel)."""
        self._check_same(other)
        self._flush()
        other._flush()
        n = C.c_int64()
        check(_lib.lib().bxg_bits_and_count(self._h, other._h, C.byref(n)))
        return n.value

    def runs(self):
        """All maximal runs of set bits as (starts, ends) int32 arrays -- the next_set/next_clear idiom of
        scripts/bed_intersect_basewise.py:30-38 / lib/bx/bitset_utils.py:34-43 in one call."""
        self._flush()
        n = C.c_int64()
        check(_lib.lib().bxg_bits_runs_count(self._h, C.byref(n)))
        rs, re = np.empty(n.value, np.int32), np.empty(n.value, np.int32)
        check(_lib.lib().bxg_bits_runs_fetch(self._h, ptr(rs), ptr(re), n.value))
        return rs, re

    def runs_in_ranges(self, starts, ends, val=1):
        """For every range [starts[i], ends[i]) the maximal runs of bits == val inside it, clipped to the range: the
        generators bits_set_in_range / bits_clear_in_range (lib/bx/intervals/operations/__init__.py:10-33) for a whole
        file in two launches.  -> (offsets int64[n+1], run_starts int32[], run_ends int32[]) in CSR form."""
        s, e = as_i32(starts), as_i32(ends)
        if len(s) != len(e):
            raise ValueError("starts and ends must have the same length")
        self._flush()
        off = np.zeros(len(s) + 1, np.int64)
        total = C.c_int64()
        L = _lib.lib()
        check(L.bxg_bits_runs_in_ranges(self._h, ptr(s), ptr(e), len(s), 1 if val else 0, _lib.HOST, ptr(off), C.byref(total)))
        rs, re = np.empty(total.value, np.int32), np.empty(total.value, np.int32)
        check(L.bxg_bits_runs_in_ranges_fetch(self._h, ptr(rs), ptr(re), total.value))
        return off, rs, re

    def _next(self, start, end, val):
        self._flush()
        out = C.c_int32()
        check(_lib.lib().bxg_bits_next(self._h, int(start), int(end), val, C.byref(out)))
        return out.value

    def to_words(self):
        """LSB-first uint64 words of the bitmap (test / interchange helper)."""
        self._flush()
        w = np.empty((self._size + 63) // 64, np.uint64)
        check(_lib.lib().bxg_bits_export_words(self._h, ptr(w)))
        return w

    def from_words(self, words):
        w = np.ascontiguousarray(words, np.uint64)
        if len(w) != (self._size + 63) // 64:
            raise ValueError("word count does not match size")
        self._flush()
        check(_lib.lib().bxg_bits_import_words(self._h, ptr(w)))


GROUP = 1024        # sets per multi-set launch (the library's descriptor table, csrc/bits.cu BATCH_MAX_PAIRS); larger
                    # genomes (scaffold-level assemblies with thousands of contigs) are served in groups of this many


def _batch(op, dst, src, want_counts):
    dst, src = list(dst), list(src)
    if len(dst) != len(src):
        raise ValueError("dst and src must have the same length")
    for a, b in zip(dst, src):
        a._check_same(b)
        a._flush()
        b._flush()
    n = len(dst)
    if n == 0:
        return np.zeros(0, np.int64) if want_counts else None
    counts = np.empty(n, np.int64) if want_counts else None
    for g in range(0, n, GROUP):
        m = min(GROUP, n - g)
        ha = (C.c_void_p * m)(*[a._h for a in dst[g:g + m]])
        hb = (C.c_void_p * m)(*[b._h for b in src[g:g + m]])
        part = np.empty(m, np.int64) if want_counts else None
        check(_lib.lib().bxg_bits_binop_batch(op, ha, hb, m, ptr(part)))
        if want_counts:
            counts[g:g + m] = part
    return counts


def iand_many(dst, src):
    """dst[i].iand(src[i]) for every pair in one kernel launch -- the genome-wide loop of
    scripts/bed_intersect_basewise.py:25-28 (`for chrom in bits1: bits1[chrom].iand(bits2[chrom])`)."""
    _batch(0, dst, src, False)


def ior_many(dst, src):
    _batch(1, dst, src, False)


def and_count_many(dst, src):
    """iand_many fused with a popcount of every result -> int64 array (covered bases per chromosome)."""
    return _batch(0, dst, src, True)


def _check_many(sets, w, s, c):
    """Vectorised bounds check of (which, start, count) triples: the first offending entry raises the IndexError the
    scalar call on its set would raise (bitset.pyx:177-189), before any device work."""
    if len(s) == 0:
        return
    sizes = np.asarray([b._size for b in sets], np.int64)
    inside = (w >= 0) & (w < len(sets))
    size_of = np.where(inside, sizes[np.clip(w, 0, len(sets) - 1)], np.int64(1) << 40)
    s64, c64 = s.astype(np.int64), c.astype(np.int64)
    bad = inside & ((s64 < 0) | (s64 >= size_of) | (c64 < 0) | (s64 + c64 > size_of))
    if bad.any():
        k = int(np.argmax(bad))
        sets[int(w[k])]._check_range_count(int(s[k]), int(c[k]))


def _groups(nsets, w):
    """(first set, number of sets, entry selector or None) per launch group of at most GROUP sets."""
    if nsets <= GROUP:
        yield 0, nsets, None
        return
    for g in range(0, nsets, GROUP):
        sel = np.nonzero((w >= g) & (w < g + GROUP))[0]
        if len(sel):
            yield g, min(GROUP, nsets - g), sel


def set_ranges_many(sets, which, starts, counts):
    """``sets[which[i]].set_range(starts[i], counts[i])`` for every i in one kernel launch (one per 1024 sets) -- the
    per-line loop of lib/bx/bitset_builders.py:40-53 over a whole file.  Same IndexError as the scalar call for the
    first offending entry, raised before any device work."""
    sets = list(sets)
    w, s, c = as_i32(which), as_i32(starts), as_i32(counts)
    if len(w) and (w.min() < 0 or w.max() >= len(sets)):
        raise IndexError("bit set index out of range")
    _check_many(sets, w, s, c)
    for b in sets:
        b._flush()
    if len(s) == 0:
        return
    for g, m, sel in _groups(len(sets), w):
        h = (C.c_void_p * m)(*[b._h for b in sets[g:g + m]])
        gw, gs, gc = (w, s, c) if sel is None else (np.ascontiguousarray(w[sel] - g), np.ascontiguousarray(s[sel]),
                                                    np.ascontiguousarray(c[sel]))
        check(_lib.lib().bxg_bits_set_ranges_multi(h, m, ptr(gw), ptr(gs), ptr(gc), len(gs), _lib.HOST))


def count_ranges_many(sets, which, starts, counts, strict=True):
    """``sets[which[i]].count_range(starts[i], counts[i])`` for every i in one kernel launch (one per 1024 sets) -- the
    per-line lookup of scripts/bed_intersect.py:46-53 (`bitsets[chrom].count_range(start, end - start)`).  Entries whose
    ``which`` is outside the list count 0 (the script's `fields[0] in bitsets` test)."""
    sets = list(sets)
    w, s, c = as_i32(which), as_i32(starts), as_i32(counts)
    _check_many(sets, w, s, c)                       # same IndexError as the scalar call, before any device work
    for b in sets:
        b._flush()
    out = np.zeros(len(s), np.int32)
    if len(s) == 0 or not sets:
        return out
    for g, m, sel in _groups(len(sets), w):
        h = (C.c_void_p * m)(*[b._h for b in sets[g:g + m]])
        gw, gs, gc = (w, s, c) if sel is None else (np.ascontiguousarray(w[sel] - g), np.ascontiguousarray(s[sel]),
                                                    np.ascontiguousarray(c[sel]))
        part = out if sel is None else np.empty(len(gs), np.int32)
        check(_lib.lib().bxg_bits_count_ranges_multi(h, m, ptr(gw), ptr(gs), ptr(gc), len(gs), ptr(part),
                                                     1 if strict else 0, _lib.HOST))
        if sel is not None:
            out[sel] = part
    return out


class BitSet(_DeviceBits):
    """bx.bitset.BitSet (bitset.pyx:107-173) on the device."""

    _kind = "BitSet"

    def __init__(self, bitCount):
        if bitCount > MAX_INT:
            raise ValueError("%d is larger than the maximum BitSet size of %d." % (bitCount, MAX_INT))
        self._create(bitCount, 0)

    def get(self, index):
        return self._get(index)

    def clone(self):
        other = BitSet(self._size)
        other.ior(self)
        return other

    def count_range(self, start=0, count=None):
        if count is None:
            count = self._size - start
        self._check_range_count(start, count)
        return self._count_one(start, count, True)

    def _check_range(self, start, end):
        self._check_index(start)
        if end < start:
            raise IndexError("Range end (%d) must be greater than range start(%d)." % (end, start))
        if end > self._size:
            raise IndexError("End %d is larger than the size of this BitSet (%d)." % (end, self._size))

    def next_set(self, start, end=None):
        if end is None:
            end = self._size
        self._check_range(start, end)
        return self._next(start, end, 1)

    def next_clear(self, start, end=None):
        if end is None:
            end = self._size
        self._check_range(start, end)
        return self._next(start, end, 0)

    def ixor(self, other):
        self._binop(other, _lib.lib().bxg_bits_xor)

    def __iand__(self, other):
        self.iand(other)
        return self

    def __ior__(self, other):
        self.ior(other)
        return self

    def __invert__(self):
        self.invert()
        return self


class BinnedBitSet(_DeviceBits):
    """bx.bitset.BinnedBitSet (bitset.pyx:198-241 over src/binBits.c) on the device."""

    _kind = "BinnedBitSet"

    def __init__(self, size=MAX, granularity=1024, strict=True):
        if size > MAX_INT:
            raise ValueError("%d is larger than the maximum BinnedBitSet size of %d." % (size, MAX_INT))
        self._strict = bool(strict)
        self._create(size, granularity)

    @property
    def bin_size(self):
        return self._bin_size

    def count_range(self, start, count):
        self._check_range_count(start, count)
        return self._count_one(start, count, self._strict)

    def count_ranges(self, starts, counts, strict=None):
        return super().count_ranges(starts, counts, self._strict if strict is None else strict)

    def next_set(self, start):
        self._check_index(start)
        return self._next(start, self._size, 1)

    def next_clear(self, start):
        self._check_index(start)
        return self._next(start, self._size, 0)

    def bin_states(self):
        """uint8 per bin: 0 ALL_ZERO sentinel, 1 ALL_ONE sentinel, 2 allocated (src/binBits.c:5-6)."""
        self._flush()
        out = np.empty(self._nbins, np.uint8)
        check(_lib.lib().bxg_bits_states(self._h, ptr(out)))
        return out
