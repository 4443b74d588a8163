"""
oracle.py -- ctypes access to the CPU restatement (oracle/bx_oracle.c) and, when it was built, to the compiled
unmodified reference (oracle/_ref).  TEST INFRASTRUCTURE ONLY: imported by tests/, by bench.py's cpu_baseline /
--impl reference legs and by __graft_entry__.smoke(); never by bx_python_b200/.

Parity status: pinned -- see the header of bx_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build_port() -> None:
    """Compile bx_oracle.c -> liboracle.so (gcc only)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


def build_ref(reference="/root/reference") -> bool:
    """Compile the unmodified reference into oracle/_ref (only possible where /root/reference exists)."""
    if not os.path.isdir(reference):
        return False
    subprocess.check_call(["make", "-s", "-C", HERE, "ref", f"REF={reference}"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return True


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "bx_oracle.c")):
        build_port()
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    sig = {
        "orc_itree_order": (C.c_int, [_i32p, _i32p, i64, _i32p]),
        "orc_itree_build": (vp, [_i32p, _i32p, i64]),
        "orc_itree_free": (None, [vp]),
        "orc_itree_find": (C.c_int, [vp, _i32p, _i32p, i64, _i64p, C.c_void_p]),
        "orc_itree_before": (i64, [vp, i32, i32, i32, _i32p, i64]),
        "orc_itree_after": (i64, [vp, i32, i32, i32, _i32p, i64]),
        "orc_scores_set_spans": (C.c_int, [_f32p, i64, i64, _i32p, C.c_void_p, _f32p, i64]),
        "orc_summarize": (None, [_i32p, _i32p, _f32p, i64, C.c_uint32, C.c_uint32, i32, _f64p, _f64p, _f64p, _f64p, _f64p]),
        "orc_join": (i64, [_i32p, _i32p, _i32p, i64, _i32p, _i32p, _i32p, i64, i32, _i64p, C.c_void_p, _u8p]),
        "orc_bb_geometry": (None, [i32, i32, C.POINTER(i32), C.POINTER(i32)]),
        "orc_bb_new": (vp, [i32, i32]),
        "orc_bb_free": (None, [vp]),
        "orc_bb_size": (i32, [vp]),
        "orc_bb_bin_size": (i32, [vp]),
        "orc_bb_nbins": (i32, [vp]),
        "orc_bb_state": (i32, [vp, i32]),
        "orc_bb_get": (C.c_int, [vp, i32]),
        "orc_bb_set": (None, [vp, i32]),
        "orc_bb_clear": (None, [vp, i32]),
        "orc_bb_set_range": (None, [vp, i32, i32]),
        "orc_bb_count_range": (i32, [vp, i32, i32]),
        "orc_bb_next": (i32, [vp, i32, C.c_int]),
        "orc_bb_and": (None, [vp, vp]),
        "orc_bb_or": (None, [vp, vp]),
        "orc_bb_not": (None, [vp]),
        "orc_bb_set_ranges": (None, [vp, _i32p, _i32p, i64]),
        "orc_bb_count_ranges": (None, [vp, _i32p, _i32p, i64, _i32p]),
        "orc_bb_read": (None, [vp, _i32p, i64, _u8p]),
        "orc_bb_export_words": (None, [vp, _u64p]),
        "orc_bb_runs": (i64, [vp, C.c_void_p, C.c_void_p, i64]),
        "orc_bits_new": (vp, [i32]),
        "orc_bits_free": (None, [vp]),
        "orc_bits_get": (C.c_int, [vp, i32]),
        "orc_bits_set": (None, [vp, i32]),
        "orc_bits_clear": (None, [vp, i32]),
        "orc_bits_set_range": (None, [vp, i32, i32]),
        "orc_bits_count_range": (i32, [vp, i32, i32]),
        "orc_bits_next": (i32, [vp, i32, i32, C.c_int]),
        "orc_bits_binop": (None, [vp, vp, C.c_int]),
        "orc_bits_not": (None, [vp]),
        "orc_aggregate": (None, [_f32p, i64, C.c_void_p, _i32p, _i32p, i64, _f32p, _f32p, _i32p, _f32p, _f32p]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def _a32(x):
    return np.ascontiguousarray(x, dtype=np.int32)


# ---------------------------------------------------------------------------------------------------------------
# Interval index
# ---------------------------------------------------------------------------------------------------------------
class OracleIntervalTree:
    """Array-in / array-out restatement of bx.intervals.intersection.IntervalTree (find + neighbours)."""

    def __init__(self, start, end):
        self.start, self.end = _a32(start), _a32(end)
        self.n = len(self.start)
        self._h = lib().orc_itree_build(self.start, self.end, self.n)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_itree_free(self._h)
            self._h = None

    def order(self):
        perm = np.empty(self.n, np.int32)
        lib().orc_itree_order(self.start, self.end, self.n, perm)
        return perm

    def find(self, qs, qe):
        """-> (offsets int64[nq+1], hits int32[total]) ; hits are insertion indices in reference order."""
        qs, qe = _a32(qs), _a32(qe)
        nq = len(qs)
        off = np.empty(nq + 1, np.int64)
        lib().orc_itree_find(self._h, qs, qe, nq, off, None)
        hits = np.empty(int(off[-1]), np.int32)
        lib().orc_itree_find(self._h, qs, qe, nq, off, hits.ctypes.data_as(C.c_void_p))
        return off, hits

    def before(self, position, n=1, max_dist=2500):
        out = np.empty(max(self.n, 1), np.int32)
        m = lib().orc_itree_before(self._h, position, n, max_dist, out, len(out))
        return out[:m].copy()

    def after(self, position, n=1, max_dist=2500):
        out = np.empty(max(self.n, 1), np.int32)
        m = lib().orc_itree_after(self._h, position, n, max_dist, out, len(out))
        return out[:m].copy()


# ---------------------------------------------------------------------------------------------------------------
# BinnedBitSet / BitSet
# ---------------------------------------------------------------------------------------------------------------
class OracleBinnedBitSet:
    def __init__(self, size=512 * 1024 * 1024, granularity=1024):
        self._h = lib().orc_bb_new(size, granularity)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_bb_free(self._h)
            self._h = None

    size = property(lambda s: lib().orc_bb_size(s._h))
    bin_size = property(lambda s: lib().orc_bb_bin_size(s._h))
    nbins = property(lambda s: lib().orc_bb_nbins(s._h))

    def states(self):
        return np.array([lib().orc_bb_state(self._h, i) for i in range(self.nbins)], np.uint8)

    def __getitem__(self, i): return lib().orc_bb_get(self._h, i)
    def set(self, i): lib().orc_bb_set(self._h, i)
    def clear(self, i): lib().orc_bb_clear(self._h, i)
    def set_range(self, s, c): lib().orc_bb_set_range(self._h, s, c)
    def count_range(self, s, c): return lib().orc_bb_count_range(self._h, s, c)
    def next_set(self, s): return lib().orc_bb_next(self._h, s, 1)
    def next_clear(self, s): return lib().orc_bb_next(self._h, s, 0)
    def iand(self, o): lib().orc_bb_and(self._h, o._h)
    def ior(self, o): lib().orc_bb_or(self._h, o._h)
    def invert(self): lib().orc_bb_not(self._h)

    def set_ranges(self, s, c):
        s, c = _a32(s), _a32(c)
        lib().orc_bb_set_ranges(self._h, s, c, len(s))

    def count_ranges(self, s, c):
        s, c = _a32(s), _a32(c)
        out = np.empty(len(s), np.int32)
        lib().orc_bb_count_ranges(self._h, s, c, len(s), out)
        return out

    def read(self, pos):
        pos = _a32(pos)
        out = np.empty(len(pos), np.uint8)
        lib().orc_bb_read(self._h, pos, len(pos), out)
        return out

    def words(self):
        w = np.empty((self.size + 63) // 64, np.uint64)
        lib().orc_bb_export_words(self._h, w)
        return w

    def runs(self):
        n = lib().orc_bb_runs(self._h, None, None, 0)
        rs, re = np.empty(n, np.int32), np.empty(n, np.int32)
        lib().orc_bb_runs(self._h, rs.ctypes.data_as(C.c_void_p), re.ctypes.data_as(C.c_void_p), n)
        return rs, re


class OracleBitSet:
    def __init__(self, n):
        self.n = n
        self._h = lib().orc_bits_new(n)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_bits_free(self._h)
            self._h = None

    size = property(lambda s: s.n)
    def __getitem__(self, i): return lib().orc_bits_get(self._h, i)
    get = __getitem__
    def set(self, i): lib().orc_bits_set(self._h, i)
    def clear(self, i): lib().orc_bits_clear(self._h, i)
    def set_range(self, s, c): lib().orc_bits_set_range(self._h, s, c)
    def count_range(self, s=0, c=None): return lib().orc_bits_count_range(self._h, s, self.n - s if c is None else c)
    def next_set(self, s, e=None): return lib().orc_bits_next(self._h, s, self.n if e is None else e, 1)
    def next_clear(self, s, e=None): return lib().orc_bits_next(self._h, s, self.n if e is None else e, 0)
    def iand(self, o): lib().orc_bits_binop(self._h, o._h, 0)
    def ior(self, o): lib().orc_bits_binop(self._h, o._h, 1)
    def ixor(self, o): lib().orc_bits_binop(self._h, o._h, 2)
    def invert(self): lib().orc_bits_not(self._h)


# ---------------------------------------------------------------------------------------------------------------
# aggregate
# ---------------------------------------------------------------------------------------------------------------
def aggregate(scores, ws, we, mask_words=None):
    """-> dict(sum, avg, count, min, max) per window; see orc_aggregate."""
    scores = np.ascontiguousarray(scores, np.float32)
    ws, we = _a32(ws), _a32(we)
    nw = len(ws)
    out = dict(sum=np.empty(nw, np.float32), avg=np.empty(nw, np.float32), count=np.empty(nw, np.int32),
               min=np.empty(nw, np.float32), max=np.empty(nw, np.float32))
    mp = None
    if mask_words is not None:
        mask_words = np.ascontiguousarray(mask_words, np.uint64)
        mp = mask_words.ctypes.data_as(C.c_void_p)
    lib().orc_aggregate(scores, len(scores), mp, ws, we, nw, out["sum"], out["avg"], out["count"], out["min"], out["max"])
    return out


# ---------------------------------------------------------------------------------------------------------------
# score sources, bigWig summary, join
# ---------------------------------------------------------------------------------------------------------------
def scores_set_spans(track, origin, start, end, val):
    """In place: the sequential per-base assignment loop (last record wins); see orc_scores_set_spans."""
    assert track.dtype == np.float32 and track.flags.c_contiguous
    start, val = _a32(start), np.ascontiguousarray(val, np.float32)
    ep = None
    if end is not None:
        end = _a32(end)
        ep = end.ctypes.data_as(C.c_void_p)
    rc = lib().orc_scores_set_spans(track, len(track), origin, start, ep, val, len(start))
    if rc != 0:
        raise IndexError("span outside the track")
    return track


def summarize(start, end, val, rstart, rend, size, init_min=np.inf, init_max=-np.inf):
    """-> dict(valid_count, min_val, max_val, sum_data, sum_squares) float64[size]; see orc_summarize."""
    start, end, val = _a32(start), _a32(end), np.ascontiguousarray(val, np.float32)
    out = dict(valid_count=np.zeros(size), min_val=np.full(size, float(init_min)), max_val=np.full(size, float(init_max)),
               sum_data=np.zeros(size), sum_squares=np.zeros(size))
    lib().orc_summarize(start, end, val, len(start), rstart, rend, size, out["valid_count"], out["min_val"],
                        out["max_val"], out["sum_data"], out["sum_squares"])
    return out


def join(itree, istart, iend, qtree, qs, qe, mincols=1):
    """-> (pair_offsets, pair_items, visited); see orc_join."""
    itree, istart, iend = _a32(itree), _a32(istart), _a32(iend)
    qtree, qs, qe = _a32(qtree), _a32(qs), _a32(qe)
    nq, n = len(qs), len(istart)
    off = np.zeros(nq + 1, np.int64)
    vis = np.zeros(max(n, 1), np.uint8)
    total = lib().orc_join(itree, istart, iend, n, qtree, qs, qe, nq, mincols, off, None, vis)
    items = np.empty(max(total, 1), np.int32)
    vis[:] = 0
    lib().orc_join(itree, istart, iend, n, qtree, qs, qe, nq, mincols, off, items.ctypes.data_as(C.c_void_p), vis)
    return off, items[:total], vis[:n].astype(bool)


# ---------------------------------------------------------------------------------------------------------------
# Compiled reference (oracle/_ref), when present
# ---------------------------------------------------------------------------------------------------------------
def ref_available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "bx"))


def ref_modules():
    """-> (bx.bitset, bx.intervals.intersection) of the compiled UNMODIFIED reference, or raises ImportError."""
    if not ref_available():
        raise ImportError("oracle/_ref not built (make -C oracle ref needs /root/reference)")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib
    bitset = importlib.import_module("bx.bitset")
    inter = importlib.import_module("bx.intervals.intersection")
    return bitset, inter
