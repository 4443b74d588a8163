"""
Device form of the inner loop of ``scripts/aggregate_scores_in_intervals.py`` (:107-134).

``ScoreTrack`` is the dense float32 score array of one chromosome (the role of ``BinnedArray``,
lib/bx/binned_array.py:72-136: NaN where no score was set); ``aggregate`` reduces it over BED windows with the
script's exact semantics: positions are visited left to right, a score is skipped when it is 0.0 (Python
truthiness, :115), masked (:117-119) or NaN (:122); the running total is float32 (NumPy >= 2 scalar rules), and
avg = total / count in float32.  Windows with no counted base yield count 0 and NaN avg/min/max (the script prints
"nan").
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_i32, check, ptr
from .bitset import _DeviceBits


class ScoreTrack:
    def __init__(self, scores, origin=0):
        """scores[i] is the score of position origin + i (float32; NaN = unset)."""
        v = np.ascontiguousarray(scores, np.float32)
        self.n, self.origin = len(v), int(origin)
        self._h = C.c_void_p()
        check(_lib.lib().bxg_scores_create(ptr(v), len(v), self.origin, _lib.HOST, C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib._lib is not None:
            _lib._lib.bxg_scores_free(h)
            self._h = None

    def aggregate(self, starts, ends, mask=None):
        """-> dict(sum, avg, count, min, max) arrays, one entry per window [start, end)."""
        ws, we = as_i32(starts), as_i32(ends)
        nw = len(ws)
        out = dict(sum=np.empty(nw, np.float32), avg=np.empty(nw, np.float32), count=np.empty(nw, np.int32),
                   min=np.empty(nw, np.float32), max=np.empty(nw, np.float32))
        mh = None
        if mask is not None:
            if not isinstance(mask, _DeviceBits):
                raise TypeError("mask must be a bx_python_b200.bitset BitSet / BinnedBitSet")
            mask._flush()
            mh = mask._h
        check(_lib.lib().bxg_aggregate(self._h, mh, ptr(ws), ptr(we), nw, _lib.HOST, ptr(out["sum"]), ptr(out["avg"]),
                                      ptr(out["count"]), ptr(out["min"]), ptr(out["max"])))
        return out


def aggregate_genome(tracks, track_ids, starts, ends, masks=None):
    """The whole BED file in one launch: window w is reduced over ``tracks[track_ids[w]]`` (the script's
    ``scores_by_chrom[chrom]``); ``masks`` is an optional list (entries may be None) of bit sets, one per track.
    A track id outside the list behaves like a chromosome without scores (count 0, NaN columns)."""
    tracks = list(tracks)
    wt, ws, we = as_i32(track_ids), as_i32(starts), as_i32(ends)
    nw, nt = len(ws), len(tracks)
    out = dict(sum=np.empty(nw, np.float32), avg=np.empty(nw, np.float32), count=np.empty(nw, np.int32),
               min=np.empty(nw, np.float32), max=np.empty(nw, np.float32))
    for t in tracks:
        if hasattr(t, "_flush"):                       # device BinnedArray: apply queued scalar sets first
            t._flush()
    ht = (C.c_void_p * nt)(*[t._h for t in tracks])
    hm = None
    if masks is not None:
        for m in masks:
            if m is not None:
                m._flush()
        hm = (C.c_void_p * nt)(*[(m._h if m is not None else None) for m in masks])
    check(_lib.lib().bxg_aggregate_multi(ht, hm, nt, ptr(wt), ptr(ws), ptr(we), nw, _lib.HOST, ptr(out["sum"]),
                                        ptr(out["avg"]), ptr(out["count"]), ptr(out["min"]), ptr(out["max"])))
    return out


def format_line(res, w):
    """The three columns the script prints for window w (:126-134)."""
    if res["count"][w] == 0:
        return ["nan", "nan", "nan"]
    return [str(np.float32(res[k][w])) for k in ("avg", "min", "max")]


def aggregate_scores_in_intervals(scores_by_chrom, interval_lines, out_file, masks=None):
    """The main loop of scripts/aggregate_scores_in_intervals.py:105-134 as one launch: for every line
    ``chrom start stop ...`` of ``interval_lines`` write ``chrom, start, stop, avg, min, max`` (tab separated;
    ``nan`` columns when nothing was counted) to ``out_file``.  ``scores_by_chrom`` maps chromosome ->
    ``ScoreTrack`` / device ``BinnedArray`` / ``FileBinnedArray`` (``wiggle.load_scores_wiggle`` builds one);
    ``masks`` maps chromosome -> bit set (``bitset_builders.binned_bitsets_from_file``) or is None."""
    chroms, starts, stops = [], [], []
    for line in interval_lines:
        fields = line.split()
        chroms.append(fields[0])
        starts.append(int(fields[1]))
        stops.append(int(fields[2]))
    try:
        names = list(scores_by_chrom.keys())
    except (AttributeError, NotImplementedError, TypeError):
        names = None
    if names is None:                                  # a dict-like without iteration (FileBinnedArrayDir, :30-57)
        names = []
        for c in dict.fromkeys(chroms):
            try:
                scores_by_chrom[c]
                names.append(c)
            except KeyError:
                pass
    tid = {c: i for i, c in enumerate(names)}
    if not chroms:
        return
    if not names:
        res = None
    else:
        tracks = [scores_by_chrom[c] for c in names]
        mlist = None if not masks else [masks[c] if c in masks else None for c in names]
        wt = np.asarray([tid.get(c, -1) for c in chroms], np.int32)
        res = aggregate_genome(tracks, wt, starts, stops, mlist)
    for w, (c, a, b) in enumerate(zip(chroms, starts, stops)):
        cols = ["nan", "nan", "nan"] if res is None else format_line(res, w)
        print("\t".join([c, str(a), str(b)] + cols), file=out_file)
