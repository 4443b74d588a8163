"""
``bx.binned_array`` on the device: ``BinnedArray`` (lib/bx/binned_array.py:72-136) and
``FileBinnedArray`` (:177-271) over one dense float32 track in HBM.

The reference keeps ``nbins`` lazily allocated numpy bins; a bin that was never written reads as ``default``
(:89-94).  Here the whole array is one device track pre-filled with ``default`` (``bxg_scores_alloc`` /
``bxg_scores_reserve``), grown bin by bin to the highest bin written, and a host ``bool[nbins]`` remembers which bins
the reference would have allocated -- that only shows through the *type* of what ``get`` returns (the ``default``
object for an untouched bin, a ``numpy.float32`` otherwise) and through ``to_file`` (untouched bins are not stored).

Scalar ``set`` calls are queued on the host and flushed as one ``bxg_scores_set_spans`` batch (applied in order, last
write wins) before anything reads the track.  Bulk callers use ``set_many`` / ``set_spans`` / ``get_many``.  The
track plugs straight into ``bx_python_b200.aggregate`` (``ScoreTrack`` protocol: ``_h`` after ``_flush()``).

``FileBinnedArray`` uploads every stored bin when it is opened (the reference reads bins lazily behind an LRU of 32,
:214-222; its ``cache`` argument is accepted and ignored here): opening a chromosome costs one decompress + H2D of the
whole track, after which every ``get`` / ``get_range`` / aggregate launch finds it resident.

Only ``typecode="f"`` (float32, what the aggregate path stores) lives on the device; other typecodes raise
``NotImplementedError``.  The on-disk format (``to_file`` / ``FileBinnedArray``) is byte-compatible with the
reference's, version 2, ``zlib`` / ``none`` compression (``lzo`` only if the optional module is importable).
"""
from __future__ import annotations

import ctypes as C
import math
import struct
import zlib

import numpy as np

from . import _lib
from ._lib import as_i32, check, ptr

MAGIC = 0x4AB04612
VERSION = 2
MAX = 512 * 1024 * 1024

comp_types = {"none": (lambda b: b, lambda b: b), "zlib": (zlib.compress, zlib.decompress)}
try:                                                   # optional, as in the reference (:55-60)
    import lzo                                         # type: ignore
    comp_types["lzo"] = (lzo.compress, lzo.decompress)
except Exception:                                      # pragma: no cover
    pass

_QUEUE_LIMIT = 1 << 16


class _DeviceTrack:
    """Shared device plumbing: dense float32 cells [0, ncells) + which bins count as allocated."""

    def _init_track(self, bin_size, default, max_size, typecode, nbins):
        if typecode != "f":
            raise NotImplementedError("only typecode 'f' (float32) is kept on the device")
        self.max_size, self.bin_size, self.nbins = max_size, bin_size, nbins
        self.default, self.typecode = default, typecode
        self._allocated = np.zeros(self.nbins, bool)
        self._h = C.c_void_p()
        self._ncells = 0
        self._q_pos, self._q_val = [], []
        fill = np.float32(default)
        check(_lib.lib().bxg_scores_alloc(0, 0, C.c_float(float(fill)), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib._lib is not None:
            _lib._lib.bxg_scores_free(h)
            self._h = None

    # -- geometry --------------------------------------------------------------------------------------------------
    def get_bin_offset(self, index):
        return int(index // self.bin_size), int(index % self.bin_size)

    def _locate(self, key):
        """(bin, offset) with the reference's list-indexing behaviour: a negative bin wraps once, anything else
        outside ``bins`` is an IndexError (bins[bin], :90-91)."""
        b, off = self.get_bin_offset(key)
        if b < 0:
            b += self.nbins
        if not 0 <= b < self.nbins:
            raise IndexError("list index out of range")
        return b, off

    def _reserve_bins(self, last_bin):
        need = min((last_bin + 1) * self.bin_size, self.nbins * self.bin_size)
        if need > self._ncells:
            check(_lib.lib().bxg_scores_reserve(self._h, need))
            self._ncells = need

    def _flush(self):
        if self._q_pos:
            pos, val = np.asarray(self._q_pos, np.int64), np.asarray(self._q_val, np.float32)
            self._q_pos, self._q_val = [], []
            # queued point writes: keep the LAST write of every position and hand the batch over sorted, so the kernel
            # takes its sorted / disjoint path whatever order `set` was called in (otherwise a few scattered writes would
            # cost an owner array over their whole bounding range -- ADVICE r1)
            order = np.argsort(pos, kind="stable")
            ps = pos[order]
            last = np.ones(len(ps), bool)
            last[:-1] = ps[1:] != ps[:-1]
            sel = order[last]
            self._apply(pos[sel], None, val[sel])

    def _apply(self, start, end, val):
        if len(start) == 0:
            return
        start = np.asarray(start, np.int64)
        last = (start if end is None else np.asarray(end, np.int64) - 1)
        live = slice(None) if end is None else (last >= start)
        if end is not None and not np.any(live):
            return
        lo, hi = int(start[live].min()), int(last[live].max())
        if lo < 0 or hi >= self.nbins * self.bin_size:
            raise IndexError("list index out of range")
        b0, b1 = start[live] // self.bin_size, last[live] // self.bin_size
        if end is None or np.array_equal(b0, b1):
            self._allocated[np.unique(b0)] = True
        else:                                          # spans crossing bin borders: mark the whole bin run
            d = np.zeros(self.nbins + 1, np.int64)
            np.add.at(d, b0, 1)
            np.add.at(d, b1 + 1, -1)
            self._allocated |= np.cumsum(d[:-1]) > 0
        self._reserve_bins(hi // self.bin_size)
        s32 = as_i32(start)
        e32 = None if end is None else as_i32(end)
        v32 = np.ascontiguousarray(val, np.float32)
        check(_lib.lib().bxg_scores_set_spans(self._h, ptr(s32), ptr(e32), ptr(v32), len(s32), _lib.HOST))

    # -- reads -----------------------------------------------------------------------------------------------------
    def get(self, key):
        b, off = self._locate(key)
        if not self._allocated[b]:
            return self.default
        self._flush()
        return self.get_many([b * self.bin_size + off])[0]

    def get_many(self, positions):
        """float32 array of the values at ``positions`` (default where nothing was set)."""
        self._flush()
        pos = as_i32(positions)
        out = np.empty(len(pos), np.float32)
        check(_lib.lib().bxg_scores_get(self._h, ptr(pos), len(pos), ptr(out), _lib.HOST))
        return out

    def get_range(self, start, end):
        size = end - start
        assert size >= 0
        if size == 0:
            raise ValueError("need at least one array to concatenate")       # numpy.concatenate([]) in the reference
        if start < 0 or end > self.nbins * self.bin_size:
            raise IndexError("list index out of range")
        self._flush()
        out = np.empty(size, np.float32)
        check(_lib.lib().bxg_scores_get_range(self._h, int(start), int(end), ptr(out)))
        return out

    def __getitem__(self, key):
        if isinstance(key, slice):
            start, stop, stride = key.indices(self.max_size)
            assert stride == 1, "Slices with strides are not supported"
            return self.get_range(start, stop)
        return self.get(key)


class BinnedArray(_DeviceTrack):
    def __init__(self, bin_size=512 * 1024, default=np.nan, max_size=MAX, typecode="f"):
        self._init_track(bin_size, default, max_size, typecode, int(math.ceil(max_size / bin_size)))

    # -- writes ----------------------------------------------------------------------------------------------------
    def set(self, key, value):
        b, off = self._locate(key)
        self._allocated[b] = True                      # init_bin (:84-87)
        self._q_pos.append(b * self.bin_size + off)
        self._q_val.append(value)
        if len(self._q_pos) >= _QUEUE_LIMIT:
            self._flush()

    def __setitem__(self, key, value):
        return self.set(key, value)

    def set_many(self, positions, values):
        """``for p, v in zip(positions, values): self[p] = v`` as one batch (later entries win)."""
        self._flush()
        self._apply(np.asarray(positions, np.int64), None, values)

    def set_spans(self, starts, ends, values):
        """``for s, e, v in zip(...): for p in range(s, e): self[p] = v`` as one batch -- the loop of
        load_scores_wiggle (scripts/aggregate_scores_in_intervals.py:60-70) over wiggle.Reader (wiggle.py:71-85)."""
        self._flush()
        self._apply(np.asarray(starts, np.int64), np.asarray(ends, np.int64), values)

    # -- file format (:137-174) ------------------------------------------------------------------------------------
    def to_file(self, f, comp_type="zlib"):
        compress = comp_types[comp_type][0]
        self._flush()
        f.write(struct.pack(">5I", MAGIC, VERSION, self.max_size, self.bin_size, self.nbins))
        f.write(struct.pack("c", self.typecode.encode()))
        f.write(comp_type[:4].ljust(4).encode())
        f.write(np.array(self.default, ">f4").tobytes())
        index_pos = f.tell()
        f.seek(8 * self.nbins, 1)
        index = []
        for b in range(self.nbins):
            if not self._allocated[b]:
                index.append((0, 0))
                continue
            cells = self.get_range(b * self.bin_size, (b + 1) * self.bin_size)
            blob = compress(cells.astype(">f4").tobytes())
            index.append((f.tell(), len(blob)))
            f.write(blob)
        f.seek(index_pos)
        for pos, size in index:
            f.write(struct.pack(">2I", pos, size))


class FileBinnedArray(_DeviceTrack):
    """Read-only view of a binned-array file (:177-271).  Every stored bin is decompressed once and uploaded; the
    reference's LRU cache of 32 bins has no observable effect, so ``cache`` is accepted and ignored."""

    def __init__(self, f, cache=32):
        self.f = f
        M, V, max_size, bin_size, nbins = struct.unpack(">5I", f.read(20))
        assert M == MAGIC
        assert V <= VERSION, f"File is version {V} but I don't know about anything beyond {VERSION}"
        typecode = f.read(1).decode() if V >= 1 else "f"
        self.comp_type = f.read(4).strip().decode() if V >= 2 else "zlib"
        self.decompress = comp_types[self.comp_type][1]
        if typecode != "f":
            raise NotImplementedError("only typecode 'f' (float32) is kept on the device")
        default = np.frombuffer(f.read(4), ">f4")[0].astype(np.float32)
        self._init_track(bin_size, default, max_size, typecode, nbins)
        table = np.frombuffer(f.read(8 * nbins), ">u4").reshape(nbins, 2)
        self.bin_pos, self.bin_sizes = table[:, 0].tolist(), table[:, 1].tolist()
        stored = [b for b in range(nbins) if self.bin_pos[b]]
        if stored:
            self._reserve_bins(stored[-1])
        for b in stored:
            f.seek(self.bin_pos[b])
            cells = np.frombuffer(self.decompress(f.read(self.bin_sizes[b])), ">f4").astype(np.float32)
            assert len(cells) == bin_size
            self._allocated[b] = True
            check(_lib.lib().bxg_scores_write(self._h, b * bin_size, ptr(cells), bin_size, _lib.HOST))
