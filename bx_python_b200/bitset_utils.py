"""
Drop-in for ``bx.bitset_utils`` (``/root/reference/lib/bx/bitset_utils.py``): lists of (start, end) as bit sets.
The next_set/next_clear loops of the reference (:34-43, :46-69, :72-85) become one run-extraction call.
"""
from __future__ import annotations

import numpy as np

from .bitset import MAX, BinnedBitSet


def list2bits(ex):
    """bitset_utils.py:27-31."""
    bits = BinnedBitSet(MAX)
    ex = list(ex)
    if ex:
        a = np.asarray(ex, np.int64).reshape(-1, 2)
        bits.set_ranges(a[:, 0], a[:, 1] - a[:, 0])
    return bits


def bits2list(bits):
    """bitset_utils.py:34-43."""
    rs, re = bits.runs()
    return list(zip(rs.tolist(), re.tolist()))


def bitset_intersect(ex1, ex2):
    bits1, bits2 = list2bits(ex1), list2bits(ex2)
    bits1.iand(bits2)
    return bits2list(bits1)


def bitset_subtract(ex1, ex2):
    bits1, bits2 = list2bits(ex1), list2bits(ex2)
    bits2.invert()
    bits1.iand(bits2)
    return bits2list(bits1)


def bitset_interval_intersect(bits, istart, iend):
    """bitset_utils.py:72-85: runs of `bits` that begin inside [istart, iend); a run already open at istart is
    reported from istart, and -- exactly like the reference loop -- the last run is NOT clipped to iend."""
    rs, re = bits.runs()
    out = []
    for s, e in zip(rs.tolist(), re.tolist()):
        if e <= istart:
            continue
        s = max(s, istart)
        if s >= iend:
            break
        out.append((s, e))
        if e >= iend:
            break
    return out


def bitset_complement(exons):
    """bitset_utils.py:46-69: gaps between the first start and the last end of `exons`."""
    exons = list(exons)
    bits = list2bits(exons)
    bits.invert()
    ex_start = min(a[0] for a in exons)
    ex_end = max(a[1] for a in exons)
    introns = []
    rs, re = bits.runs()
    for s, e in zip(rs.tolist(), re.tolist()):
        if e <= ex_start:
            continue
        s = max(s, ex_start)
        if s >= ex_end:          # the reference would emit (s, ex_end) here only if s < ex_end
            break
        e = min(e, ex_end)
        if s != e:
            introns.append((s, e))
        if e == ex_end:
            break
    return introns


def bitset_union(exons):
    return bits2list(list2bits(exons))
