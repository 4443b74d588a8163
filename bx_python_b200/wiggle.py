"""
``bx.wiggle`` (lib/bx/wiggle.py:16-85) plus the batched loader the aggregate path needs.

``IntervalReader`` / ``Reader`` yield exactly what the reference's generators yield (zero-based, half-open; ``bed``
mode until the first ``variableStep`` / ``fixedStep`` declaration, then the declared mode for the rest of the file).
Text parsing stays on the host -- it is not on the hot path.  What is expensive in the reference is what happens
*after* parsing: ``load_scores_wiggle`` (scripts/aggregate_scores_in_intervals.py:60-70) assigns one Python-level
``BinnedArray.__setitem__`` per covered base.  ``read_spans`` keeps the file's records as ``(start, end, value)``
arrays per chromosome, in file order, and ``load_scores_wiggle`` hands each chromosome's batch to
``BinnedArray.set_spans`` (one device launch; overlapping records resolve as the reference's sequential loop does:
the last one wins).
"""
from __future__ import annotations

import numpy as np

_SKIP = ("track", "#", "browser")


def parse_header(line):
    return dict(field.split("=") for field in line.split()[1:])


def _records(f):
    """(chrom, start, end, strand, value) per data line -- the state machine of wiggle.py:16-68."""
    chrom = pos = step = None
    span = 1
    mode = "bed"
    for line in f:
        if line.isspace() or line.startswith(_SKIP):
            continue
        if line.startswith("variableStep"):
            h = parse_header(line)
            chrom, pos, step = h["chrom"], None, None
            span = int(h["span"]) if "span" in h else 1
            mode = "variableStep"
        elif line.startswith("fixedStep"):
            h = parse_header(line)
            chrom, pos, step = h["chrom"], int(h["start"]) - 1, int(h["step"])
            span = int(h["span"]) if "span" in h else 1
            mode = "fixedStep"
        elif mode == "bed":
            fields = line.split()
            if len(fields) > 3:
                yield fields[0], int(fields[1]), int(fields[2]), (fields[5] if len(fields) > 5 else "+"), float(fields[3])
        elif mode == "variableStep":
            fields = line.split()
            p = int(fields[0]) - 1
            yield chrom, p, p + span, "+", float(fields[1])
        else:                                          # fixedStep
            yield chrom, pos, pos + span, "+", float(line.split()[0])
            pos += step


def IntervalReader(f):
    """Iterator yielding chrom, start, end, strand, value (wiggle.py:16-68)."""
    return _records(f)


class Reader:
    """Iterator yielding chrom, position, value, one per covered base (wiggle.py:71-85)."""

    def __init__(self, f):
        self.file = f

    def __iter__(self):
        for chrom, start, end, _, val in _records(self.file):
            for p in range(start, end):
                yield chrom, p, val


def read_spans(f):
    """-> {chrom: (starts int64[], ends int64[], values float32[])} in file order (dict order = first appearance)."""
    acc = {}
    for chrom, start, end, _, val in _records(f):
        rec = acc.get(chrom)
        if rec is None:
            rec = acc[chrom] = ([], [], [])
        rec[0].append(start)
        rec[1].append(end)
        rec[2].append(val)
    return {c: (np.asarray(s, np.int64), np.asarray(e, np.int64), np.asarray(v, np.float64).astype(np.float32))
            for c, (s, e, v) in acc.items()}


def load_scores_wiggle(f):
    """Read a wiggle file (name or open text file) into a dict of device ``BinnedArray`` keyed by chromosome --
    scripts/aggregate_scores_in_intervals.py:60-70.  A chromosome whose records are all empty still gets its (empty)
    array only if the reference would have created it, i.e. if at least one base was assigned."""
    from .binned_array import BinnedArray
    if isinstance(f, (str, bytes)):
        with open(f) as fh:
            spans = read_spans(fh)
    else:
        spans = read_spans(f)
    out = {}
    for chrom, (s, e, v) in spans.items():
        if not np.any(e > s):
            continue
        ba = out[chrom] = BinnedArray()
        ba.set_spans(s, e, v)
    return out
