"""
bigWig-style summaries on the device: ``SummarizedData`` (lib/bx/bbi/bbi_file.pyx:66-111) with the
per-interval ``accumulate_interval_value`` loop replaced by one launch per batch (``bxg_summarize``), plus the
``summarize_from_full`` / ``query`` post-processing of ``BigWigFile`` / ``BBIFile`` (bigwig_file.pyx:176-185,
bbi_file.pyx:232-262) for callers that already hold the file's intervals as arrays.  Reading the bigWig container
itself (B+ tree, R-tree, zlib blocks) is file-format code and stays out of scope.
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib
from ._lib import as_i32, check, ptr


class SummarizedData:
    def __init__(self, start, end, size):
        self.start, self.end, self.size = int(start), int(end), int(size)
        self.valid_count = np.zeros(self.size, np.float64)
        self.min_val = np.zeros(self.size, np.float64)
        self.max_val = np.zeros(self.size, np.float64)
        self.sum_data = np.zeros(self.size, np.float64)
        self.sum_squares = np.zeros(self.size, np.float64)

    def accumulate_intervals(self, starts, ends, vals):
        """``for s, e, v in zip(...): self.accumulate_interval_value(s, e, v)`` (bbi_file.pyx:80-111)."""
        s, e = as_i32(starts), as_i32(ends)
        v = np.ascontiguousarray(vals, np.float32)
        check(_lib.lib().bxg_summarize(ptr(s), ptr(e), ptr(v), len(s), _lib.HOST, self.start, self.end, self.size,
                                      ptr(self.valid_count), ptr(self.min_val), ptr(self.max_val), ptr(self.sum_data),
                                      ptr(self.sum_squares)))
        return self

    def accumulate_interval_value(self, s, e, val):
        return self.accumulate_intervals([s], [e], [val])


def summarize_from_full(starts, ends, vals, start, end, summary_size):
    """BigWigFile._summarize_from_full (bigwig_file.pyx:176-185) for intervals already in memory: min / max start at
    +inf / -inf (:98-105) and valid_count is rounded in place (:182-184).  Returns None when start >= end."""
    if start >= end:
        return None
    sd = SummarizedData(start, end, summary_size)
    sd.min_val[:] = np.inf
    sd.max_val[:] = -np.inf
    sd.accumulate_intervals(starts, ends, vals)
    sd.valid_count[:] = np.round(sd.valid_count)
    return sd


def query(sd, start, end, summary_size):
    """BBIFile.query's view of a summary (bbi_file.pyx:232-262): a list of dicts with mean, max, min, coverage,
    std_dev.  Plain Python float arithmetic on the five arrays, as in the reference."""
    if sd is None:
        return None
    out = []
    for i in range(summary_size):
        sum_data, valid_count = sd.sum_data[i], sd.valid_count[i]
        with np.errstate(divide="ignore", invalid="ignore"):
            mean = sum_data / valid_count
            coverage = summary_size / (end - start) * valid_count
            variance = sd.sum_squares[i] - sum_data * sum_data / valid_count
            if valid_count > 1:
                variance /= valid_count - 1
        out.append({"mean": mean, "max": sd.max_val[i], "min": sd.min_val[i], "coverage": coverage,
                    "std_dev": math.sqrt(max(variance, 0))})
    return out
