"""
Device-backed counterparts of ``bx.intervals.operations`` (``/root/reference/lib/bx/intervals/operations/``).

``bits_set_in_range`` / ``bits_clear_in_range`` keep the reference's generator form (``__init__.py:10-33``, including its
IndexError when the scan runs off the end of the bit set); the whole-file operations live in ``arrays`` (array in, array
out: ``intersect``, ``subtract``, ``merge``, ``complement``, ``coverage``, ``base_coverage``), ``quicksect`` / ``join`` sit on
the interval index.  The reference's own generator functions run unmodified on top of the shadowed
``bx.bitset`` (bx_python_b200.shadow; tests/test_gpu_dropin.py).
"""
BED_DEFAULT_COLS = 0, 1, 2, 5
MAX_END = 512 * 1024 * 1024


def _range_runs(bits, range_start, range_end, val):
    off, rs, re = bits.runs_in_ranges([max(int(range_start), 0)], [min(int(range_end), bits.size)], val)
    return rs.tolist(), re.tolist()


def bits_set_in_range(bits, range_start, range_end):
    """Yield start,end tuples for each span of set bits in [range_start,range_end) -- __init__.py:10-20.  One ranged
    run-extraction call instead of a next_set / next_clear round trip per span; like the reference loop it raises
    IndexError when it has to look for the next set bit at or beyond ``bits.size``."""
    range_start, range_end = int(range_start), int(range_end)
    if range_start >= bits.size or range_start < 0:
        bits.next_set(range_start)                      # the reference's first call raises here
    rs, re = _range_runs(bits, range_start, range_end, 1)
    for s, e in zip(rs, re):
        yield s, e
    # after the last span the reference calls next_set(end) -- IndexError if end == size -- and, when no set bit is
    # left in the whole bit set, next_clear(size) -- IndexError as well; otherwise it stops quietly
    end = re[-1] if re else range_start
    start = bits.next_set(end)
    bits.next_clear(start)


def bits_clear_in_range(bits, range_start, range_end):
    """Yield start,end tuples for each span of clear bits in [range_start,range_end) -- __init__.py:23-33."""
    range_start, range_end = int(range_start), int(range_end)
    if range_start >= bits.size or range_start < 0:
        bits.next_clear(range_start)
    rs, re = _range_runs(bits, range_start, range_end, 0)
    for s, e in zip(rs, re):
        yield s, e
    end = re[-1] if re else range_start
    start = bits.next_clear(end)                        # IndexError if end == size, as in the reference
    if start < range_end:
        bits.next_set(start)
