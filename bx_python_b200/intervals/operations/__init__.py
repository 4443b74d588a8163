"""Device-backed counterparts of ``bx.intervals.operations`` that sit on the interval index."""
