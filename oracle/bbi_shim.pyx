# cython: language_level=3
"""
Test-infrastructure shim: exposes the reference's `cdef` method SummarizedData.accumulate_interval_value
(/root/reference/lib/bx/bbi/bbi_file.pyx:80-111) to Python so that arbitrary interval batches -- not only what the
one bigWig file of the reference's test data holds -- can be pushed through the UNMODIFIED compiled reference.
It adds no arithmetic of its own.  Built by oracle/Makefile into oracle/_ref/ next to the reference's bx.bbi modules.
"""
from bx.bbi.bbi_file cimport SummarizedData
from bx.bbi.types cimport bits32


def accumulate(SummarizedData sd, starts, ends, vals):
    cdef Py_ssize_t i
    cdef bits32 s, e
    cdef float v
    for i in range(len(starts)):
        s = starts[i]
        e = ends[i]
        v = vals[i]
        sd.accumulate_interval_value(s, e, v)
    return sd
