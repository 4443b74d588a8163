"""
Route A of INTEGRATION.md: make an UNMODIFIED bx-python installation use the B200 implementations.

    import bx_python_b200.shadow
    bx_python_b200.shadow.install()          # before anything imports bx.bitset / bx.intervals

After ``install()`` the names ``bx.bitset`` and ``bx.intervals.intersection`` resolve to
``bx_python_b200.bitset`` / ``bx_python_b200.intervals.intersection`` -- both as ``sys.modules`` entries (for
``from bx.bitset import BinnedBitSet``) and as attributes of their parent packages (for ``import bx.bitset`` followed by
``bx.bitset.BitSet(...)``, which is how lib/bx/bitset_tests.py:7,11 uses it).  Everything else of the ``bx`` package (the
pure-Python callers lib/bx/bitset_builders.py, bitset_utils.py, intervals/io.py, intervals/operations/*, the scripts)
keeps running unmodified on top of them.  ``scores=True`` also shadows the score sources of
scripts/aggregate_scores_in_intervals.py (``bx.binned_array``, ``bx.wiggle``), ``operations=True`` the quicksect / join
pair.

tests/test_gpu_dropin.py runs the reference's own unit tests for the path and the scripts BASELINE.json names through
this shadow.
"""
from __future__ import annotations

import importlib
import sys

_SHADOWS = {
    "bx.bitset": "bx_python_b200.bitset",
    "bx.intervals.intersection": "bx_python_b200.intervals.intersection",
}
_SCORES = {
    "bx.binned_array": "bx_python_b200.binned_array",
    "bx.wiggle": "bx_python_b200.wiggle",
}
_OPERATIONS = {
    "bx.intervals.operations.quicksect": "bx_python_b200.intervals.operations.quicksect",
    "bx.intervals.operations.join": "bx_python_b200.intervals.operations.join",
}


def install(scores: bool = False, operations: bool = False) -> dict:
    """Shadow the compiled modules of an importable ``bx`` package.  Returns {shadowed name: module}.
    Raises ImportError when there is no ``bx`` package to shadow, RuntimeError when the real extension module was
    imported first (objects created from it would not interoperate)."""
    table = dict(_SHADOWS)
    if scores:
        table.update(_SCORES)
    if operations:
        table.update(_OPERATIONS)
    done = {}
    # sys.modules first: importing the parent packages below runs bx/intervals/__init__.py, which itself does
    # `from bx.intervals.intersection import ...` (lib/bx/intervals/__init__.py:7-14)
    for name, ours in table.items():
        mod = importlib.import_module(ours)
        prev = sys.modules.get(name)
        if prev is not None and prev is not mod:
            raise RuntimeError(f"{name} was already imported from {getattr(prev, '__file__', '?')}; call "
                               "bx_python_b200.shadow.install() before anything imports it")
        sys.modules[name] = mod
        done[name] = mod
    try:
        for name, mod in done.items():
            parent, _, leaf = name.rpartition(".")
            setattr(importlib.import_module(parent), leaf, mod)
    except ImportError:
        for name in done:                              # no `bx` package to shadow: leave sys.modules as it was
            sys.modules.pop(name, None)
        raise
    return done


def installed() -> bool:
    return all(sys.modules.get(name) is sys.modules.get(ours) and name in sys.modules for name, ours in _SHADOWS.items())
