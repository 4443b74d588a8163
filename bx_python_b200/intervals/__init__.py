"""Drop-in for ``bx.intervals`` (lib/bx/intervals/__init__.py:7-14): re-exports the intersection classes."""
from .intersection import Intersecter, Interval, IntervalForest, IntervalNode, IntervalTree

__all__ = ["Intersecter", "Interval", "IntervalNode", "IntervalTree", "IntervalForest"]
