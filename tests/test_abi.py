"""CPU: libbxb200.so loads, exports every symbol include/bxb200.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bxb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bxg_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from bx_python_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()


def test_header_declares_something():
    syms = declared_symbols()
    assert len(syms) >= 50 and "bxg_itree_find" in syms and "bxg_bits_and" in syms


def test_every_declared_symbol_is_exported(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_header():
    from bx_python_b200 import _lib
    assert set(_lib.SIGNATURES) == set(declared_symbols())


def test_no_cpu_fallback(lib):
    """Without a CUDA device every compute entry point must fail with a message, never fall back."""
    n = C.c_int()
    rc = lib.bxg_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is visible")
    from bx_python_b200 import bitset
    from bx_python_b200.intervals import IntervalTree
    with pytest.raises(RuntimeError):
        bitset.BinnedBitSet(100)
    t = IntervalTree()
    t.insert(1, 5, "a")           # host-side queueing works without a device
    with pytest.raises(RuntimeError):
        t.find(0, 10)


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (the checker) anywhere."""
    pkg = os.path.join(ROOT, "bx_python_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "oracle" not in txt.replace("bx_oracle", "").lower() or f == "synth.py", f


def test_host_side_checks_match_reference_messages():
    """Argument validation happens on the host, before any device call (bitset.pyx:177-192 messages)."""
    from bx_python_b200 import bitset
    with pytest.raises(ValueError, match="larger than the maximum BinnedBitSet size"):
        bitset.BinnedBitSet(4000000000)
    with pytest.raises(ValueError, match="larger than the maximum BitSet size"):
        bitset.BitSet(4000000000)
    from bx_python_b200.intervals import Interval, IntervalTree
    with pytest.raises(AssertionError):
        Interval(5, 3)
    with pytest.raises(OverflowError):
        IntervalTree().insert(0, 2**31)
    assert repr(Interval(3, 7)) == "Interval(3, 7)"
    assert repr(Interval(3, 7, value=5)) == "Interval(3, 7, value=5)"
    assert IntervalTree().find(100, 300) == []
    assert IntervalTree().traverse(lambda x: None) is None


def header_prototypes():
    """{name: [parameter type class, ...]} parsed from include/bxb200.h (classes: ptr, i32, i64, u32, f32, f64)."""
    src = open(os.path.join(ROOT, "include", "bxb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for name, params in re.findall(r"\b(bxg_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src, flags=re.S):
        classes = []
        for p in (q.strip() for q in params.split(",")):
            if p in ("", "void"):
                continue
            if "*" in p or "[" in p:
                classes.append("ptr")
            elif "float" in p:
                classes.append("f32")
            elif "double" in p:
                classes.append("f64")
            elif "int64_t" in p:
                classes.append("i64")
            elif "uint32_t" in p:
                classes.append("u32")
            elif "int32_t" in p:
                classes.append("i32")
            else:
                assert re.match(r"(const )?int\b", p), (name, p)
                classes.append("i32")                  # int is 32 bits on every platform this library targets
        out[name] = classes
    return out


def test_python_binding_matches_header_prototypes():
    """Guards the hand-written ctypes table against ABI drift: same parameter count and the same kind of type
    (pointer / 32-bit / 64-bit / float / double) in every position as the C prototype."""
    from bx_python_b200 import _lib
    kind = {C.c_int32: "i32", C.c_int: "i32", C.c_int64: "i64", C.c_uint32: "u32", C.c_float: "f32", C.c_double: "f64"}

    def classify(t):
        if t in kind:
            return kind[t]
        assert t in (C.c_void_p, C.c_char_p) or issubclass(t, C._Pointer), t
        return "ptr"
    protos = header_prototypes()
    assert set(protos) == set(_lib.SIGNATURES)
    for name, classes in protos.items():
        got = [classify(t) for t in _lib.SIGNATURES[name]]
        assert got == classes, (name, got, classes)
