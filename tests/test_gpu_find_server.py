"""
GPU tests of the lingering find server (csrc/itree.cu: k_find_server): scalar `IntervalTree.find` answered by a resident
one-warp kernel through a request line in mapped host memory.  Every case runs with the server on AND off and compares
both with the oracle's in-order treap walk (intersection.pyx:400-406 semantics, oracle.OracleIntervalTree).
"""
import ctypes as C
import gc
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from bx_python_b200 import _lib
    return _lib.lib()


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture()
def server(L):
    """Switches the server on for the test and restores the built-in default afterwards."""
    from bx_python_b200._lib import check
    check(L.bxg_set_find_server(1))
    yield lambda on: check(L.bxg_set_find_server(int(on)))
    check(L.bxg_set_find_server(-1))


def stats(L):
    a, b, c = C.c_int64(), C.c_int64(), C.c_int32()
    L.bxg_find_server_stats(C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def make_tree(rng, n, span=100000, maxlen=300):
    from bx_python_b200.intervals import IntervalTree
    s = rng.integers(-50, span, n).astype(np.int32)
    e = s + rng.integers(-3, maxlen, n).astype(np.int32)          # some empty / inverted items, as the reference allows
    t = IntervalTree()
    for i in range(n):
        t.insert(int(s[i]), int(e[i]), i)
    return t, s, e


def queries(rng, nq, span=100000):
    qs = rng.integers(-100, span + 100, nq).astype(np.int32)
    qe = qs + rng.integers(-2, 800, nq).astype(np.int32)
    return qs, qe


def test_scalar_finds_match_oracle_with_and_without_server(L, orc, server):
    rng = np.random.default_rng(5)
    t, s, e = make_tree(rng, 5000)
    qs, qe = queries(rng, 3000)
    off, hits = orc.OracleIntervalTree(s, e).find(qs, qe)
    l0, r0, _ = stats(L)
    for on in (1, 0, 1):
        server(on)
        for q in range(len(qs)):
            assert t.find(int(qs[q]), int(qe[q])) == hits[off[q]:off[q + 1]].tolist(), (on, q)
    l1, r1, alive = stats(L)
    assert r1 - r0 == 2 * len(qs)                                  # both server passes went through the mailbox
    assert 1 <= l1 - l0 < len(qs) // 4                             # ... and almost never through a launch
    assert alive == 1
    server(0)
    assert stats(L)[2] == 0


def test_alternating_trees_rebuilds_and_frees(L, orc, server):
    """One Intersecter per chromosome, lines in file order (bed_count_overlapping.py:27-33): consecutive finds address
    different indexes; inserts between finds go through the tail / rebuild; trees are dropped while the server lingers."""
    rng = np.random.default_rng(6)
    trees = [make_tree(rng, 400 + 300 * k, span=20000) for k in range(5)]
    for rnd in range(3):
        for q in range(400):
            k = int(rng.integers(0, len(trees)))
            t, s, e = trees[k]
            a = int(rng.integers(-10, 20010))
            b = a + int(rng.integers(0, 500))
            exp = [i for i in np.nonzero((s < b) & (e > a))[0].tolist()]
            assert sorted(t.find(a, b)) == exp
        # grow one tree past the host-side tail (forces a device rebuild while the server is resident) ...
        t, s, e = trees[rnd]
        ns = rng.integers(0, 20000, 1500).astype(np.int32)
        ne = ns + rng.integers(1, 200, 1500).astype(np.int32)
        for i in range(len(ns)):
            t.insert(int(ns[i]), int(ne[i]), len(s) + i)
        trees[rnd] = (t, np.concatenate([s, ns]), np.concatenate([e, ne]))
        t, s, e = trees[rnd]
        off, hits = orc.OracleIntervalTree(s, e).find(np.asarray([5000], np.int32), np.asarray([5600], np.int32))
        assert t.find(5000, 5600) == hits.tolist()
        # ... and drop another one
        trees[4 - rnd] = make_tree(rng, 700, span=20000)
        gc.collect()
        assert stats(L)[2] == 0                                     # freeing an index tells the server to leave
    t, s, e = trees[0]
    got = t.find(0, 20000)
    alive = stats(L)[2]                                             # (asked at once: it leaves after ~100 us without work)
    assert sorted(got) == np.nonzero((s < 20000) & (e > 0))[0].tolist()
    assert alive == 1


def test_idle_exit_and_relaunch(L, server):
    from bx_python_b200.intervals import IntervalTree
    t = IntervalTree()
    for i in range(100):
        t.insert(10 * i, 10 * i + 15, i)
    assert t.find(0, 12) == [0, 1]
    l0 = stats(L)[0]
    for k in range(5):
        time.sleep(0.02)                                            # far beyond the idle time-out: the kernel has left
        assert stats(L)[2] == 0
        assert t.find(100 * k, 100 * k + 12) == [10 * k - 1, 10 * k, 10 * k + 1][(k == 0):]
    assert stats(L)[0] - l0 == 5


def test_overflow_falls_back_to_the_general_path(L, server):
    from bx_python_b200.intervals import IntervalTree
    n = 70000                                                       # > SMALL_CAP hits for one query
    t = IntervalTree()
    for i in range(n):
        t.insert(0, 10, i)
    got = t.find(1, 2)
    assert len(got) == n and sorted(got) == list(range(n))
    assert t.find(11, 12) == []


def test_latency_report(L, server, capsys):
    """Not an assertion on speed -- prints the per-call time of both paths for profiles/."""
    rng = np.random.default_rng(7)
    t, s, e = make_tree(rng, 20000, span=2_000_000)
    qs, qe = queries(rng, 20000, span=2_000_000)
    out = {}
    for on in (0, 1, 0, 1):
        server(on)
        for q in range(200):
            t.find(int(qs[q]), int(qe[q]))
        t0 = time.perf_counter()
        for q in range(len(qs)):
            t.find(int(qs[q]), int(qe[q]))
        out.setdefault(on, []).append((time.perf_counter() - t0) / len(qs) * 1e6)
    with capsys.disabled():
        print("\nscalar IntervalTree.find us/call: launch per call %s, lingering server %s" %
              (["%.2f" % v for v in out[0]], ["%.2f" % v for v in out[1]]))
