"""GPU: a plain C program drives libbxb200.so through include/bxb200.h (no Python shim, no torch) -- the drop-in
boundary itself.  Expected values come from the oracle and from the probes recorded in SURVEY 8(a) addendum 2."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_client(tmp_path):
    from oracle import oracle as orc
    exe = str(tmp_path / "abi_client")
    pkg = os.path.join(ROOT, "bx_python_b200")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "abi_client.c"), "-L", pkg, "-l:libbxb200.so",
                           "-Wl,-rpath," + pkg])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    a, b = orc.OracleBinnedBitSet(10000, 10), orc.OracleBinnedBitSet(10000, 10)
    a.set_ranges([0, 100, 5000], [10, 900, 2500])
    b.set_ranges([50, 6000], [500, 100])
    a.iand(b)
    exp = ["and_count %d" % a.count_range(0, 10000)]
    rs, re = a.runs()
    exp += ["run %d %d" % (x, y) for x, y in zip(rs.tolist(), re.tolist())]
    a.invert()
    exp.append("strict_counts %d %d %d" % tuple(a.count_range(s, c) for s, c in ((1500, 100), (1000, 1000), (0, 10000))))
    exp.append("next_set %d" % a.next_set(100))
    exp += ["find 0: 3 0 1 4 2", "find 1: 0", "find 2: 0 4", "find 3: 2", "total 9"]
    exp += ["small 0: 3 0 1 4 2", "small 1: 0", "small 2: 0 4", "small 3: 2", "off32 0 5 6 8 9"]
    s, e = [10, 15, 30, -5, 20], [20, 12, 30, 3, 25]
    off, items, vis = orc.join([0] * 5, s, e, [0, 0], [0, 18], [16, 40], 3)
    pairs = "".join(" %d>%d" % (q, i) for q in range(2) for i in sorted(items[off[q]:off[q + 1]].tolist()))
    got_join = lines[len(exp)].split(" visited ")
    head, got_pairs = got_join[0].split(":")
    assert head == "join %d" % len(items)
    assert sorted(got_pairs.split()) == sorted(pairs.split())              # per-left order: index order vs id order
    assert got_join[1] == "".join(str(int(v)) for v in vis)
    exp.append(lines[len(exp)])
    import numpy as np
    track = np.full(64, -1.0, np.float32)
    orc.scores_set_spans(track, 0, [2, 4, 3], [6, 5, 4], [1.0, 2.0, 3.0])
    exp.append("cells " + " ".join("%g" % v for v in track[:8]))
    o = orc.summarize([0, 5, 9], [5, 8, 20], [1.5, 2.5, -1.0], 0, 21, 3, 0.0, 0.0)
    exp.append("summary %g %g %g | %g %g %g | %g %g %g" % (*o["valid_count"], *o["sum_data"], *o["sum_squares"]))
    exp.append("multi 300 120")
    # g[0] = [0,300) of 1000, g[1] = [10,30) + [400,500) of 500; lines: (0: 0+100) (1: 0+500) (1: 390+100) (5: no such set)
    exp += ["per_line 100 120 90 0", "stats 1 100 2 210", "all_multi 300 120",
            "pieces 3 | 0:10-30 0:400-450 1:400-500", "cleared 0", "find1 2: 0 4"]
    assert lines == exp
