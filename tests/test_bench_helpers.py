"""CPU: the host-only legs of bench.py keep working without a GPU -- the reference timings for the bitset / bed_intersect /
aggregate halves of the metric (oracle/_ref), the LPT sharding, and the NUMA helper's never-raise contract."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_cpu_reference_extras():
    from oracle import oracle as orc
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built")
    import bench
    r = bench.cpu_reference_extras()
    assert "error" not in r, r
    assert r["kind"] == "reference" and r["cores"] == 1
    assert r["bitset_and"]["gbs"] > 0 and r["count_all"]["bits_set"] > 0
    assert r["set_range"]["calls_per_s"] > 0 and r["count_range"]["calls_per_s"] > 0 and r["aggregate"]["windows_per_s"] > 0


def test_workload_shards_cover_every_chromosome_once():
    import bench
    for world in (1, 2, 8):
        _, _, shards = bench.make_workload(world, 20000, 2000)
        flat = sorted(c for s in shards for c in s)
        assert flat == list(range(24)) and len(shards) == world and all(shards)


def test_numa_helper_never_raises():
    from bx_python_b200 import _lib
    msg = _lib.bind_to_gpu_numa_node()          # no GPU here: must come back with a description, affinity untouched
    assert isinstance(msg, str) and msg
    assert isinstance(_lib._numa_nodes(), dict)
