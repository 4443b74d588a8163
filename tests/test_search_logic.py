"""CPU: the find kernels' search / walk arithmetic (csrc/itree_search.cuh, shared by nvcc and g++) fuzzed against
std::lower_bound / a naive scan -- multi-tree segments, tiny splitter budgets (deep sampled levels), chromosome-long
items in front (max-hierarchy skipping, the probe's fallback), all three search forms (two lock-step searches, 16-ary
probe, 8-ary probe), and coordinates saturating at both ends of the int32 range (where the padding values live)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("offset,trials,extra", [(0, 1200, []), (2147400000, 400, []), (-2147400000, 400, []),
                                                 (0, 400, ["-DFUZZ_WALK_PREFETCH"])])
def test_search_and_walk_fuzz(tmp_path, offset, trials, extra):
    exe = str(tmp_path / "search_fuzz")
    subprocess.check_call(["g++", "-O2", "-std=c++17", f"-DFUZZ_OFFSET={offset}LL", *extra, "-o", exe,
                           os.path.join(ROOT, "tests", "search_fuzz.cpp")])
    out = subprocess.run([exe, str(trials)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().startswith("ok"), out.stdout[-500:]
