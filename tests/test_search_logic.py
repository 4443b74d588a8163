"""CPU: the find kernels' search / walk arithmetic (csrc/itree_search.cuh, shared by nvcc and g++) fuzzed against
std::lower_bound / a naive scan -- multi-tree segments, tiny splitter budgets (deep 16-ary levels), chromosome-long
items in front (max-hierarchy skipping)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_search_and_walk_fuzz(tmp_path):
    exe = str(tmp_path / "search_fuzz")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "search_fuzz.cpp")])
    out = subprocess.run([exe, "1200"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().startswith("ok"), out.stdout[-500:]
