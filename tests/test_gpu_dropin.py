"""
The drop-in claim, executed: INTEGRATION.md route A (bx_python_b200.shadow.install) under the reference's OWN, UNMODIFIED
unit tests for this path (lib/bx/bitset_tests.py:10-119, lib/bx/intervals/intersection_tests.py:17-201) and under the
scripts BASELINE.json names (bed_intersect_basewise.py, bed_intersect.py, bed_count_overlapping.py,
aggregate_scores_in_intervals.py) plus the other bitset-builder callers -- stdout compared line by line with the same
scripts run on the compiled reference (tests/golden/scripts.json, tests/golden/make_golden.py:golden_scripts).
The reference files are staged copies under oracle/_ref/ (`make -C oracle ref`); each run is its own interpreter.
"""
import json
import os
import subprocess
import sys

import pytest

import dropin

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "scripts.json")))
HARNESS = os.path.join(ROOT, "tests", "dropin.py")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    if not dropin.available():
        pytest.fail("oracle/_ref/pylib is missing: `make -C oracle ref` stages it (it is shipped to the GPU box by gpurun)")
    d = tmp_path_factory.mktemp("dropin")
    dropin.make_inputs(str(d), GOLD["inputs_seed"])
    return str(d)


def test_reference_unit_tests_pass_on_the_shadowed_modules():
    r = subprocess.run([sys.executable, HARNESS, "--impl", "b200", "unittests"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["failed"] == 0 and res["rc"] == 0, (res["failures"], res["tail"])
    assert res["passed"] == GOLD["unittests_passed"] == 31
    assert res["modules"]["bx.bitset"].endswith("bx_python_b200/bitset.py")
    assert res["modules"]["bx.intervals.intersection"].endswith("bx_python_b200/intervals/intersection.py")


@pytest.mark.parametrize("k", range(len(dropin.SCRIPT_RUNS)), ids=[f"{n}:{'_'.join(a)}" for n, a, _ in dropin.SCRIPT_RUNS])
def test_reference_script_output_identical(workdir, k):
    name, argv, scores = dropin.SCRIPT_RUNS[k]
    g = GOLD["runs"][k]
    assert (g["script"], g["argv"]) == (name, argv)
    r = subprocess.run([sys.executable, HARNESS, "--impl", "b200", "script", name, workdir] + argv,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    mods = json.loads(r.stderr.strip().splitlines()[-1])
    assert mods["bx.bitset"].endswith("bx_python_b200/bitset.py"), mods                  # it really ran on the shim ...
    if name not in ("bed_count_overlapping", "bed_count_by_interval"):                    # (those only use bx.intervals)
        assert mods["bx.bitset_builders"].endswith("oracle/_ref/pylib/bx/bitset_builders.py"), mods
    if scores:
        assert mods["bx.binned_array"].endswith("bx_python_b200/binned_array.py"), mods
    got = r.stdout.splitlines()
    assert len(got) == len(g["stdout"]), (len(got), len(g["stdout"]))
    assert got == g["stdout"]                                                             # ... and printed the same lines


@pytest.mark.parametrize("seed", [0, 3])
def test_reference_interval_operations_run_on_the_shim(seed):
    """lib/bx/intervals/operations/{intersect,subtract,merge,complement,coverage,base_coverage}.py and
    GenomicIntervalReader.binned_bitsets (lib/bx/intervals/io.py:190-216), unmodified, on the shadowed bx.bitset: row for row
    what they produce on the compiled reference (tests/golden/operations.json)."""
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "operations.json")))[seed]
    r = subprocess.run([sys.executable, HARNESS, "--impl", "b200", "operations", str(seed)], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    got = json.loads(r.stdout.strip().splitlines()[-1])
    assert got.pop("modules")["bx.bitset"].endswith("bx_python_b200/bitset.py")
    assert sorted(got) == sorted(gold)
    for k in gold:
        assert got[k] == gold[k], k
