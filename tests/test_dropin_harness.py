"""CPU: the drop-in harness itself (tests/dropin.py) against the compiled reference -- the reference's unit tests pass in
it and the committed script goldens (tests/golden/scripts.json) are reproduced, so a difference seen on the GPU box is the
shim's, not the harness's.  Also: bx_python_b200.shadow refuses to shadow a module that was imported first."""
import json
import os
import subprocess
import sys

import pytest

import dropin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "scripts.json")))
HARNESS = os.path.join(ROOT, "tests", "dropin.py")

needs_ref = pytest.mark.skipif(not (dropin.available() and os.path.isdir(os.path.join(dropin.REF, "bx"))),
                               reason="oracle/_ref not built")


@needs_ref
def test_reference_unit_tests_in_harness():
    r = subprocess.run([sys.executable, HARNESS, "--impl", "reference", "unittests"], capture_output=True, text=True, timeout=600)
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["rc"] == 0 and res["passed"] == GOLD["unittests_passed"] and res["failed"] == 0
    assert "oracle/_ref/bx/bitset" in res["modules"]["bx.bitset"]


@needs_ref
@pytest.mark.parametrize("k", [0, 1, 5, 7, 10, 13, 14, 15])
def test_script_goldens_reproducible(tmp_path, k):
    dropin.make_inputs(str(tmp_path), GOLD["inputs_seed"])
    name, argv, _ = dropin.SCRIPT_RUNS[k]
    r = subprocess.run([sys.executable, HARNESS, "--impl", "reference", "script", name, str(tmp_path)] + argv,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.splitlines() == GOLD["runs"][k]["stdout"]


def test_shadow_refuses_late_install():
    code = ("import sys, types; sys.path.insert(0, %r); m = types.ModuleType('bx.bitset'); sys.modules['bx.bitset'] = m\n"
            "import bx_python_b200.shadow as s\n"
            "try:\n    s.install()\nexcept RuntimeError as e:\n    print('refused')\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert "refused" in r.stdout, r.stderr[-1500:]


@needs_ref
def test_operations_goldens_reproducible():
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "operations.json")))[1]
    r = subprocess.run([sys.executable, HARNESS, "--impl", "reference", "operations", "1"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    got = json.loads(r.stdout.strip().splitlines()[-1])
    got.pop("modules")
    assert got == gold
