"""
dropin.py -- harness that runs the reference's OWN unit tests and scripts, unmodified, on top of either implementation.

    python tests/dropin.py --impl b200|reference unittests              -> JSON {rc, passed, failed, modules}
    python tests/dropin.py --impl b200|reference script NAME WORKDIR ARG...   -> the script's stdout (meta on stderr)
    python tests/dropin.py --impl b200|reference operations SEED              -> JSON rows of the reference's interval operations

  --impl reference : the compiled unmodified reference (oracle/_ref/bx/*.so) + its pure-Python package (oracle/_ref/pylib)
  --impl b200      : the same pure-Python package with bx.bitset / bx.intervals.intersection (and, for the aggregate
                     script, bx.binned_array / bx.wiggle) shadowed by bx_python_b200 -- INTEGRATION.md route A
                     (bx_python_b200.shadow.install)

oracle/_ref/pylib and oracle/_ref/scripts are plain copies staged by `make -C oracle ref` (build outputs, git-ignored,
shipped to the GPU box); nothing under /root/reference is read at run time.  Always run as a subprocess: one
implementation per interpreter.
"""
import contextlib
import importlib.util
import io
import json
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "oracle", "_ref")
PYLIB = os.path.join(REF, "pylib")
SCRIPTS = os.path.join(REF, "scripts")

UNIT_TESTS = ["bx/bitset_tests.py", "bx/intervals/intersection_tests.py"]

# (script, argv, shadow the score sources too) -- file names are relative to the work directory (make_inputs)
SCRIPT_RUNS = [
    ("bed_intersect_basewise", ["a.bed", "b.bed"], False),
    ("bed_intersect", ["a.bed", "b.bed"], False),
    ("bed_intersect", ["-m", "60", "a.bed", "b.bed"], False),
    ("bed_intersect", ["-v", "a.bed", "b.bed"], False),
    ("bed_intersect", ["-b", "a.bed", "b.bed"], False),
    ("bed_count_overlapping", ["a.bed", "b.bed"], False),
    ("aggregate_scores_in_intervals", ["scores.wig", "windows.bed"], True),
    ("aggregate_scores_in_intervals", ["--mask", "mask.bed", "scores.wig", "windows.bed"], True),
    ("bed_subtract_basewise", ["a.bed", "b.bed"], False),
    ("bed_merge_overlapping", ["a.bed", "b.bed"], False),
    ("bed_complement", ["a.bed", "chrom.len"], False),
    ("bed_coverage", ["a.bed", "b.bed"], False),
    ("bed_coverage_by_interval", ["a.bed", "b.bed"], False),
    ("bed_coverage_by_interval", ["a.bed", "b.bed", "mask.bed"], False),
    ("bed_diff_basewise_summary", ["a.bed", "b.bed"], False),
    ("bed_count_by_interval", ["a.bed", "b.bed"], False),
]


def available():
    return os.path.isdir(PYLIB) and os.path.isdir(SCRIPTS)


def make_inputs(workdir, seed=0):
    """Deterministic small inputs (a few hundred lines per file, three chromosomes + one that only file a has)."""
    import numpy as np

    sys.path.insert(0, ROOT)
    from bx_python_b200 import synth
    rng = np.random.default_rng(77000 + seed)
    lens = {"chr1": 200000, "chr2": 90000, "chrX": 40000}
    os.makedirs(workdir, exist_ok=True)

    def bed(name, n, extra_chrom=None, max_len=400):
        rows = []
        names = list(lens) + ([extra_chrom] if extra_chrom else [])
        for i in range(n):
            c = names[int(rng.integers(0, len(names)))]
            L = lens.get(c, 50000)
            s = int(rng.integers(0, L - max_len))
            e = s + int(rng.integers(1, max_len))
            rows.append(f"{c}\t{s}\t{e}\tf{i}\t0\t{'+-'[int(rng.integers(0, 2))]}\n")
        with open(os.path.join(workdir, name), "w") as f:
            f.writelines(rows)
    bed("a.bed", 400, extra_chrom="chr9")
    bed("b.bed", 350)
    bed("mask.bed", 60, max_len=150)
    with open(os.path.join(workdir, "chrom.len"), "w") as f:
        for c, L in lens.items():
            f.write(f"{c}\t{L}\n")
        f.write("chr7\t12345\n")                               # a chromosome without intervals
    with open(os.path.join(workdir, "scores.wig"), "w") as f:
        f.write(synth.wiggle_text(3, n=6000, chroms=("chr1", "chr2")))
    rows = []
    for i in range(250):
        c = ("chr1", "chr2", "chrX")[int(rng.integers(0, 3))]
        s = int(rng.integers(0, 6000))
        rows.append(f"{c}\t{s}\t{s + int(rng.integers(0, 45))}\n")
    with open(os.path.join(workdir, "windows.bed"), "w") as f:
        f.writelines(rows)
    with open(os.path.join(workdir, "mask.bed"), "a") as f:    # make sure the mask also covers scored positions
        for _ in range(40):
            s = int(rng.integers(0, 6000))
            f.write(f"chr1\t{s}\t{s + int(rng.integers(1, 30))}\n")


def _load_ext(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def use_reference():
    """The compiled reference's two extension modules inside the staged pure-Python package."""
    import glob
    sys.path.insert(0, PYLIB)
    _load_ext("bx.bitset", glob.glob(os.path.join(REF, "bx", "bitset.*.so"))[0])
    _load_ext("bx.intervals.intersection", glob.glob(os.path.join(REF, "bx", "intervals", "intersection.*.so"))[0])
    import bx
    import bx.intervals
    bx.bitset = sys.modules["bx.bitset"]
    bx.intervals.intersection = sys.modules["bx.intervals.intersection"]


def use_b200(scores=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, PYLIB)
    import bx_python_b200.shadow as shadow
    shadow.install(scores=scores)
    assert shadow.installed()


def modules_report():
    names = ["bx.bitset", "bx.intervals.intersection", "bx.binned_array", "bx.wiggle", "bx.bitset_builders"]
    return {n: getattr(sys.modules.get(n), "__file__", None) for n in names if n in sys.modules}


def run_unittests():
    import pytest

    class Counter:
        passed = failed = 0
        failures = []

        def pytest_runtest_logreport(self, report):
            if report.when == "call" and report.passed:
                self.passed += 1
            elif report.failed:
                self.failed += 1
                self.failures.append(report.nodeid)
    c = Counter()
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        rc = pytest.main([os.path.join(PYLIB, t) for t in UNIT_TESTS] +
                         ["-q", "-p", "no:cacheprovider", "--rootdir", PYLIB, "-o", "python_files=*_tests.py"], plugins=[c])
    print(json.dumps({"rc": int(rc), "passed": c.passed, "failed": c.failed, "failures": c.failures,
                      "modules": modules_report(), "tail": out.getvalue()[-1500:]}))
    return 0


def run_script(name, workdir, argv):
    path = os.path.join(SCRIPTS, name + ".py")
    os.chdir(workdir)
    sys.argv = [path] + list(argv)
    runpy.run_path(path, run_name="__main__")
    print(json.dumps(modules_report()), file=sys.stderr)
    return 0


def run_operations(seed):
    """The reference's own interval operations (oracle/_ref/pylib/bx/intervals/operations/*.py over its NiceReaderWrapper /
    GenomicIntervalReader.binned_bitsets, lib/bx/intervals/io.py:190-216) on synth.ops_case(seed) -- exactly what
    tests/golden/make_golden.py:golden_operations recorded from the compiled reference."""
    sys.path.insert(0, ROOT)
    from bx_python_b200 import synth
    from bx.intervals.io import GenomicInterval, NiceReaderWrapper
    from bx.intervals.operations.base_coverage import base_coverage
    from bx.intervals.operations.complement import complement
    from bx.intervals.operations.coverage import coverage
    from bx.intervals.operations.intersect import intersect
    from bx.intervals.operations.merge import merge
    from bx.intervals.operations.subtract import subtract

    def rd(lines):
        return NiceReaderWrapper(iter([ln + "\n" for ln in lines]), chrom_col=0, start_col=1, end_col=2, strand_col=5,
                                 fix_strand=True)

    def rows(gen):
        out = []
        for r in gen:
            if isinstance(r, GenomicInterval):
                out.append([str(f).rstrip("\n") for f in r.fields])
            elif isinstance(r, list):
                out.append([str(f) for f in r])
        return out
    p, s2, s3, lens = synth.ops_case(seed)
    c = {"seed": seed}
    for pieces in (True, False):
        for mincols in (1, 40):
            key = f"pieces{int(pieces)}_min{mincols}"
            c["intersect_" + key] = rows(intersect([rd(p), rd(s2)], mincols=mincols, pieces=pieces, lens=lens))
            c["subtract_" + key] = rows(subtract([rd(p), rd(s2)], mincols=mincols, pieces=pieces, lens=lens))
    c["intersect3"] = rows(intersect([rd(p), rd(s2), rd(s3)], lens=lens))
    c["subtract3"] = rows(subtract([rd(p), rd(s2), rd(s3)], lens=lens))
    c["merge"] = rows(merge(rd(p)))
    c["complement"] = rows(complement(rd(s2), lens))
    c["coverage"] = rows(coverage([rd(p), rd(s2)]))
    c["coverage3"] = rows(coverage([rd(p), rd(s2), rd(s3)]))
    c["base_coverage"] = int(base_coverage(rd(p)))
    c["modules"] = modules_report()
    print(json.dumps(c))
    return 0


def main(args):
    assert args[0] == "--impl" and args[1] in ("b200", "reference"), __doc__
    impl, cmd, rest = args[1], args[2], args[3:]
    wants_scores = cmd == "script" and rest[0] == "aggregate_scores_in_intervals"
    if impl == "reference":
        use_reference()
    else:
        use_b200(scores=wants_scores)
    if cmd == "unittests":
        return run_unittests()
    if cmd == "script":
        return run_script(rest[0], rest[1], rest[2:])
    if cmd == "operations":
        return run_operations(int(rest[0]))
    raise SystemExit(__doc__)


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
