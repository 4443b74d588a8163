"""
Build libbxb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m bx_python_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbxb200.so")
SOURCES = ["runtime.cu", "bits.cu", "itree.cu", "aggregate.cu", "scores.cu", "join.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "itree_search.cuh"), os.path.join(CSRC, "scores.cuh"),
           os.path.join(HERE, "..", "include", "bxb200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "--expt-extended-lambda",
    "-Wno-deprecated-declarations",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libbxb200.so cannot be built (there is no CPU fallback)")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    objs = []
    # tuning experiments: BXB200_NVCC_FLAGS="-DFIND_MIN_CTAS=5" rebuilds itree.cu with extra defines
    extra = os.environ.get("BXB200_NVCC_FLAGS", "").split()
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + HEADERS) or (extra and src in ("itree.cu", "bits.cu")):
            cmd = [nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            subprocess.check_call(cmd)
        objs.append(o)
    if force or _stale(LIB, objs):
        cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ldl"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
