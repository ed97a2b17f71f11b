"""Build recipe for libsandcrate.so (the C-ABI library of include/sandcrate.h): nvcc, sm_100a only, in-tree.

    python -m sand_crate_b200.build [--force]

-fmad=false: the fp64 parity mode must not contract a*b+c into FMA because NumPy never does (SURVEY.md
section 8(c)); the fp32 production kernels spell out fmaf() where fusion is wanted.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# SC_LIB: developer override (a variant build to A/B against); the product path is the in-tree default
LIB = os.environ.get("SC_LIB") or os.path.join(HERE, "libsandcrate.so")
SOURCES = ["sc_api.cu"]
HEADERS = ["sc_common.cuh", "sc_pair.cuh", "sc_tile.cuh", "sc_sort.cuh", "sc_dist.cuh", os.path.join("..", "..", "include", "sandcrate.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "--shared", "-Xptxas", "-v", "-ldl",
]


def needs_build() -> bool:
    if os.environ.get("SC_LIB") and os.path.exists(LIB):
        return False
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    out = out or LIB
    cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("nvcc failed building libsandcrate.so")
    if out == LIB:
        with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
            f.write(res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
