"""ORACLE build recipe (test infrastructure, NOT product code).

Compiles oracle/csrc/*.cpp (our CPU restatement) into oracle/_build/liboracle.so.
The reference itself (/root/reference) cannot be compiled here: it hard-requires
deal.II >= 9.3 + PETSc + p4est + METIS + MPI (reference CMakeLists.txt:6-62), none
of which exist in this image, so there is no oracle/_ref.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    srcs = sorted(
        os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith(".cpp")
    )
    deps = srcs + [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith(".h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-std=c++17", "-o", OUT] + srcs
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
