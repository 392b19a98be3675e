"""Build recipe for the product library (in-tree, so the .so travels with the repo to the GPU box).

nvcc cross-compiles every translation unit under openifem_b200/csrc for sm_100a
only and links them into openifem_b200/lib/libopenifem_b200.so, the C-ABI shared
library declared in include/openifem_b200.h.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "lib", "libopenifem_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
    "-Xcompiler", "-fPIC,-fopenmp,-O3", "-Xptxas", "-v",
]


def _needs(out, deps):
    return not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    srcs = sorted(f for f in os.listdir(SRC) if f.endswith((".cu", ".cpp")))
    hdrs = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "openifem_b200.h"))
    jobs = []
    objs = []
    for f in srcs:
        o = os.path.join(OBJ, f + ".o")
        objs.append(o)
        if force or _needs(o, [os.path.join(SRC, f)] + hdrs):
            cmd = [NVCC] + FLAGS + (["-x", "cu"] if f.endswith(".cu") else []) + ["-c", os.path.join(SRC, f), "-o", o]
            jobs.append((f, cmd))

    def run(job):
        f, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {f}:\n{r.stdout}\n{r.stderr}")
        with open(os.path.join(OBJ, f + ".ptxas.log"), "w") as fh:
            fh.write(r.stderr)
        if verbose:
            print(r.stderr)
        return f

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, jobs))
    if jobs or force or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-Xcompiler", "-fopenmp", "-ldl", "-lgomp"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
