"""Probe: setup time, Newton / FGMRES / inner-solver iteration counts and section times of the 3-D cavity
at a given number of cells per direction (development aid, not a test)."""
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
from util import cavity_prm

import openifem_b200 as ifem

n = int(sys.argv[1])
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
a_inv_rel = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-3
fp32 = int(sys.argv[4]) if len(sys.argv) > 4 else 0
max_it = int(sys.argv[5]) if len(sys.argv) > 5 else 2000
ifem.init(0)
t0 = time.time()
tria = ifem.Triangulation(3)
ifem.GridGenerator.subdivided_hyper_rectangle(tria, (n, n, n), (0, 0, 0), (1, 1, 1), True)
params = ifem.Parameters.AllParameters(text=cavity_prm(3))
s = ifem.Fluid.MPI.InsIM(tria, params)
s.setup()
print(f"n={n} setup {time.time()-t0:.1f}s sizes={s.sizes()}", flush=True)
s.set_control(a_inv_rel=a_inv_rel, a_inv_fp32=fp32, a_inv_max_it=max_it)
s.set_verbose(True)
for k in range(steps):
    t0 = time.time()
    s.run_one_step(k == 0)
    print(f"step {k} wall {time.time()-t0:.2f}s", flush=True)
for sec in ["Assemble system", "Solve linear system", "CG for Mp", "CG for Sm", "A_inv"]:
    print(sec, f"{s.timer_ms(sec):.1f} ms")
ms, b = s.bench_spmv_uu(10)
print(f"spmv uu: {ms:.3f} ms, {b/ms/1e6:.1f} GB/s")
ms, b = s.bench_vmult(10)
print(f"block vmult: {ms:.3f} ms, {b/ms/1e6:.1f} GB/s")
print(f"assemble: {s.bench_assemble(3):.2f} ms")
