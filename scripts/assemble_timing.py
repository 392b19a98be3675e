"""Timing of InsIM::assemble on one GPU: python scripts/assemble_timing.py [cells] - ms per assembly (all colours, CUDA events),
achieved FP64 rate against the measured FMA peak and RMW traffic rate of the matrix scatter against the copy peak."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from util import cavity_prm

import openifem_b200 as ifem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ifem.init(0)
tria = ifem.Triangulation(3)
ifem.GridGenerator.subdivided_hyper_rectangle(tria, (n, n, n), (0, 0, 0), (1, 1, 1), True)
flow = ifem.Fluid.MPI.InsIM(tria, ifem.Parameters.AllParameters(text=cavity_prm(3)))
flow.setup()
rng = np.random.default_rng(0)
flow.set_vector(flow.EVALUATION_POINT, 0.1 * rng.uniform(-1, 1, flow.n_dofs))
flow.bench_assemble(1)
ms = flow.bench_assemble(5)
cells = n ** 3
peak = ifem.bench_fp64_peak()
flop = cells * (27 ** 3 * 25 * 2 + 60000)  # uu blocks: 25 FMA per (row node, column node, quadrature point); ~6e4 for the other phases
rmw = cells * 89 * 89 * 16.0
print(json.dumps({"cells": cells, "ms_per_assembly": ms, "fp64_peak_tflops": peak, "achieved_tflops": flop / (ms * 1e-3) / 1e12,
                  "frac_fp64": flop / (ms * 1e-3) / 1e12 / peak, "rmw_GBps": rmw / (ms * 1e-3) / 1e9}))
