"""one INS assembly at cells^3 for a profiler run: python scripts/ncu_assemble.py [cells]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from util import cavity_prm

import openifem_b200 as ifem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
ifem.init(0)
tria = ifem.Triangulation(3)
ifem.GridGenerator.subdivided_hyper_rectangle(tria, (n, n, n), (0, 0, 0), (1, 1, 1), True)
flow = ifem.Fluid.MPI.InsIM(tria, ifem.Parameters.AllParameters(text=cavity_prm(3)))
flow.setup()
flow.set_vector(flow.EVALUATION_POINT, 0.1 * np.random.default_rng(0).uniform(-1, 1, flow.n_dofs))
flow.assemble(False)
flow.assemble(False)
