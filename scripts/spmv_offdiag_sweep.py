"""Sweep of the BCSR mat-vec kernel shape for the off-diagonal blocks of the config-3 system (A_up, A_pu; A_uu keeps its tuned
kernel): block mat-vec time per variant (key = 10 * lanes per row + unroll; 0 = defaults).  python scripts/spmv_offdiag_sweep.py [cells]"""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from util import cavity_prm

import openifem_b200 as ifem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ifem.init(0)
tria = ifem.Triangulation(3)
ifem.GridGenerator.subdivided_hyper_rectangle(tria, (n, n, n), (0, 0, 0), (1, 1, 1), True)
flow = ifem.Fluid.MPI.InsIM(tria, ifem.Parameters.AllParameters(text=cavity_prm(3)))
flow.setup()
flow.assemble(True)
ms_uu, b_uu = flow.bench_spmv_uu(10)
print(f"A_uu alone: {ms_uu:.3f} ms ({b_uu / 1e9:.1f} GB)", flush=True)
for key in (0, 81, 84, 161, 162, 321):
    ifem.set_spmv_short_variant(key)
    flow.bench_vmult(2)
    ms, b = flow.bench_vmult(10)
    print(f"variant {key:4d}: block mat-vec {ms:.3f} ms, {b / ms / 1e6:.0f} GB/s; off-diagonal part {ms - ms_uu:.3f} ms for {(b - b_uu) / 1e9:.1f} GB = "
          f"{(b - b_uu) / (ms - ms_uu) / 1e6:.0f} GB/s", flush=True)
