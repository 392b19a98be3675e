"""Sweep of the BCSR mat-vec kernel shape for short rows (Q1/Q1 SCnsIM system of the config-5 mesh at a reduced scale): block
mat-vec time and achieved GB/s per variant (key = 10 * lanes per row + unroll; 0 = default kernel).
    python scripts/spmv_short_sweep.py [scale]"""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import bench

import openifem_b200 as ifem

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ifem.init(0)
ftria, _ = bench.fsi_meshes(5, scale, 1)
fluid = ifem.Fluid.MPI.SCnsIM(ftria, ifem.Parameters.AllParameters(bench.fsi_prm_path(5)))
fluid.setup()
fluid.assemble(True)
for key in (0, 41, 44, 48, 81, 84, 161, 162, 321):
    ifem.set_spmv_short_variant(key)
    fluid.bench_vmult(3)
    ms, b = fluid.bench_vmult(20)
    print(f"variant {key:4d}: block mat-vec {ms:.4f} ms, {b / ms / 1e6:.0f} GB/s ({b / 1e9:.2f} GB)", flush=True)
