"""Target for ncu: set up the 3-D cavity at <cells>^3, assemble once, then launch the kernels to be
profiled a few times (development / profiling aid; numbers printed under ncu are not bench values)."""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from util import cavity_prm

import openifem_b200 as ifem

n = int(sys.argv[1])
what = sys.argv[2] if len(sys.argv) > 2 else "spmv"
ifem.init(0)
tria = ifem.Triangulation(3)
ifem.GridGenerator.subdivided_hyper_rectangle(tria, (n, n, n), (0, 0, 0), (1, 1, 1), True)
flow = ifem.Fluid.MPI.InsIM(tria, ifem.Parameters.AllParameters(text=cavity_prm(3)))
flow.setup()
flow.assemble(True)
if what == "sell":
    # product kernels of the fp32 inner solvers (SELL-32 copies of A_uu): fp16 values, then fp32 values
    print("sell16", flow.bench_spmv_uu_sell(3, precision=16, check_error=False))
    print("sell32", flow.bench_spmv_uu_sell(3, precision=32, check_error=False))
    print("uu f64", flow.bench_spmv_uu(3))
elif what == "spmv":
    print("uu", flow.bench_spmv_uu(3))
    print("uu fp32", flow.bench_spmv_uu_fp32(3))
    print("block", flow.bench_vmult(2))
else:
    print("assemble", flow.bench_assemble(2))
