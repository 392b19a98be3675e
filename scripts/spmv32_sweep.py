"""Sweep IFEM_SPMV32_VARIANT (10 * lanes per row + min CTAs per SM) of the fp32-streamed velocity-block SpMV."""
import os
import subprocess
import sys

n = sys.argv[1] if len(sys.argv) > 1 else "128"
code = r'''
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from util import cavity_prm
import openifem_b200 as ifem
n = int(sys.argv[1])
ifem.init(0)
tria = ifem.Triangulation(3)
ifem.GridGenerator.subdivided_hyper_rectangle(tria, (n, n, n), (0, 0, 0), (1, 1, 1), True)
flow = ifem.Fluid.MPI.InsIM(tria, ifem.Parameters.AllParameters(text=cavity_prm(3)))
flow.setup(); flow.assemble(True)
flow.bench_spmv_uu(3); ms, b = flow.bench_spmv_uu(20)
flow.bench_spmv_uu_fp32(3); ms32, b32 = flow.bench_spmv_uu_fp32(20)
print(f"f64 {ms:.3f} ms {b/ms/1e6:.0f} GB/s | f32 {ms32:.3f} ms {b32/ms32/1e6:.0f} GB/s")
'''
for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1", "0", "1"]):
    env = dict(os.environ, IFEM_SPMV_L2HINT=v)
    r = subprocess.run([sys.executable, "-c", code, n], env=env, capture_output=True, text=True)
    print("variant", v, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
