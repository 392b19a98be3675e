"""One-process sweep of the product-kernel variants of the fp32 inner solver (SELL-32 copy of A_uu) at config-3
size, followed by a few time steps with the inner solver switched on. Usage: sell_sweep.py [cells] [steps]"""
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from util import cavity_prm

import openifem_b200 as ifem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ifem.init(0)
t0 = time.perf_counter()
tria = ifem.Triangulation(3)
ifem.GridGenerator.subdivided_hyper_rectangle(tria, (n, n, n), (0, 0, 0), (1, 1, 1), True)
flow = ifem.Fluid.MPI.InsIM(tria, ifem.Parameters.AllParameters(text=cavity_prm(3)))
flow.setup()
flow.assemble(True)
print(f"setup {time.perf_counter() - t0:.1f} s", flush=True)
flow.bench_spmv_uu(3)
ms, b = flow.bench_spmv_uu(10)
print(f"f64 BCSR  {ms:.3f} ms  {b / ms / 1e6:.0f} GB/s", flush=True)
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 3
best = {}
for prec, variants in [(32, [24]), (16, [24])]:
    if mode == 2 and prec == 16:
        continue
    t0 = time.perf_counter()
    best[prec] = (1e9, 0)
    for k, v in enumerate(variants):
        ms, b, pad, err = flow.bench_spmv_uu_sell(20, variant=v, check_error=(k == 0), precision=prec)
        if k == 0:
            print(f"SELL({prec}) setup + first product {time.perf_counter() - t0:.1f} s, padding {pad:.4f}", flush=True)
        print(f"sell{prec} variant {v:3d}  {ms:.3f} ms  {b / ms / 1e6:.0f} GB/s  err {err:.2e}", flush=True)
        best[prec] = min(best[prec], (ms, v))
print("best", best, flush=True)
flow.set_inner_variant(best[16 if mode == 3 else 32][1])
sm_mode = int(sys.argv[4]) if len(sys.argv) > 4 else 2
flow.set_control(a_inv_rel=1e-1, a_inv_fp32=mode, cg_sm_fp32=sm_mode, a_inv_max_it=int(sys.argv[5]) if len(sys.argv) > 5 else 2000)
flow.set_verbose(True)
for k in range(steps):
    t0 = time.perf_counter()
    flow.run_one_step(k == 0)
    print(f"step {k} wall {time.perf_counter() - t0:.2f}s", flush=True)
for sec in ["Assemble system", "Solve linear system", "CG for Mp", "CG for Sm", "A_inv"]:
    print(sec, f"{flow.timer_ms(sec):.1f} ms")
