"""Two-rank diagnostics of the peer link (openifem_b200/csrc/peer.h): launched by torchrun, prints per rank the error of the
SELL-32 product against the fp64 product (checks the ghost push), then FGMRES / inner iteration counts of two time steps.
    IFEM_PEER=3|2|1|0 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/peer_diag.py --cells 32"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=32)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    import numpy as np
    from util import cavity_prm

    import openifem_b200 as ifem

    rank, world = ifem.init_distributed()
    n = args.cells
    tria = ifem.Triangulation(3)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, (n, n, n), (0, 0, 0), (1, 1, 1), True)
    flow = ifem.Fluid.MPI.InsIM(tria, ifem.Parameters.AllParameters(text=cavity_prm(3)))
    flow.setup()
    flow.set_control(a_inv_rel=1e-1, a_inv_fp32=3, cg_sm_fp32=1, a_inv_max_it=400)
    rng = np.random.default_rng(1)
    flow.set_vector(flow.EVALUATION_POINT, 0.1 * rng.uniform(-1, 1, flow.n_dofs))
    flow.assemble(True)
    errs = [flow.bench_spmv_uu_sell(1, check_error=True, precision=p)[3] for p in (32, 16)]
    zero = np.zeros(flow.n_dofs)
    flow.set_vector(flow.EVALUATION_POINT, zero)
    flow.set_vector(flow.PRESENT, zero)
    for k in range(args.steps):
        flow.run_one_step(k == 0)
    h = flow.history()
    sol = flow.get_current_solution()
    print(f"[rank {rank}/{world}] IFEM_PEER={os.environ.get('IFEM_PEER', '(unset)')} sell product error fp32 {errs[0]:.2e} fp16 {errs[1]:.2e}; "
          f"fgmres its {[r['gmres_its'] for r in h]} a_inv its {[r['a_inv_its'] for r in h]} cg_sm its {[r['cg_sm_its'] for r in h]} "
          f"true_res {[float('%.1e' % r['true_res']) for r in h]} |sol| {np.linalg.norm(sol):.12e}", flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        ifem.comm_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
