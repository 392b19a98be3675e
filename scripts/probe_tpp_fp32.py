"""A/B of the inner T_pp solve of the SUPG block preconditioner on fp32 copies of the blocks (control.a_inv_fp32 on SCnsIM,
csrc/scnsim.cu precondition_supg) on a config-5-shaped FSI case (fsi-wall-3D at `scale`): the same coupled object runs
warm-up + `steps` time steps with fp64 blocks, then `steps` more with fp32 blocks; per-step device time of the sections and the
inner / outer iteration counts are printed as one JSON line. No torch import (a cold import costs more than the measurement).
    python scripts/probe_tpp_fp32.py [scale] [steps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    import numpy as np

    import openifem_b200 as ifem
    from bench import fsi_meshes, fsi_prm_path

    ifem.init(0)
    t0 = time.perf_counter()
    ftria, stria = fsi_meshes(5, scale, 1)
    params = ifem.Parameters.AllParameters(fsi_prm_path(5))
    fluid = ifem.Fluid.MPI.SCnsIM(ftria, params)
    fluid.setup()
    solid = ifem.Solid.MPI.SharedHyperElasticity(stria, params)
    solid.setup()
    coupling = ifem.MPI.FSI(fluid, solid, params, False)
    out = {"scale": scale, "fluid_cells": int(ftria.n_active_cells()), "fluid_dofs": int(fluid.n_dofs), "setup_s": time.perf_counter() - t0}
    secs = ["Assemble system", "Solve linear system", "Solving Tpp"]
    coupling.run_one_step(True)  # warm-up
    fields = {}
    for mode in (0, 1):
        fluid.set_control(a_inv_fp32=mode)
        if mode == 1:
            coupling.run_one_step(False)  # first step of the mode allocates the copies
        before = {k: fluid.timer_ms(k) for k in secs}
        n_hist = len(fluid.history())
        t = time.perf_counter()
        for _ in range(steps):
            coupling.run_one_step(False)
        wall = (time.perf_counter() - t) / steps
        h = fluid.history()[n_hist:]
        out[f"mode{mode}"] = {"wall_s_per_step": wall, **{k: (fluid.timer_ms(k) - before[k]) / steps for k in secs},
                              "fgmres_its": [r["gmres_its"] for r in h], "inner_tpp_its": [r["a_inv_its"] for r in h],
                              "abs_res": [r["abs_res"] for r in h]}
        fields[mode] = fluid.get_current_solution()
    out["finite"] = bool(np.isfinite(fields[1]).all())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
