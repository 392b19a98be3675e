"""A/B of the SCnsIM cell-kernel variants (csrc/scnsim.cu, IFEM_SCNS_ASM) on the fluid mesh of config 5 at `scale`: device time of
`reps` assemblies per variant, and the assembled right-hand side / system mat-vec of every variant against variant 0.
    python scripts/probe_scns_assemble.py [scale] [reps]            (no torch import)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    import numpy as np

    import openifem_b200 as ifem
    from bench import fsi_meshes, fsi_prm_path

    ifem.init(0)
    ftria, _ = fsi_meshes(5, scale, 1)
    fluid = ifem.Fluid.MPI.SCnsIM(ftria, ifem.Parameters.AllParameters(fsi_prm_path(5)))
    fluid.setup()
    n = fluid.n_dofs
    k = np.arange(n)
    fluid.set_vector(fluid.EVALUATION_POINT, 0.3 * np.sin(0.013 * k) + 0.1)
    fluid.set_vector(fluid.PRESENT, 0.28 * np.sin(0.013 * k + 0.05) + 0.1)
    fluid.set_vector(fluid.FSI_ACCELERATION, 0.2 * np.cos(0.007 * k))
    ind = (np.arange(ftria.n_active_cells()) % 11 == 0).astype(np.int32)
    fluid.set_indicator(ind)
    x = np.cos(0.021 * k)
    out = {"scale": scale, "cells": int(ftria.n_active_cells()), "dofs": int(n), "variants": {}}
    ref = None
    for v in (0, 1, 2, 3, 4, 0):
        os.environ["IFEM_SCNS_ASM"] = str(v)
        fluid.assemble(True)  # warm
        t0 = fluid.timer_ms("Assemble system")
        for _ in range(reps):
            fluid.assemble(True)
        ms = (fluid.timer_ms("Assemble system") - t0) / reps
        rhs, y = fluid.get_vector(fluid.SYSTEM_RHS), fluid.vmult(x)
        if ref is None:
            ref = (rhs, y)
        rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
        out["variants"].setdefault(str(v), []).append({"ms_per_assembly": ms, "ns_per_cell": 1e6 * ms / out["cells"],
                                                       "rhs_vs_v0": rel(rhs, ref[0]), "matvec_vs_v0": rel(y, ref[1])})
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
