"""BASELINE config 5 at test size - the reference's tests/fsi-wall-3D case (fsi-wall-3D.cpp:33-63; .prm in
tests/golden/fsi_wall_3d.prm with the NeoHookean plate SURVEY 8(d) prescribes) at half its resolution: MPI::FSI<3>(SCnsIM Q1/Q1 on
{5,5,20} cells of [0,1]^2 x [0,4] whose band 2 <= z <= 2.4 is refined once - hanging nodes on two planes -, pressure 5e2 on
boundary id 4, SharedHyperElasticity plate of {10,10,4} cells at z in [2, 2.4], use_dirichlet_bc = false), dt = 1e-6.
bench.py --config 5 times the same case at {70,70,280} cells (9.6 M fluid DoF) on 1 and 8 GPUs.

The reference pins nothing for this case (it is not even registered as a test: "parity unpinned"); the device loop is compared with
the oracle's FSI loop step by step: indicator exact, fluid velocity / pressure and solid displacement 1e-6 relative (device FGMRES
to 1e-10 |rhs| for the comparison, oracle sparse direct)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def test_fsi_wall_3d_steps_match_oracle(golden_dir):
    import openifem_b200 as ifem
    from oracle import fem, fsi, grid, prm, scns, solid

    path = os.path.join(golden_dir, "fsi_wall_3d.prm")
    params = ifem.Parameters.AllParameters(path)
    ftria = ifem.Triangulation(3)
    ifem.GridGenerator.subdivided_hyper_rectangle(ftria, (5, 5, 20), (0, 0, 0), (1, 1, 4), True)
    v, c, _ = ftria.get_mesh()
    cz = v[c].mean(axis=1)[:, 2]
    ftria.execute_refinement(((cz >= 2) & (cz <= 2.4)).astype(np.uint8))
    stria = ifem.Triangulation(3)
    ifem.GridGenerator.subdivided_hyper_rectangle(stria, (10, 10, 4), (0, 0, 2), (1, 1, 2.4), True)
    fluid = ifem.Fluid.MPI.SCnsIM(ftria, params)
    fluid.setup()
    fluid.set_control(fgmres_rel=1e-10)
    sol = ifem.Solid.MPI.SharedHyperElasticity(stria, params)
    sol.setup()
    coupling = ifem.MPI.FSI(fluid, sol, params, False)

    P = prm.Params(path)
    v, c, b = ftria.get_mesh()
    o_fluid = scns.SCnsIM(grid.HexMesh(v, c, b), P)
    o_solid = solid.HyperElasticity(fem.BoxMesh((10, 10, 4), (0, 0, 2), (1, 1, 2.4)), P)
    loop = fsi.FSI(o_fluid, o_solid, False)
    assert ftria.n_active_cells() == 500 + 7 * 50 and len(o_fluid.dofs.hanging_u) > 0 and fluid.n_dofs == o_fluid.n
    for k in range(2):
        loop.run_one_step(k == 0)
        coupling.run_one_step(k == 0)
        assert np.array_equal(coupling.get_indicator(), o_fluid.indicator)
        fsol = fluid.get_current_solution()
        eu, ep = _rel(fsol[: o_fluid.n_u], o_fluid.velocity()), _rel(fsol[o_fluid.n_u:], o_fluid.pressure())
        assert eu < 1e-6 and ep < 1e-6, (k, eu, ep)
        us = sol.get_current_solution()
        # (the pressure wave has not reached the plate after two steps of 1e-6: the plate moves by round-off, 1e-23)
        assert np.abs(us - o_solid.cur_u).max() <= 1e-6 * np.abs(o_solid.cur_u).max() + 1e-18
    assert o_fluid.indicator.sum() > 0 and np.abs(o_fluid.pressure()).max() > 0
