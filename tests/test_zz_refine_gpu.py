"""FSI::refine_mesh on the device path (reference source/mpi_fsi.cpp:1024-1117, called from FSI::run at :1164-1168 and
:1215-1218): cells near the boundary of the deformed solid are refined, the others coarsened (between Global refinements and
Global refinements + 3), the fluid solution is transferred (SolutionTransfer) and the fluid solver set up again.

Oracle: oracle/fsi.py FSI.refine_mesh - flags and transfer (old FE field evaluated at the new support points) restated there; the
mesh operation itself (families, 2:1 balance) is the product's host-side mesh class on both sides, pinned by
tests/test_grid_cpu.py. The reference pins nothing here (no test refines an FSI case): "parity unpinned".
Tolerance: meshes identical, transferred solution 1e-13, fields after coupled steps on the adapted mesh 1e-6."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def test_refine_mesh_flags_transfer_and_steps_match_oracle():
    import openifem_b200 as ifem
    from oracle import fsi
    from test_fsi_gpu import _fsi_pair

    reps, s_lo, s_hi = (12, 12), (0.3125, 0.0), (0.5625, 0.6875)
    o_fluid, o_solid, fluid, sol, coupling = _fsi_pair(2, reps, (4, 6), s_lo, s_hi, True)
    otria = ifem.Triangulation(2)  # the oracle's own copy of the fluid triangulation
    ifem.GridGenerator.subdivided_hyper_rectangle(otria, reps, (0.0, 0.0), (1.0, 1.0), True)
    fluid.set_control(fgmres_rel=1e-10)
    loop = fsi.FSI(o_fluid, o_solid, True)

    def refine_both():
        loop.refine_mesh(otria, 0, 3)
        coupling.refine_mesh(0, 3)
        a, b = otria.get_mesh(), fluid.tria.get_mesh()
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        assert fluid.n_dofs == loop.fluid.n
        assert _rel(fluid.get_current_solution(), loop.fluid.present) < 1e-13 or np.abs(loop.fluid.present).max() == 0

    refine_both()  # FSI::run refines twice before the first step (:1164-1168)
    refine_both()
    lv = fluid.tria.levels()
    assert lv.max() == 2 and lv.min() == 0 and fluid.tria.hanging()[0].size > 0
    n_cells = [fluid.tria.n_active_cells()]
    for k in range(4):
        loop.run_one_step(k == 0)
        coupling.run_one_step(k == 0)
        if k == 1:  # a refinement interval of two steps (:1215-1218): the solution is nonzero now, the solid has moved
            refine_both()
            n_cells.append(fluid.tria.n_active_cells())
            fluid.set_control(fgmres_rel=1e-10)
    of = loop.fluid
    fsol = fluid.get_current_solution()
    assert np.abs(of.velocity()).max() > 0 and lv.max() >= 2
    assert _rel(fsol[: of.n_u], of.velocity()) < 1e-6 and _rel(fsol[of.n_u:], of.pressure()) < 1e-6
    assert _rel(sol.get_current_solution(), o_solid.cur_u) < 1e-6


def test_fsi_run_refines_like_the_reference_driver_loop():
    """FSI::run with `Refinement interval` < `End time`: two refinements before the first step and one at every interval (:1164-1168,
    :1215-1218) - the loop runs through and leaves a locally refined, balanced mesh around the solid"""
    import openifem_b200 as ifem
    from test_fsi_gpu import _fsi_text

    text = _fsi_text(2).replace("set End time = 1.0", "set End time = 3e-3").replace("set Refinement interval = 1e6", "set Refinement interval = 2e-3")
    assert "End time = 3e-3" in text and "Refinement interval = 2e-3" in text
    params = ifem.Parameters.AllParameters(text=text)
    ftria, stria = ifem.Triangulation(2), ifem.Triangulation(2)
    ifem.GridGenerator.subdivided_hyper_rectangle(ftria, (8, 8), (0.0, 0.0), (1.0, 1.0), True)
    ifem.GridGenerator.subdivided_hyper_rectangle(stria, (2, 3), (0.375, 0.0), (0.625, 0.5), True)
    fluid, solid = ifem.Fluid.MPI.SCnsIM(ftria, params), ifem.Solid.MPI.SharedHyperElasticity(stria, params)
    fluid.setup()
    solid.setup()
    coupling = ifem.MPI.FSI(fluid, solid, params, False)
    coupling.run()
    lv = ftria.levels()
    assert lv.max() == 3 and lv.min() == 0  # three refinements, capped at Global refinements + 3
    assert ftria.hanging()[0].size > 0 and fluid.n_dofs == 3 * ftria.n_vertices()
    assert len({r["timestep"] for r in fluid.history()}) == 3
    assert np.isfinite(fluid.get_current_solution()).all()
