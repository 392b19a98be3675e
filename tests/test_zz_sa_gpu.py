"""GPU parity of the Spalart-Allmaras turbulence model (reference source/mpi_spalart_allmaras.cpp, attached through
FluidSolver::attach_turbulence_model, source/mpi_fluid_solver.cpp:53-63) against the oracle (oracle/spalart_allmaras.py), all
through the C ABI (ifem_insim_attach_turbulence_model, ifem_turbulence_*).

Parity is UNPINNED on the reference side: no reference test attaches the model, and `r` of its destruction term is
indeterminate in the reference source (:757-770); product and oracle both use r = min(nu~ / (S~ kappa^2 d^2), 10).

Tolerances: wall distance 1e-14 relative; assembled matrix / rhs 1e-12 relative (two routes through the constraints: in-kernel
elimination + condensation on the device, C^T A C in the oracle); fields after time steps 1e-6 relative (device FGMRES to
1e-8 |rhs|, oracle sparse direct); shear velocity 1e-14."""
import numpy as np
import pytest
import scipy.sparse as sp

from test_hanging_oracle_cpu import refined_mesh
from test_scns_gpu import scns_prm

pytestmark = pytest.mark.gpu

SA = """
subsection Spalart Allmaras model
  set Number of S-A model BCs = 3
  set S-A model boundary id = 0, 2, 3
  set S-A model boundary types = 1, 0, 0
  set Initial condition coefficient = 3.0
  set Wall function image distance = 0.02
end
"""


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def _make(dim, refined=False, **kw):
    import openifem_b200 as ifem
    from oracle import fem, prm, scns

    text = scns_prm(dim, **kw) + SA
    if refined:
        tria, mesh = refined_mesh(dim)
    else:
        reps, hi = ((7, 5), (2.0, 1.0)) if dim == 2 else ((4, 3, 3), (2.0, 1.0, 0.9))
        mesh = fem.BoxMesh(reps, (0.0,) * dim, hi)
        tria = ifem.Triangulation(dim)
        ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0.0,) * dim, hi, True)
    o = scns.SCnsIM(mesh, prm.Params(text, is_text=True))
    g = ifem.Fluid.MPI.SCnsIM(tria, ifem.Parameters.AllParameters(text=text))
    g.setup()
    to, tg = o.attach_turbulence_model("Spalart-Allmaras"), g.attach_turbulence_model("Spalart-Allmaras")
    assert tg.n_dofs == to.n
    return o, g, to, tg


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("refined", [False, True])
def test_setup_wall_distance_and_initial_condition(dim, refined):
    o, g, to, tg = _make(dim, refined)
    d = tg.get_vector(tg.WALL_DISTANCE)
    assert np.abs(d - to.fixed_wall_distance).max() <= 1e-14 * to.fixed_wall_distance.max()
    assert to.fixed_wall_distance.max() > 0.2 and (to.fixed_wall_distance[to.con != 0] >= 0).all()
    assert np.abs(tg.get_vector(tg.PRESENT) - to.present).max() <= 1e-18
    assert to.present.max() > 0 and np.count_nonzero(to.present == 0.0) > 0  # the Dirichlet lines were distributed


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("refined", [False, True])
@pytest.mark.parametrize("nonzero", [True, False])
def test_assembly_matches_oracle(dim, refined, nonzero):
    o, g, to, tg = _make(dim, refined, mu=1e-3, rho=1.2, dt=1e-2)
    rng = np.random.default_rng(5 + dim)
    nu = to.nu_laminar
    # both branches of the model: a few negative values of nu~ (negative S-A), values far above the laminar viscosity
    pr = nu * rng.uniform(-0.5, 6.0, to.n)
    ev = pr + nu * rng.uniform(-0.3, 0.3, to.n)
    fl = rng.uniform(-1, 1, o.n)
    ind = (rng.uniform(size=o.mesh.n_cells) < 0.25).astype(np.int32)
    to.present[:], to.evaluation_point[:], o.present[:], o.indicator[:] = pr, ev, fl, ind
    tg.set_vector(tg.PRESENT, pr)
    tg.set_vector(tg.EVALUATION_POINT, ev)
    g.set_vector(g.PRESENT, fl)
    g.set_indicator(ind)
    A_ref, b_ref = to.assemble(nonzero)
    tg.assemble(nonzero)
    A, b = tg.get_matrix(), tg.get_vector(tg.SYSTEM_RHS)
    err_A = sp.linalg.norm(A - A_ref) / sp.linalg.norm(A_ref)
    assert err_A < 1e-12 and _rel(b, b_ref) < 1e-12, (err_A, _rel(b, b_ref))
    if refined:
        h = np.asarray(sorted(to.hanging))
        D = sp.diags(A.diagonal()).tocsr()
        assert h.size and abs(A.tocsr()[h] - D[h]).max() == 0.0 and abs(A.tocsc()[:, h] - D.tocsc()[:, h]).max() == 0.0


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("refined", [False, True])
def test_time_steps_with_the_model_attached(dim, refined):
    """SUPGFluidSolver::run with a turbulence model (source/mpi_supg_solver.cpp:456-468): model step, then fluid step reading the
    eddy viscosity (source/mpi_scnsim.cpp:198-216)"""
    kw = dict(mu=1e-3, rho=1.0, dt=1e-2, newton_tol=1e-8)
    o, g, to, tg = _make(dim, refined, **kw)
    o.prm.end_time = 3 * o.dt
    g.set_control(fgmres_rel=1e-10)
    o.run(max_steps=3)
    for k in range(3):
        tg.run_one_step(k == 0)
        g.run_one_step(k == 0)
    assert _rel(tg.get_vector(tg.PRESENT), to.present) < 1e-6
    assert _rel(tg.get_eddy_viscosity(), to.eddy_viscosity) < 1e-6 and to.eddy_viscosity.max() > 0.1 * o.prm.viscosity
    sol = g.get_current_solution()
    assert _rel(sol[: o.n_u], o.velocity()) < 1e-6 and _rel(sol[o.n_u:], o.pressure()) < 1e-6
    ho, hg = to.history, tg.history()
    assert len(ho) == len(hg)
    for a, b in zip(ho, hg):
        assert abs(a[1] - b[0]) <= 1e-6 * max(a[1], 1e-12) + 1e-13
    # the eddy viscosity changes the fluid step: the same steps without the model give a different field
    import openifem_b200 as ifem

    g0 = ifem.Fluid.MPI.SCnsIM(g.tria, g.params)
    g0.setup()
    g0.set_control(fgmres_rel=1e-10)
    for k in range(3):
        g0.run_one_step(k == 0)
    assert _rel(g0.get_current_solution()[: o.n_u], sol[: o.n_u]) > 1e-5


def test_run_drives_the_model():
    """ifem_insim_run: the model is advanced before every fluid step"""
    text = scns_prm(2, mu=1e-3, rho=1.0, dt=1e-2, newton_tol=1e-8).replace("set End time = 1.0", "set End time = 0.03") + SA
    import openifem_b200 as ifem
    from oracle import fem, prm, scns

    p = prm.Params(text, is_text=True)
    assert abs(p.end_time - 0.03) < 1e-12
    tria = ifem.Triangulation(2)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, (7, 5), (0.0, 0.0), (2.0, 1.0), True)
    g = ifem.Fluid.MPI.SCnsIM(tria, ifem.Parameters.AllParameters(text=text))
    tg = g.attach_turbulence_model("Spalart-Allmaras")  # before setup: initialised by initialize_system
    g.setup()
    g.set_control(fgmres_rel=1e-10)
    g.run()
    o = scns.SCnsIM(fem.BoxMesh((7, 5), (0.0, 0.0), (2.0, 1.0)), p)
    to = o.attach_turbulence_model("Spalart-Allmaras")
    o.run()
    assert o.timestep == 3 and _rel(tg.get_vector(tg.PRESENT), to.present) < 1e-6
    assert _rel(g.get_current_solution()[: o.n_u], o.velocity()) < 1e-6


def test_lines_of_cells_inside_the_solid():
    """update_boundary_condition (:133-224): nu~ is driven to zero on every node of a cell with indicator 1"""
    o, g, to, tg = _make(2, mu=1e-3, rho=1.0, dt=1e-2)
    ind = np.zeros(o.mesh.n_cells, dtype=np.int32)
    ind[[9, 10, 16, 17]] = 1
    o.indicator[:] = ind
    g.set_indicator(ind)
    for first in (True, False):
        to.make_constraints()
        to.update_boundary_condition(first)
        to.run_one_step(True)
        tg.update_boundary_condition(first)  # restores the model's boundary lines first, like the reference's per-step make_constraints
        tg.run_one_step(True)
        x = tg.get_vector(tg.PRESENT)
        nodes = np.unique(to.nodes[ind == 1])
        assert np.abs(x[nodes]).max() < 1e-18 and _rel(x, to.present) < 1e-6


def test_shear_velocity():
    o, g, to, tg = _make(2, mu=1.8e-5, rho=1.2)
    for vel, guess in [(0.0, 0.1), (1e-3, 0.0), (0.5, 0.0), (5.0, 0.3), (40.0, 1.0)]:
        a, b = tg.get_shear_velocity(vel, guess), to.get_shear_velocity(vel, guess)
        assert abs(a - b) <= 1e-14 * max(abs(b), 1.0), (vel, a, b)
    assert to.get_shear_velocity(5.0, 0.3) > 0.1


def test_unknown_model_is_rejected():
    import openifem_b200 as ifem

    o, g, to, tg = _make(2)
    with pytest.raises(ifem.IfemError):
        g.attach_turbulence_model("k-epsilon")


def test_coupled_fsi_steps_with_the_model():
    """MPI::FSI::run with a turbulence model attached to the fluid solver (source/mpi_fsi.cpp:1199-1210): per pass the model's lines
    are re-made, the cells inside the solid get nu~ -> 0, the model steps, then the fluid steps with the eddy viscosity"""
    import openifem_b200 as ifem
    from oracle import fem, fsi, prm, scns, solid
    from test_fsi_gpu import _fsi_text

    dim, f_reps, s_reps, s_lo, s_hi = 2, (12, 12), (4, 6), (0.3125, 0.0), (0.5625, 0.6875)
    text = _fsi_text(dim) + SA
    lo, hi = (0.0,) * dim, (1.0,) * dim
    P = prm.Params(text, is_text=True)
    o_fluid = scns.SCnsIM(fem.BoxMesh(f_reps, lo, hi), P)
    to = o_fluid.attach_turbulence_model("Spalart-Allmaras")
    o_solid = solid.HyperElasticity(fem.BoxMesh(s_reps, s_lo, s_hi), P)
    ftria = ifem.Triangulation(dim)
    ifem.GridGenerator.subdivided_hyper_rectangle(ftria, f_reps, lo, hi, True)
    params = ifem.Parameters.AllParameters(text=text)
    fluid = ifem.Fluid.MPI.SCnsIM(ftria, params)
    tg = fluid.attach_turbulence_model("Spalart-Allmaras")
    fluid.setup()
    fluid.set_control(fgmres_rel=1e-10)
    stria = ifem.Triangulation(dim)
    ifem.GridGenerator.subdivided_hyper_rectangle(stria, s_reps, s_lo, s_hi, True)
    sol = ifem.Solid.MPI.HyperElasticity(stria, params)
    sol.setup()
    coupling = ifem.MPI.FSI(fluid, sol, params, False)
    loop = fsi.FSI(o_fluid, o_solid, False)
    for k in range(2):
        loop.run_one_step(k == 0)
        coupling.run_one_step(k == 0)
    inside = np.unique(to.nodes[o_fluid.indicator == 1])
    x = tg.get_vector(tg.PRESENT)
    assert inside.size and np.abs(x[inside]).max() < 1e-18 and x.max() > 0
    assert _rel(x, to.present) < 1e-6
    fsol = fluid.get_current_solution()
    assert _rel(fsol[: o_fluid.n_u], o_fluid.velocity()) < 1e-6 and _rel(fsol[o_fluid.n_u:], o_fluid.pressure()) < 1e-6
    assert _rel(sol.get_current_solution(), o_solid.cur_u) < 1e-6


def test_refine_mesh_carries_the_model():
    """FSI::refine_mesh with a turbulence model (source/mpi_fsi.cpp:1093-1096, 1113-1116): nu~ is transferred like the fluid's
    solution, the model's lines / wall distances / system are rebuilt on the new mesh; coupled steps before and after agree"""
    import openifem_b200 as ifem
    from oracle import fem, fsi, prm, scns, solid
    from test_fsi_gpu import _fsi_text

    reps, s_reps, s_lo, s_hi = (12, 12), (4, 6), (0.3125, 0.0), (0.5625, 0.6875)
    text = _fsi_text(2) + SA
    P = prm.Params(text, is_text=True)
    o_fluid = scns.SCnsIM(fem.BoxMesh(reps, (0.0, 0.0), (1.0, 1.0)), P)
    o_fluid.attach_turbulence_model("Spalart-Allmaras")
    o_solid = solid.HyperElasticity(fem.BoxMesh(s_reps, s_lo, s_hi), P)
    params = ifem.Parameters.AllParameters(text=text)
    ftria, stria, otria = ifem.Triangulation(2), ifem.Triangulation(2), ifem.Triangulation(2)
    for t in (ftria, otria):
        ifem.GridGenerator.subdivided_hyper_rectangle(t, reps, (0.0, 0.0), (1.0, 1.0), True)
    ifem.GridGenerator.subdivided_hyper_rectangle(stria, s_reps, s_lo, s_hi, True)
    fluid = ifem.Fluid.MPI.SCnsIM(ftria, params)
    fluid.setup()
    tg = fluid.attach_turbulence_model("Spalart-Allmaras")
    fluid.set_control(fgmres_rel=1e-10)
    sol = ifem.Solid.MPI.HyperElasticity(stria, params)
    sol.setup()
    coupling = ifem.MPI.FSI(fluid, sol, params, False)
    loop = fsi.FSI(o_fluid, o_solid, False)

    def refine_both():
        loop.refine_mesh(otria, 0, 2)
        coupling.refine_mesh(0, 2)
        assert all(np.array_equal(x, y) for x, y in zip(otria.get_mesh(), ftria.get_mesh()))
        to = loop.fluid.turbulence_model
        assert tg.n_dofs == to.n and _rel(tg.get_vector(tg.PRESENT), to.present) < 1e-13
        assert np.abs(tg.get_vector(tg.WALL_DISTANCE) - to.fixed_wall_distance).max() < 1e-14
        fluid.set_control(fgmres_rel=1e-10)

    refine_both()
    assert ftria.hanging()[0].size > 0
    for k in range(3):
        loop.run_one_step(k == 0)
        coupling.run_one_step(k == 0)
        if k == 1:
            refine_both()  # nu~ is no longer uniform here: wall and solid lines have acted for two steps
    of, to = loop.fluid, loop.fluid.turbulence_model
    fsol = fluid.get_current_solution()
    assert to.present.max() > 0 and _rel(tg.get_vector(tg.PRESENT), to.present) < 1e-6
    assert _rel(fsol[: of.n_u], of.velocity()) < 1e-6 and _rel(fsol[of.n_u:], of.pressure()) < 1e-6


def test_no_wall_boundary_and_no_model():
    """edge cases: without a wall boundary the wall distance is DBL_MAX (:521) - a uniform nu~ at rest is then a steady state of the
    transport equation; the ifem_turbulence_* entry points fail loudly on a solver that has no model attached"""
    import openifem_b200 as ifem
    from oracle import fem, prm, scns

    text = scns_prm(2, mu=1e-3, rho=1.0, dt=1e-2) + SA.replace("= 3\n", "= 0\n", 1).replace("= 0, 2, 3", "= 0").replace("= 1, 0, 0", "= 0")
    o = scns.SCnsIM(fem.BoxMesh((5, 4), (0.0, 0.0), (1.0, 1.0)), prm.Params(text, is_text=True))
    to = o.attach_turbulence_model("Spalart-Allmaras")
    assert not to.sa["bcs"] and to.con.sum() == 0
    tria = ifem.Triangulation(2)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, (5, 4), (0.0, 0.0), (1.0, 1.0), True)
    g = ifem.Fluid.MPI.SCnsIM(tria, ifem.Parameters.AllParameters(text=text))
    g.setup()
    with pytest.raises(ifem.IfemError):
        ifem._TurbulenceModel(g).get_vector(0)
    tg = g.attach_turbulence_model("Spalart-Allmaras")
    assert (tg.get_vector(tg.WALL_DISTANCE) == np.finfo(np.float64).max).all()
    tg.set_vector(tg.EVALUATION_POINT, tg.get_vector(tg.PRESENT))  # run_one_step starts from evaluation_point = present_solution (:307)
    tg.assemble(False)
    assert np.abs(tg.get_vector(tg.SYSTEM_RHS)).max() < 1e-18
    tg.run_one_step(True)
    assert np.allclose(tg.get_vector(tg.PRESENT), 3.0 * to.nu_laminar, rtol=1e-14, atol=0)
    A_ref, _ = to.assemble(False)
    assert sp.linalg.norm(tg.get_matrix() - A_ref) < 1e-12 * sp.linalg.norm(A_ref)
