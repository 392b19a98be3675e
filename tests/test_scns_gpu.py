"""GPU parity of the slightly compressible path: SCnsIM::assemble
(source/mpi_scnsim.cpp:15-568), FluidSolver::update_stress (source/mpi_fluid_solver.cpp:716-811) and the
SUPGFluidSolver Newton / FGMRES loop (source/mpi_supg_solver.cpp:297-425) against oracle/scns.py, plus the
reference goldens tests/fluid_initial_condition_mpi (max p = 1e4, 1e-8) and tests/fluid_body_force_mpi
(p_max - p_min = 1e3, 1e-3) through the device path.

Tolerances: assembled matrix / rhs / nodal stress 1e-12 relative; fields after time steps 1e-6 relative
(device: FGMRES to 1e-10 |rhs| for this comparison; oracle: sparse direct)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

SCNS_PRM = """
subsection Simulation
  set Simulation type = Fluid
  set Dimension = {dim}
  set Global refinements = 0, 0
  set End time = 1.0
  set Time step size = {dt}
  set Output interval = 1e6
  set Refinement interval = 1e6
  set Save interval = 1e6
  set Gravity = {gravity}
  set Initial velocity = {zeros}
end
subsection Fluid finite element system
  set Pressure degree = 1
  set Velocity degree = 1
end
subsection Fluid material properties
  set Dynamic viscosity = {mu}
  set Fluid density = {rho}
end
subsection Fluid solver control
  set Grad-Div stabilization = 0.1
  set Max Newton iterations = 10
  set Nonlinear system tolerance = {newton_tol}
end
subsection Fluid Dirichlet BCs
  set Use hard-coded boundary values = 0
  set Number of Dirichlet BCs = {n_dir}
  set Dirichlet boundary id = {dir_ids}
  set Dirichlet boundary components = {dir_comps}
  set Dirichlet boundary values = {dir_vals}
end
subsection Fluid Neumann BCs
  set Number of Neumann BCs = {n_neu}
  set Neumann boundary id = {neu_ids}
  set Neumann boundary values = {neu_vals}
end
subsection Solid material properties
  set Solid type = NeoHookean
  set Solid density = {rho_s}
  set Hyperelastic parameters = 1.0e3, 1.0e4
end
"""


def scns_prm(dim, dt=1e-3, mu=1.8e-4, rho=1.3e-3, rho_s=1.2, gravity=None, dirichlet=None, neumann=None, newton_tol=1e-8):
    if dirichlet is None:
        full = 3 if dim == 2 else 7
        dirichlet = {i: (full, [0.0] * dim) for i in range(2 * dim)}
        dirichlet[0] = (full, [1.0] + [0.0] * (dim - 1))
        del dirichlet[1]  # open outflow
    neumann = neumann or {}
    gravity = gravity or [0.0] * dim
    ids = sorted(dirichlet)
    return SCNS_PRM.format(dim=dim, dt=dt, mu=mu, rho=rho, rho_s=rho_s, zeros=", ".join(["0.0"] * dim),
                           gravity=", ".join(str(g) for g in gravity), newton_tol=newton_tol, n_dir=len(ids),
                           dir_ids=", ".join(str(i) for i in ids), dir_comps=", ".join(str(dirichlet[i][0]) for i in ids),
                           dir_vals=", ".join(str(v) for i in ids for v in dirichlet[i][1]), n_neu=len(neumann),
                           neu_ids=", ".join(str(i) for i in sorted(neumann)) or "0",
                           neu_vals=", ".join(str(neumann[i]) for i in sorted(neumann)) or "0")


def _make(text, reps, lo, hi, body_force=None, sigma=None, ic=None):
    import openifem_b200 as ifem
    from oracle import fem, prm, scns

    dim = len(reps)
    o = scns.SCnsIM(fem.BoxMesh(reps, lo, hi), prm.Params(text, is_text=True), body_force=body_force, sigma_pml_field=sigma,
                    initial_condition=ic)
    tria = ifem.Triangulation(dim)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, lo, hi, True)
    g = ifem.Fluid.MPI.SCnsIM(tria, ifem.Parameters.AllParameters(text=text))
    if body_force:
        g.set_body_force(body_force)
    if sigma:
        g.set_sigma_pml_field(sigma)
    if ic:
        g.set_initial_condition(ic)
    g.setup()
    return o, g


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


BF = lambda p, c: (30.0 * (1 + p[0]) if c == 0 else -12.0 * p[1])
SIG = lambda p, c: 50.0 * p[0] ** 2


@pytest.mark.parametrize("dim,reps,hi", [(2, (6, 5), (1.0, 0.8)), (3, (3, 4, 3), (1.0, 1.2, 0.9))])
@pytest.mark.parametrize("nonzero", [True, False])
def test_scns_assembly_matches_oracle(dim, reps, hi, nonzero):
    text = scns_prm(dim, gravity=[1.0, -9.8, 0.5][:dim], neumann={1: 3.5})
    o, g = _make(text, reps, (0,) * dim, hi, body_force=BF, sigma=SIG)
    rng = np.random.default_rng(3)
    ev, pr = rng.uniform(-1, 1, o.n), rng.uniform(-1, 1, o.n)
    acc = rng.uniform(-1, 1, o.n)
    ind = (rng.uniform(size=o.mesh.n_cells) < 0.4).astype(np.int32)
    stress = rng.uniform(-1, 1, o.stress.shape)
    fsis = rng.uniform(-1, 1, o.fsi_stress.shape)
    o.evaluation_point[:], o.present[:], o.fsi_acceleration[:], o.indicator[:] = ev, pr, acc, ind
    o.stress, o.fsi_stress = stress.copy(), fsis.copy()
    g.set_vector(g.EVALUATION_POINT, ev)
    g.set_vector(g.PRESENT, pr)
    g.set_vector(g.FSI_ACCELERATION, acc)
    g.set_indicator(ind)
    g.set_field(0, stress)
    g.set_field(1, fsis)
    A_ref, rhs_ref = o.assemble(nonzero)
    g.assemble(nonzero)
    A = g.get_matrix(0)
    assert sp.linalg.norm(A - A_ref) / sp.linalg.norm(A_ref) < 1e-12
    assert _rel(g.get_vector(g.SYSTEM_RHS), rhs_ref) < 1e-12
    x = rng.uniform(-1, 1, o.n)
    assert _rel(g.vmult(x), A_ref @ x) < 1e-12


@pytest.mark.parametrize("dim,reps,hi", [(2, (6, 5), (1.0, 0.8)), (3, (3, 4, 3), (1.0, 1.2, 0.9))])
def test_update_stress_matches_oracle(dim, reps, hi):
    o, g = _make(scns_prm(dim), reps, (0,) * dim, hi)
    rng = np.random.default_rng(4)
    pr = rng.uniform(-1, 1, o.n)
    o.present[:] = pr
    g.set_vector(g.PRESENT, pr)
    ref = o.update_stress()
    g.update_stress()
    assert _rel(g.get_field(0), ref) < 1e-12


@pytest.mark.parametrize("dim,reps,hi", [(2, (10, 6), (2.0, 1.0)), (3, (5, 4, 4), (2.0, 1.0, 1.0))])
def test_scns_time_steps_match_oracle(dim, reps, hi):
    o, g = _make(scns_prm(dim, dt=1e-3), reps, (0,) * dim, hi, body_force=lambda p, c: 5.0 if c == 0 else 0.0)
    g.set_control(fgmres_rel=1e-10)
    for k in range(3):
        o.run_one_step(k == 0)
        g.run_one_step(k == 0)
    sol = g.get_current_solution()
    assert _rel(sol[: o.n_u], o.velocity()) < 1e-6
    assert _rel(sol[o.n_u:], o.pressure()) < 1e-6
    assert _rel(g.get_field(0), o.stress) < 1e-6
    assert [(h["timestep"], h["iteration"]) for h in g.history()] == [(h[0], h[1]) for h in o.history]


@pytest.mark.parametrize("dim,reps,hi", [(2, (6, 5), (1.0, 0.8)), (3, (3, 4, 3), (1.0, 1.2, 0.9))])
@pytest.mark.parametrize("ilu", [0, 1])
def test_inner_tpp_solve_on_fp32_blocks(dim, reps, hi, ilu):
    """control.a_inv_fp32 = 1 on a SUPG solver: the products of the INNER T_pp solve (1e-3, inside the preconditioner of a flexible
    GMRES) stream fp32 copies of A_vp / A_pv / A_pp; operator, residuals and Krylov bases of the outer solve stay fp64, so the
    converged fields agree with the oracle to the same 1e-6 and the outer iteration counts do not move"""
    text = scns_prm(dim, dt=1e-3)
    o, g = _make(text, reps, (0,) * dim, hi, body_force=lambda p, c: 5.0 if c == 0 else 0.0)
    _, g64 = _make(text, reps, (0,) * dim, hi, body_force=lambda p, c: 5.0 if c == 0 else 0.0)
    g.set_control(fgmres_rel=1e-10, a_inv_fp32=1, supg_ilu=ilu)
    g64.set_control(fgmres_rel=1e-10, supg_ilu=ilu)
    for k in range(3):
        o.run_one_step(k == 0)
        g.run_one_step(k == 0)
        g64.run_one_step(k == 0)
    sol = g.get_current_solution()
    assert _rel(sol[: o.n_u], o.velocity()) < 1e-6 and _rel(sol[o.n_u:], o.pressure()) < 1e-6
    h32, h64 = g.history(), g64.history()
    assert [(h["timestep"], h["iteration"]) for h in h32] == [(h[0], h[1]) for h in o.history]
    assert all(abs(a["gmres_its"] - b["gmres_its"]) <= 1 for a, b in zip(h32, h64))


@pytest.mark.parametrize("dim,reps,hi", [(2, (6, 5), (1.0, 0.8)), (3, (3, 4, 3), (1.0, 1.2, 0.9))])
def test_cell_kernel_variants_agree(dim, reps, hi, monkeypatch):
    """IFEM_SCNS_ASM: 0 = the cell kernel with the reference's divisions by dt / atm / kappa_s, 1 = reciprocals + batched scatter
    (the 3-D default, profiles/r02_scns_assemble_ncu_summary.md), 2-4 = its register / arithmetic variants: one system to round-off"""
    import scipy.sparse as sp

    o, g = _make(scns_prm(dim, gravity=[0.3, -9.8, 0.5][:dim], neumann={1: 2.0}), reps, (0,) * dim, hi, body_force=BF, sigma=SIG)
    rng = np.random.default_rng(2)
    g.set_vector(g.EVALUATION_POINT, rng.uniform(-1, 1, o.n))
    g.set_vector(g.PRESENT, rng.uniform(-1, 1, o.n))
    g.set_vector(g.FSI_ACCELERATION, rng.uniform(-1, 1, o.n))
    g.set_indicator((rng.uniform(size=o.mesh.n_cells) < 0.4).astype(np.int32))
    out = {}
    for v in (0, 1, 2, 3, 4):
        monkeypatch.setenv("IFEM_SCNS_ASM", str(v))
        g.assemble(True)
        out[v] = (g.get_matrix(0), g.get_vector(g.SYSTEM_RHS))
    A0, b0 = out[0]
    for v in (1, 2, 3, 4):
        A, b = out[v]
        assert sp.linalg.norm(A - A0) <= 1e-14 * sp.linalg.norm(A0) and _rel(b, b0) < 1e-14, v


def test_initial_condition_reference_golden(golden_dir):
    """tests/fluid_initial_condition_mpi/fluid_initial_condition_mpi.cpp:31-60"""
    import openifem_b200 as ifem

    def ic(p, c):
        if c == 2:
            if 4.0 < p[0] < 5.0:
                return 1e4 * (p[0] - 4.0)
            if 5.0 <= p[0] < 12.0:
                return 1e4
        return 0.0

    tria = ifem.Triangulation(2)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, (150, 20), (0, 0), (15, 2), True)
    flow = ifem.Fluid.MPI.SCnsIM(tria, ifem.Parameters.AllParameters(os.path.join(golden_dir, "scns_initial_condition_2d.prm")))
    flow.set_initial_condition(ic)
    flow.run()
    p = flow.get_current_solution()[flow.n_u:]
    assert abs(p.max() - 1e4) / 1e4 < 1e-8


def test_body_force_reference_golden(golden_dir):
    """tests/fluid_body_force_mpi/fluid_body_force_mpi.cpp:33-81: 500 steps, p_max - p_min = 1e3 +- 1e-3"""
    import openifem_b200 as ifem

    def body_force(p, c):
        return 1.0e3 / 1.3e-3 if (3.5 - 5e-4 < p[0] < 4.5 + 5e-4 and c == 0) else 0.0

    def sigma_pml(p, c):
        s = 0.0
        for b in (0.0, 8.0):
            if abs(p[0] - b) < 3.0:
                s = 340000 * ((3.0 - abs(p[0] - b)) / 3.0) ** 4
        return s

    tria = ifem.Triangulation(2)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, (160, 30), (0, 0), (8, 2), True)
    flow = ifem.Fluid.MPI.SCnsIM(tria, ifem.Parameters.AllParameters(os.path.join(golden_dir, "scns_body_force_2d.prm")))
    flow.set_body_force(body_force)
    flow.set_sigma_pml_field(sigma_pml)
    flow.run()
    p = flow.get_current_solution()[flow.n_u:]
    assert abs((p.max() - p.min()) - 1e3) / 1e3 < 1e-3


def test_cylinder_flow_scnsim_golden_on_gpu(golden_dir):
    """reference golden tests/fluid_cylinder_mpi_scnsim/fluid_cylinder_mpi_scnsim.cpp:75-85 through the device path:
    SCnsIM Q1/Q1 on GridCreator<2>::flow_around_cylinder refined 3 times, hard-coded inflow (active while t < 2 dt), one
    time step; max velocity 4.5, max pressure 1.03544 to 1e-3"""
    import openifem_b200 as ifem

    tria = ifem.Triangulation(2)
    ifem.GridCreator.flow_around_cylinder(tria)
    params = ifem.Parameters.AllParameters(os.path.join(golden_dir, "scns_cylinder_2d.prm"))
    flow = ifem.Fluid.MPI.SCnsIM(tria, params)
    dt = 1e-2
    flow.add_hard_coded_boundary_condition(
        0, lambda p, c, t: 4 * 4.5 * p[1] * (0.41 - p[1]) / 0.41 ** 2 if c == 0 and abs(p[0]) < 1e-10 and t < 2 * dt else 0.0)
    flow.run()
    sol = flow.get_current_solution()
    vmax, pmax = sol[: flow.n_u].max(), sol[flow.n_u:].max()
    assert abs(vmax - 4.5) / 4.5 < 1e-3, vmax
    assert abs(pmax - 1.03544) / 1.03544 < 1e-3, pmax
