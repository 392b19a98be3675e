"""GPU parity of Solid::MPI::LinearElasticity / SharedLinearElasticity and of FSI::apply_contact_model
against the CPU oracle (oracle/solid.py, oracle/fsi.py), and the reference's own goldens through the device path:
solid_beam_bending_mpi_linearelastic / _shared_linearelastic (u_min = -0.1337) and fsi_contact_model_mpi (u_min = -0.01999).

STATUS: 14 passed on a B200 (profiles/r01f_linear_elasticity_contact_gpu_tests.txt).

Tolerances: assembled matrices / rhs / nodal stress 1e-12 relative; displacement after time steps 1e-7 relative (linear
solves: CG to 1e-8 |b| on the device, sparse direct in the oracle); goldens 1e-3 as in the reference's drivers."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

PRM = """
subsection Simulation
  set Simulation type = {sim}
  set Dimension = {dim}
  set Global refinements = 0, 0
  set End time = 1.0
  set Time step size = 0.05
  set Output interval = 1.0
  set Refinement interval = 100
  set Save interval = 100
  set Gravity = {gravity}
  set Initial velocity = {zeros}
end
subsection Solid finite element system
  set Degree = 1
end
subsection Solid material properties
  set Solid type = LinearElastic
  set Solid density = 3.0
  set Young's modulus = 250
  set Poisson's ratio = 0.3
  set Viscosity = 0.7
end
subsection Solid solver control
  set Damping = 0.1
  set Max Newton iterations = 10
  set Displacement tolerance  = 1.0e-6
  set Force tolerance  = 1.0e-6
end
subsection Solid Dirichlet BCs
  set Number of Dirichlet BCs = 1
  set Dirichlet boundary id = 0
  set Dirichlet boundary components = {full}
end
subsection Solid Neumann BCs
  set Number of Neumann BCs = 1
  set Neumann boundary id = {top}
  set Neumann boundary type = Traction
  set Neumann boundary values = {traction}
end
"""


def _prm(dim, sim="Solid"):
    return PRM.format(sim=sim, dim=dim, gravity=", ".join(["0.3", "-2.0", "0.1"][:dim]), zeros=", ".join(["0.0"] * dim),
                      full=(1 << dim) - 1, top=3, traction=", ".join(["0.2", "-1.5", "0.4"][:dim]))


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def _mrel(A, B):
    return sp.linalg.norm(A - B) / max(sp.linalg.norm(B), 1e-300)


def _make(dim, shared, sim="Solid"):
    import openifem_b200 as ifem
    from oracle import fem, prm, solid

    reps, hi = ((7, 3), (4.0, 1.0)) if dim == 2 else ((5, 2, 3), (4.0, 1.0, 1.2))
    text = _prm(dim, sim)
    o = solid.LinearElasticity(fem.BoxMesh(reps, (0,) * dim, hi), prm.Params(text, is_text=True), shared=shared)
    tria = ifem.Triangulation(dim)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0,) * dim, hi, True)
    cls = ifem.Solid.MPI.SharedLinearElasticity if shared else ifem.Solid.MPI.LinearElasticity
    g = cls(tria, ifem.Parameters.AllParameters(text=text))
    g.setup()
    return o, g


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("shared", [False, True])
def test_linear_assembly_matches_oracle(dim, shared):
    """mpi_linear_elasticity.cpp:27-198 / mpi_shared_linear_elasticity.cpp:26-299: every matrix the twin assembles and the
    right-hand side (gravity + traction faces), 1e-12"""
    o, g = _make(dim, shared)
    assert g.n_dofs == o.n
    o.assemble_system(True)
    g.assemble_system(True)
    assert _mrel(g.get_matrix(g.SYSTEM), o.system_matrix) < 1e-12
    assert _rel(g.get_vector(g.SYSTEM_RHS), o.system_rhs) < 1e-12
    if shared:
        assert _mrel(g.get_matrix(g.MASS), o.mass_matrix) < 1e-12
        assert _mrel(g.get_matrix(g.STIFFNESS), o.stiffness_matrix) < 1e-12
        assert _mrel(g.get_matrix(g.DAMPING), o.damping_matrix) < 1e-12
    o.assemble_system(False)
    g.assemble_system(False)
    assert _rel(g.get_vector(g.SYSTEM_RHS), o.system_rhs) < 1e-12
    if not shared:
        assert _mrel(g.get_matrix(g.SYSTEM), o.system_matrix) < 1e-12
        assert _mrel(g.get_matrix(g.STIFFNESS), o.stiffness_matrix) < 1e-12


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("shared", [False, True])
def test_linear_time_steps_match_oracle(dim, shared):
    """run_one_step (mpi_linear_elasticity.cpp:199-262 / mpi_shared_linear_elasticity.cpp:300-398): five Newmark steps"""
    o, g = _make(dim, shared)
    for k in range(5):
        o.run_one_step(k == 0)
        g.run_one_step(k == 0)
    assert _rel(g.get_vector(g.CUR_U), o.cur_u) < 1e-7
    assert _rel(g.get_vector(g.CUR_V), o.cur_v) < 1e-7
    assert _rel(g.get_vector(g.CUR_A), o.cur_a) < 1e-6
    if shared:
        # update_strain_and_stress ran inside the step (mpi_shared_linear_elasticity.cpp:377)
        assert _rel(g.get_nodal_tensor(0), o.stress) < 1e-6
        assert _rel(g.get_nodal_tensor(1), o.strain) < 1e-6


@pytest.mark.parametrize("dim", [2, 3])
def test_linear_strain_and_stress_match_oracle(dim):
    """SharedLinearElasticity::update_strain_and_stress (mpi_shared_linear_elasticity.cpp:401-531) on a random displacement"""
    o, g = _make(dim, True)
    rng = np.random.default_rng(5)
    u = 0.01 * rng.uniform(-1, 1, o.n)
    o.cur_u = u.copy()
    g.set_vector(g.CUR_U, u)
    stress, strain = o.update_strain_and_stress()
    g.update_strain_and_stress()
    assert _rel(g.get_nodal_tensor(0), stress) < 1e-12
    assert _rel(g.get_nodal_tensor(1), strain) < 1e-12


@pytest.mark.parametrize("shared", [False, True])
def test_beam_linearelastic_reference_golden(golden_dir, shared):
    """reference goldens tests/solid_beam_bending_mpi_linearelastic/...cpp:50-53 and ..._shared_linearelastic/...cpp:50-53:
    u_min = -0.1337 to 1e-3 after 200 steps, through run() (which refines the 32 x 4 mesh once)"""
    import openifem_b200 as ifem

    tria = ifem.Triangulation(2)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, (32, 4), (0, 0), (8.0, 1.0), True)
    cls = ifem.Solid.MPI.SharedLinearElasticity if shared else ifem.Solid.MPI.LinearElasticity
    s = cls(tria, ifem.Parameters.AllParameters(os.path.join(golden_dir, "solid_beam_linearelastic_2d.prm")))
    s.run()
    assert tria.n_active_cells() == 64 * 8
    umin = s.get_current_solution().min()
    assert abs(umin + 0.1337) / 0.1337 < 1e-3, umin
    assert abs(umin + 0.13370340734763897) < 1e-6, umin  # the oracle's value


def _contact_pair(golden_dir):
    import openifem_b200 as ifem

    path = os.path.join(golden_dir, "fsi_contact_model_2d.prm")
    params = ifem.Parameters.AllParameters(path)
    tf, ts = ifem.Triangulation(2), ifem.Triangulation(2)
    ifem.GridGenerator.subdivided_hyper_rectangle(tf, (50, 25), (0, 0), (2.0, 1.0), True)
    ifem.GridGenerator.subdivided_hyper_rectangle(ts, (10, 11), (0.25, 0.0), (1.25, 1.02), True)
    fluid = ifem.Fluid.MPI.SCnsIM(tf, params)
    solid = ifem.Solid.MPI.SharedLinearElasticity(ts, params)
    fluid.setup()
    solid.setup()
    fsi = ifem.MPI.FSI(fluid, solid, params)
    fsi.set_penetration_criterion(lambda p: p[1] - 1.0, [0.0, -1.0])
    return fluid, solid, fsi


def test_fsi_contact_model_reference_golden(golden_dir):
    """reference golden tests/fsi_contact_model_mpi/fsi_contact_model_mpi.cpp:46-60 through the device path: one coupled
    step of SCnsIM + SharedLinearElasticity with apply_contact_model; solid u_min = -0.01999 to 1e-3, and the oracle's
    38 contact iterations / displacement field"""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_oracle_goldens import contact_problem

    fluid, solid, fsi = _contact_pair(golden_dir)
    fsi.run()
    u = solid.get_current_solution()
    assert abs(u.min() + 0.01999) / 0.01999 < 1e-3, u.min()
    c = contact_problem(golden_dir)
    c.run()
    assert fsi.contact_iterations() == c.contact_iterations == 38
    assert _rel(u, c.solid.cur_u) < 1e-6
    # the fluid step that follows the contact loop (FGMRES to the reference's 1e-6 |rhs| on the device, sparse direct in the
    # oracle): velocity to 1e-4
    fsol = fluid.get_current_solution()
    assert _rel(fsol[: c.fluid.n_u], c.fluid.present[: c.fluid.n_u]) < 1e-4


def test_cpp_contact_driver_reference_golden(golden_dir):
    """the reference's own driver body (tests/cpp/fsi_contact_model_mpi.cpp) compiled against the C++ facade"""
    if os.environ.get("IFEM_CPU_EMULATION"):
        pytest.skip("compiled drivers link the product library: not replayable on the emulated device")
    import subprocess
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_cpp_facade import ROOT, _build

    _build("fsi_contact_model_mpi")
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "fsi_contact_model_mpi")
    r = subprocess.run([exe, os.path.join(golden_dir, "fsi_contact_model_2d.prm")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
