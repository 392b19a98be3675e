"""SCnsIM with element pairs other than Q1/Q1 (the reference's SCnsIM::assemble, source/mpi_scnsim.cpp:15-568, is degree-generic;
its cases all use Q1/Q1): the degree-generic device kernel (csrc/scnsim_generic.cu) against the oracle's degree-generic cell loop
(oracle/csrc/oracle_scns.cpp) for Taylor-Hood Q2/Q1 and equal-order Q2/Q2 with every term switched on.
Tolerances: assembled matrix / rhs 1e-12 relative, fields after time steps 1e-6 (device FGMRES to 1e-10 |rhs|, oracle direct)."""
import numpy as np
import pytest
import scipy.sparse as sp

from test_scns_gpu import BF, SIG, scns_prm

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def _make(dim, pu, pp, reps, hi, **kw):
    import openifem_b200 as ifem
    from oracle import fem, prm, scns

    text = scns_prm(dim, **kw).replace("set Pressure degree = 1", f"set Pressure degree = {pp}").replace("set Velocity degree = 1", f"set Velocity degree = {pu}")
    o = scns.SCnsIM(fem.BoxMesh(reps, (0,) * dim, hi), prm.Params(text, is_text=True), body_force=BF, sigma_pml_field=SIG)
    tria = ifem.Triangulation(dim)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0,) * dim, hi, True)
    g = ifem.Fluid.MPI.SCnsIM(tria, ifem.Parameters.AllParameters(text=text))
    g.set_body_force(BF)
    g.set_sigma_pml_field(SIG)
    g.setup()
    assert g.n_dofs == o.n
    return o, g


@pytest.mark.parametrize("dim,pu,pp,reps,hi", [(2, 2, 1, (4, 3), (1.0, 0.8)), (2, 2, 2, (3, 3), (1.0, 0.8)), (3, 2, 1, (2, 3, 2), (1.0, 1.2, 0.9))])
@pytest.mark.parametrize("nonzero", [True, False])
def test_generic_degree_assembly_matches_oracle(dim, pu, pp, reps, hi, nonzero):
    o, g = _make(dim, pu, pp, reps, hi, gravity=[1.0, -9.8, 0.5][:dim], neumann={1: 3.5})
    rng = np.random.default_rng(3)
    ev, pr, acc = rng.uniform(-1, 1, o.n), rng.uniform(-1, 1, o.n), rng.uniform(-1, 1, o.n)
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
    A_ref, b_ref = o.assemble(nonzero)
    g.assemble(nonzero)
    A, b = g.get_matrix(0), g.get_vector(g.SYSTEM_RHS)
    err_A = sp.linalg.norm(A - A_ref) / sp.linalg.norm(A_ref)
    assert err_A < 1e-12 and _rel(b, b_ref) < 1e-12, (err_A, _rel(b, b_ref))


@pytest.mark.parametrize("dim,pu,pp,reps,hi", [(2, 2, 1, (5, 4), (1.0, 0.8)), (3, 2, 1, (2, 2, 3), (1.0, 1.0, 1.2))])
def test_generic_degree_time_steps_match_oracle(dim, pu, pp, reps, hi):
    o, g = _make(dim, pu, pp, reps, hi, dt=1e-3, newton_tol=1e-9)
    g.set_control(fgmres_rel=1e-10)
    for k in range(2):
        o.run_one_step(k == 0)
        g.run_one_step(k == 0)
    sol = g.get_current_solution()
    assert _rel(sol[: o.n_u], o.velocity()) < 1e-6 and _rel(sol[o.n_u:], o.pressure()) < 1e-6
    assert _rel(g.get_stress(), o.stress) < 1e-6
