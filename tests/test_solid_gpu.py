"""GPU parity of the hyperelastic hot path: NeoHookean point update, tangent / mass
assembly, Newmark-Newton steps of Solid::MPI::HyperElasticity against the CPU oracle (oracle/solid.py), and
the reference's own beam goldens through the device path.

Tolerances: point history and assembled matrices 1e-12 relative; displacement after time steps 1e-6
relative (linear solves: CG to 1e-8 |b| on the device, sparse direct in the oracle)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _voigt_pairs(dim):
    return [(0, 0), (1, 1), (0, 1)] if dim == 2 else [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]


def _make(golden_dir, dim, reps, hi):
    import openifem_b200 as ifem
    from oracle import fem, prm, solid

    path = os.path.join(golden_dir, f"solid_beam_neohookean_{dim}d.prm")
    o = solid.HyperElasticity(fem.BoxMesh(reps, (0,) * dim, hi), prm.Params(path))
    tria = ifem.Triangulation(dim)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0,) * dim, hi, True)
    g = ifem.Solid.MPI.HyperElasticity(tria, ifem.Parameters.AllParameters(path))
    g.setup()
    return o, g


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


@pytest.mark.parametrize("dim,reps,hi", [(2, (10, 3), (10.0, 1.0)), (3, (6, 2, 3), (10.0, 1.0, 1.2))])
def test_qph_and_assembly_match_oracle(golden_dir, dim, reps, hi):
    o, g = _make(golden_dir, dim, reps, hi)
    rng = np.random.default_rng(11)
    u = 0.05 * rng.uniform(-1, 1, o.n)
    o.cur_u = u.copy()
    o.update_qph(u)
    g.set_vector(g.CUR_U, u)
    g.update_qph()
    Finv, tau, Jc, det = g.get_qph()
    nqp = o.mesh.n_cells * o.nq
    assert _rel(Finv, o.F_inv.reshape(nqp, dim, dim)) < 1e-12
    assert _rel(tau, o.tau.reshape(nqp, dim, dim)) < 1e-12
    assert _rel(det, o.detF.reshape(nqp)) < 1e-13
    pairs = _voigt_pairs(dim)
    Jo = o.Jc.reshape(nqp, dim, dim, dim, dim)
    Jv = np.array([[Jo[:, i, j, k, l] for (k, l) in pairs] for (i, j) in pairs]).transpose(2, 0, 1)
    assert _rel(Jc, Jv) < 1e-12
    for initial in (True, False):
        A_ref, rhs_ref = o.assemble_system(initial)
        g.assemble_system(initial)
        A = g.get_matrix(1 if initial else 0)
        assert sp.linalg.norm(A - A_ref) / sp.linalg.norm(A_ref) < 1e-12
        assert _rel(g.get_vector(g.SYSTEM_RHS), rhs_ref) < 1e-12


@pytest.mark.parametrize("dim,reps,hi", [(2, (20, 2), (10.0, 1.0)), (3, (10, 2, 2), (10.0, 1.0, 1.0))])
def test_time_steps_match_oracle(golden_dir, dim, reps, hi):
    o, g = _make(golden_dir, dim, reps, hi)
    for k in range(4):
        o.run_one_step(k == 0)
        g.run_one_step(k == 0)
    assert _rel(g.get_current_solution(), o.cur_u) < 1e-6
    assert _rel(g.get_vector(g.PREV_V), o.prev_v) < 1e-5
    hg = g.history()
    assert [(h["timestep"], h["iteration"]) for h in hg] == [(h[0], h[1]) for h in o.history]


@pytest.mark.parametrize("dim,reps,hi,umin,umax", [(2, (40, 4), (10.0, 1.0), -0.0616287, 0.00867069),
                                                   (3, (40, 4, 4), (10.0, 1.0, 1.0), -0.0617214, 0.00867507)])
def test_beam_bending_reference_golden_on_gpu(golden_dir, dim, reps, hi, umin, umax):
    """reference tests/solid_beam_bending_mpi_NeoHookean/...cpp:59-68, through run() on the device"""
    import openifem_b200 as ifem

    tria = ifem.Triangulation(dim)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0,) * dim, hi, True)
    s = ifem.Solid.MPI.HyperElasticity(tria, ifem.Parameters.AllParameters(os.path.join(golden_dir, f"solid_beam_neohookean_{dim}d.prm")))
    s.run()
    u = s.get_current_solution()
    assert abs((u.min() - umin) / umin) < 1e-3
    assert abs((u.max() - umax) / umax) < 1e-3


@pytest.mark.parametrize("dim,reps,hi", [(2, (10, 3), (10.0, 1.0)), (3, (6, 2, 3), (10.0, 1.0, 1.2))])
def test_update_strain_and_stress_matches_oracle(golden_dir, dim, reps, hi):
    """SharedHyperElasticity::update_strain_and_stress (source/mpi_shared_hyper_elasticity.cpp:599-714)"""
    o, g = _make(golden_dir, dim, reps, hi)
    rng = np.random.default_rng(12)
    u = 0.05 * rng.uniform(-1, 1, o.n)
    o.update_qph(u)
    g.set_vector(g.CUR_U, u)
    g.update_qph()
    stress, strain = o.update_strain_and_stress()
    g.update_strain_and_stress()
    assert _rel(g.get_nodal_tensor(0), stress) < 1e-12
    assert _rel(g.get_nodal_tensor(1), strain) < 1e-12
