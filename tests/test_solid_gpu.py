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


def test_two_solid_parts_pick_the_material_of_the_cell(golden_dir, tmp_path):
    """`Number of solid parts = 2`: every cell takes the (C1, kappa) of its material id (mpi_hyper_elasticity.cpp:226-228,
    8-20) - the point history and the tangent matrix must follow the oracle with the same per-cell parts."""
    import openifem_b200 as ifem
    from oracle import fem, prm, solid

    dim, reps, hi = 2, (8, 3), (8.0, 1.0)
    text = open(os.path.join(golden_dir, "solid_beam_neohookean_2d.prm")).read()
    c = prm.Params(os.path.join(golden_dir, "solid_beam_neohookean_2d.prm")).C[0]
    import re
    text, n = re.subn(r"set Hyperelastic parameters\s*=.*", f"set Hyperelastic parameters = {c[0]}, {c[1]}, {3.0 * c[0]}, {0.5 * c[1]}\n"
                      "  set Number of solid parts = 2\n  set Viscosity = 0, 0", text)
    assert n == 1
    for key in ("Young's modulus", "Poisson's ratio"):
        text, n = re.subn(rf"(set {key}\s*=\s*)([^\n,]+)\n", r"\1\2, \2\n", text)
        assert n == 1
    path = tmp_path / "two_parts.prm"
    path.write_text(text)
    mesh = fem.BoxMesh(reps, (0,) * dim, hi)
    ids = np.where(mesh.vertices[mesh.cells].mean(axis=1)[:, 0] < 4.0, 1, 2).astype(np.int32)
    o = solid.HyperElasticity(mesh, prm.Params(str(path)), material_id=ids)
    tria = ifem.Triangulation(dim)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0,) * dim, hi, True)
    tria.set_material_ids(ids)
    g = ifem.Solid.MPI.HyperElasticity(tria, ifem.Parameters.AllParameters(str(path)))
    g.setup()
    u = 0.05 * np.random.default_rng(5).uniform(-1, 1, o.n)
    o.cur_u = u.copy()
    o.update_qph(u)
    g.set_vector(g.CUR_U, u)
    g.update_qph()
    _, tau, _, _ = g.get_qph()
    assert _rel(tau, o.tau.reshape(-1, dim, dim)) < 1e-12
    A_ref, rhs_ref = o.assemble_system(False)
    g.assemble_system(False)
    assert sp.linalg.norm(g.get_matrix(0) - A_ref) / sp.linalg.norm(A_ref) < 1e-12
    assert _rel(g.get_vector(g.SYSTEM_RHS), rhs_ref) < 1e-12
    # and the two parts really differ: the same state with one part gives another stress
    o1 = solid.HyperElasticity(mesh, prm.Params(os.path.join(golden_dir, "solid_beam_neohookean_2d.prm")))
    o1.update_qph(u)
    assert _rel(o1.tau, o.tau) > 1e-2
