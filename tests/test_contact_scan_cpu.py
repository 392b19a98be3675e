"""The host function the device path uses for the penetration scan of FSI::apply_contact_model (openifem_b200/csrc/contact.h;
reference source/mpi_fsi.cpp:897-956) against the oracle's restatement (oracle/fsi.py FSI.contact_scan, pinned on the
fsi_contact_model_mpi golden), on the CPU: same mesh, same displaced solid, same criterion -> same still_penetrate flag and the
same increments of fsi_stress_rows."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "openifem_b200", "csrc")
POINT_FN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double))


@pytest.fixture(scope="module")
def harness():
    out = os.path.join(ROOT, "tests", "cpp", "_build", "libcontact_scan_cpu.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    srcs = [os.path.join(ROOT, "tests", "cpp", "contact_scan_cpu.cpp"), os.path.join(CSRC, "mesh.cpp"), os.path.join(CSRC, "fe_tables.cpp")]
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-fopenmp"] + srcs + ["-o", out])
    return C.CDLL(out)


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def _scan_both(harness, c, criterion, direction):
    s = c.solid
    dim = s.dim
    rows = np.ascontiguousarray(s.fsi_stress_rows.copy())
    cb = POINT_FN(lambda p: float(criterion([p[i] for i in range(dim)])))
    d = np.asarray(direction, dtype=float)
    bf = np.ascontiguousarray(s.mesh.boundary_faces, dtype=np.int32)
    nodes = np.ascontiguousarray(s.dofs.nodes, dtype=np.int32)
    coords = np.ascontiguousarray(s.dofs.coords, dtype=float)
    u = np.ascontiguousarray(s.cur_u)
    got = harness.cpu_contact_scan(C.c_int(dim), C.c_int(1), C.c_int(bf.shape[0]), _p(bf, C.c_int), _p(nodes, C.c_int), C.c_int(nodes.shape[1]),
                                   _p(coords), _p(u), C.c_longlong(s.n), cb, _p(d), C.c_double(s.prm.contact_force_multiplier), _p(rows))
    c.set_penetration_criterion(criterion, direction)
    ref = c.contact_scan()
    return bool(got), rows, ref, s.fsi_stress_rows


def test_contact_scan_matches_oracle_on_the_reference_case(harness, golden_dir):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_oracle_goldens import contact_problem

    c = contact_problem(golden_dir)
    s = c.solid
    # after one solid step from rest (no load yet) nothing has moved: the top row penetrates by 0.02
    s.run_one_step(True)
    for k in range(3):
        got, rows, ref, rows_ref = _scan_both(harness, c, lambda p: p[1] - 1.0, [0.0, -1.0])
        assert got is True and ref is True
        assert np.abs(rows).max() > 0
        assert np.linalg.norm(rows - rows_ref) <= 1e-14 * np.linalg.norm(rows_ref)
        # redo the step with the accumulated contact stress, as apply_contact_model does
        for v in (s.cur_a, s.cur_v, s.cur_u, s.prev_a, s.prev_v, s.prev_u):
            v[:] = 0.0
        s.time, s.timestep = 0.0, 0
        s.run_one_step(True)


@pytest.mark.parametrize("dim", [2, 3])
def test_contact_scan_matches_oracle_on_a_tilted_solid(harness, dim):
    """random displacement (tilted faces, every normal component in play), oblique direction, 2-D and 3-D"""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import test_linear_elasticity_gpu as T
    from oracle import fem, fsi, prm, solid

    reps, hi = ((5, 4), (1.0, 1.0)) if dim == 2 else ((3, 2, 3), (1.0, 0.8, 1.0))
    p = prm.Params(T._prm(dim, "FSI"), is_text=True)
    s = solid.LinearElasticity(fem.BoxMesh(reps, (0,) * dim, hi), p, shared=True)
    s.cur_u = 0.03 * np.random.default_rng(9).uniform(-1, 1, s.n)

    class _F:  # the scan touches the solid only
        con = np.zeros(1)
        nonzero_val = np.zeros(1)

    c = fsi.FSI(_F(), s)
    direction = [0.3, -1.0, 0.2][:dim]
    crit = lambda pt: pt[dim - 1] - 0.97 + 0.05 * pt[0]
    got, rows, ref, rows_ref = _scan_both(harness, c, crit, direction)
    assert got is True and ref is True
    assert np.count_nonzero(rows_ref) > 3
    assert np.linalg.norm(rows - rows_ref) <= 1e-13 * np.linalg.norm(rows_ref)
