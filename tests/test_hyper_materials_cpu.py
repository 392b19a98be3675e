"""The Kirchhoff material point of the hyperelastic device path (openifem_b200/csrc/hyper_materials.cuh: the function
update_qph_kernel calls for solid_type = Kirchhoff) compiled with g++ and compared with the oracle (oracle/solid.py
kirchhoff_update, restated from the reference's include/kirchhoff_elastic_material.h:36-76), plus properties of the model
itself: zero stress under rigid rotation (what tests/solid_rotation_mpi_shared_Kirchhoff exercises), linear elasticity in the
small-strain limit, and a full oracle run of that rotation case staying rigid."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import solid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness():
    out = os.path.join(ROOT, "tests", "cpp", "_build", "libhyper_materials_cpu.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                           os.path.join(ROOT, "tests", "cpp", "hyper_materials_cpu.cpp"), "-o", out])
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _pairs(dim):
    return [(0, 0), (1, 1), (0, 1)] if dim == 2 else [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]


def _device_points(h, grad_u, young, poisson):
    n, dim = grad_u.shape[0], grad_u.shape[-1]
    ns = dim * (dim + 1) // 2
    Finv, tau, Jc, det = np.empty((n, dim, dim)), np.empty((n, dim, dim)), np.empty((n, ns, ns)), np.empty(n)
    assert h.cpu_kirchhoff_points(C.c_int(dim), C.c_int(n), _p(np.ascontiguousarray(grad_u)), C.c_double(young), C.c_double(poisson),
                                  _p(Finv), _p(tau), _p(Jc), _p(det)) == 0
    return Finv, tau, Jc, det


@pytest.mark.parametrize("dim", [2, 3])
def test_kirchhoff_point_matches_oracle(harness, dim):
    rng = np.random.default_rng(dim)
    grad_u = 0.3 * rng.uniform(-1, 1, (200, dim, dim))
    Finv, tau, Jc, det = _device_points(harness, grad_u, 250.0, 0.3)
    Fo, to, Jo, do = solid.kirchhoff_update(grad_u, 250.0, 0.3)
    pairs = _pairs(dim)
    Jv = np.array([[Jo[:, i, j, k, l] for (k, l) in pairs] for (i, j) in pairs]).transpose(2, 0, 1)
    for a, b in ((Finv, Fo), (tau, to), (Jc, Jv), (det, do)):
        assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-14


@pytest.mark.parametrize("dim", [2, 3])
def test_kirchhoff_model_properties(harness, dim):
    # rigid rotation: F = R, E = 0, no stress
    th = 0.7
    R = np.eye(dim)
    R[:2, :2] = [[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]
    _, tau, _, det = _device_points(harness, (R - np.eye(dim))[None], 250.0, 0.3)
    assert np.abs(tau).max() < 1e-12 and abs(det[0] - 1) < 1e-14
    # small strain: tau -> lambda tr(eps) I + 2 mu eps
    rng = np.random.default_rng(7)
    g = 1e-7 * rng.uniform(-1, 1, (dim, dim))
    eps = 0.5 * (g + g.T)
    lam, mu = 250.0 * 0.3 / (1.3 * 0.4), 250.0 / 2.6
    _, tau, _, _ = _device_points(harness, g[None], 250.0, 0.3)
    assert np.abs(tau[0] - (lam * np.trace(eps) * np.eye(dim) + 2 * mu * eps)).max() < 1e-11


def test_kirchhoff_rotation_case_conserves_momentum(golden_dir):
    """the reference's solid_rotation_mpi_shared_Kirchhoff case (free unit square, traction (0, 1e4) on the face x = 0, 1e-4 s
    steps; the reference only checks that it runs) on the oracle, 4 x coarser and for 20 steps: internal forces cancel, so the
    centre of mass moves with F / m exactly - u_cm = (0, F t^2 / 2m) - while the body rotates and deforms"""
    from oracle import fem, prm

    text = open(os.path.join(golden_dir, "solid_rotation_kirchhoff_2d.prm")).read().replace("set End time = 5e-2", "set End time = 2e-3")
    o = solid.HyperElasticity(fem.BoxMesh((2, 2), (0, 0), (1.0, 1.0)).refine_global(2), prm.Params(text, is_text=True))
    o.run()
    M, _ = o.assemble_system(True)
    ex, ey = np.zeros(o.n), np.zeros(o.n)
    ex[0::2], ey[1::2] = 1.0, 1.0
    mass = ey @ (M @ ey)
    assert abs(mass - 1.0) < 1e-12
    assert abs(ey @ (M @ o.cur_u) / mass - 0.5 * 1e4 * (2e-3) ** 2) < 1e-12
    assert abs(ex @ (M @ o.cur_u) / mass) < 1e-12
    assert o.cur_u.max() - o.cur_u.min() > 0.05  # it does rotate
