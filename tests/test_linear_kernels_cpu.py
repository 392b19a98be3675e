"""The kernel bodies of the linear-elasticity device path (openifem_b200/csrc/solid_linear.cuh) run on the CPU: the
same source the CUDA kernels wrap is compiled with g++ (tests/cpp/linear_kernels_cpu.cpp) and walked over its launch grid
sequentially, then compared with the oracle (oracle/solid.py, pinned on the reference's linear-elastic beam and contact
goldens). This checks the device arithmetic, the BCSR row-plane indexing and the constrained scatter without a GPU; the
launches themselves are covered by the gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import fem, prm, solid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness():
    out = os.path.join(ROOT, "tests", "cpp", "_build", "liblinear_kernels_cpu.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "linear_kernels_cpu.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-Wno-unknown-pragmas", src, "-o", out])
    return C.CDLL(out)


def _prm_text(dim, dirichlet=True):
    import test_linear_elasticity_gpu as T

    text = T._prm(dim).replace("Number of Neumann BCs = 1", "Number of Neumann BCs = 0")
    return text if dirichlet else text.replace("Number of Dirichlet BCs = 1", "Number of Dirichlet BCs = 0")


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def _node_pattern(nodes, n_nodes):
    rows = np.repeat(nodes, nodes.shape[1], axis=1).ravel()
    cols = np.tile(nodes, (1, nodes.shape[1])).ravel()
    P = sp.coo_matrix((np.ones(rows.size), (rows, cols)), shape=(n_nodes, n_nodes)).tocsr()
    P.sort_indices()
    return P.indptr.astype(np.int64), P.indices.astype(np.int32)


def _slots(nodes, rowptr, col):
    nc, npc = nodes.shape
    s = np.zeros((nc, npc, npc), dtype=np.uint8)
    for c in range(nc):
        for a in range(npc):
            r = nodes[c, a]
            lst = col[rowptr[r]: rowptr[r + 1]]
            for b in range(npc):
                s[c, a, b] = np.searchsorted(lst, nodes[c, b])
    return s


def _to_csr(val, rowptr, col, dim):
    """row-plane BCSR (val[rowptr[i] d d + (r d + c) nb + j]) -> scalar CSR"""
    rows, cols, vals = [], [], []
    for i in range(rowptr.size - 1):
        nb = rowptr[i + 1] - rowptr[i]
        blk = val[rowptr[i] * dim * dim: rowptr[i + 1] * dim * dim].reshape(dim, dim, nb)
        cj = col[rowptr[i]: rowptr[i + 1]]
        for r in range(dim):
            for c in range(dim):
                rows.append(np.full(nb, i * dim + r)), cols.append(cj * dim + c), vals.append(blk[r, c])
    n = (rowptr.size - 1) * dim
    return sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()


def _mrel(A, B):
    return sp.linalg.norm(A - B) / max(sp.linalg.norm(B), 1e-300)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("shared", [False, True])
def test_linear_assemble_body_matches_oracle(harness, dim, shared):
    reps, hi = ((7, 3), (4.0, 1.0)) if dim == 2 else ((4, 2, 3), (4.0, 1.0, 1.2))
    p = prm.Params(_prm_text(dim), is_text=True)
    o = solid.LinearElasticity(fem.BoxMesh(reps, (0,) * dim, hi), p, shared=shared)
    nodes = np.ascontiguousarray(o.dofs.nodes, dtype=np.int32)
    n_nodes = o.dofs.n_nodes
    rowptr, col = _node_pattern(nodes, n_nodes)
    slots = _slots(nodes, rowptr, col)
    nnzb = col.size
    E, nu = p.E[0], p.nu[0]
    lam, mu, eta = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu)), p.eta[0]
    dt = p.time_step
    N, G, JxW = np.ascontiguousarray(o.N), np.ascontiguousarray(o.G), np.ascontiguousarray(o.JxW)
    grav = np.asarray(p.gravity[:dim], dtype=float)

    def run(c_mass, c_damp, c_stiff, want):
        bufs = {k: np.zeros(nnzb * dim * dim) if k in want else None for k in ("sys", "mass", "stiff", "damp")}
        rhs = np.zeros(o.n)
        rc = harness.cpu_linear_assemble(
            C.c_int(dim), C.c_int(nodes.shape[0]), _p(nodes, C.c_int), _p(slots, C.c_ubyte), _p(o.con, C.c_ubyte), _p(N), _p(G), _p(JxW),
            C.c_int(o.nq), C.c_double(p.solid_rho), C.c_double(lam), C.c_double(mu), C.c_double(eta), _p(grav), C.c_double(c_mass),
            C.c_double(c_damp), C.c_double(c_stiff), _p(rowptr, C.c_int64),
            *[(_p(bufs[k]) if bufs[k] is not None else None) for k in ("sys", "mass", "stiff", "damp")], _p(rhs))
        assert rc == 0
        return {k: (_to_csr(v, rowptr, col, dim) if v is not None else None) for k, v in bufs.items()}, rhs

    if shared:
        # the coefficients LinearElasticity::assemble_system passes for the shared twin (mpi_shared_linear_elasticity.cpp:30-32)
        alpha = -p.damping
        gamma, beta = 0.5 - alpha, (1 + alpha) ** 2 / 4
        got, rhs = run(1.0, gamma * dt * (1 + alpha), beta * dt * dt * (1 + alpha), ("sys", "mass", "stiff", "damp"))
        o.assemble_system(True)
        assert _mrel(got["sys"], o.system_matrix) < 1e-13
        assert _mrel(got["mass"], o.mass_matrix) < 1e-13
        assert _mrel(got["stiff"], o.stiffness_matrix) < 1e-13
        assert _mrel(got["damp"], o.damping_matrix) < 1e-13
        assert np.linalg.norm(rhs - o.system_rhs) / np.linalg.norm(o.system_rhs) < 1e-13
        # rhs-only pass (every later FSI step)
        got, rhs = run(1.0, 0.0, 0.0, ())
        assert np.linalg.norm(rhs - o.system_rhs) / np.linalg.norm(o.system_rhs) < 1e-13
    else:
        gamma = 0.5 + p.damping
        beta = gamma / 2
        got, rhs = run(1.0, 0.0, 0.0, ("sys",))
        o.assemble_system(True)
        assert _mrel(got["sys"], o.system_matrix) < 1e-13
        got, rhs = run(1.0, 0.0, beta * dt * dt, ("sys", "stiff"))
        o.assemble_system(False)
        assert _mrel(got["sys"], o.system_matrix) < 1e-13
        assert _mrel(got["stiff"], o.stiffness_matrix) < 1e-13
        assert np.linalg.norm(rhs - o.system_rhs) / np.linalg.norm(o.system_rhs) < 1e-13


@pytest.mark.parametrize("dim", [2, 3])
def test_linear_stress_body_matches_oracle(harness, dim):
    reps, hi = ((7, 3), (4.0, 1.0)) if dim == 2 else ((4, 2, 3), (4.0, 1.0, 1.2))
    p = prm.Params(_prm_text(dim), is_text=True)
    o = solid.LinearElasticity(fem.BoxMesh(reps, (0,) * dim, hi), p, shared=True)
    rng = np.random.default_rng(3)
    u = 0.01 * rng.uniform(-1, 1, o.n)
    o.cur_u = u.copy()
    stress_ref, strain_ref = o.update_strain_and_stress()
    nodes = np.ascontiguousarray(o.dofs.nodes, dtype=np.int32)
    n_nodes = o.dofs.n_nodes
    Mref = np.einsum("qi,qj,q->ij", o.N, o.N, o.qw)
    qpt_to_dof = np.ascontiguousarray(np.linalg.solve(Mref, (o.N * o.qw[:, None]).T))
    E, nu = p.E[0], p.nu[0]
    lam, mu = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    stress, strain, count = np.zeros((dim * dim, n_nodes)), np.zeros((dim * dim, n_nodes)), np.zeros(n_nodes)
    G = np.ascontiguousarray(o.G)
    rc = harness.cpu_linear_stress(C.c_int(dim), C.c_int(nodes.shape[0]), C.c_int(o.nq), _p(nodes, C.c_int), _p(qpt_to_dof), _p(G), _p(u),
                                   C.c_double(lam), C.c_double(mu), C.c_int(n_nodes), _p(stress), _p(strain), _p(count))
    assert rc == 0
    assert np.linalg.norm(stress / count - stress_ref) / np.linalg.norm(stress_ref) < 1e-13
    assert np.linalg.norm(strain / count - strain_ref) / np.linalg.norm(strain_ref) < 1e-13
