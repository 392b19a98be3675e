"""The OpenMP Krylov loops of the oracle (oracle/csrc/oracle_krylov.cpp, used by bench.py's CPU baseline) against the plain
numpy statement of the same algorithms in oracle/ins.py."""
import numpy as np
import scipy.sparse as sp

from oracle import ins


def _spd(n, seed):
    rng = np.random.default_rng(seed)
    A = sp.random(n, n, density=8.0 / n, random_state=seed, format="csr")
    A = (A + A.T) * 0.5 + sp.diags(np.full(n, 4.0))
    return A.tocsr(), rng.uniform(-1, 1, n)


def test_cg_c_equals_numpy():
    A, b = _spd(3000, 1)
    op = ins.CsrOp(A)
    x1, it1, r1 = ins.cg(op, b, np.zeros_like(b), 1e-9, 500)
    x2, it2, r2 = ins.cg_py(op, b, np.zeros_like(b), 1e-9, 500)
    assert it1 == it2
    assert np.linalg.norm(x1 - x2) <= 1e-12 * np.linalg.norm(x2)
    # nonzero initial guess
    x0 = 0.5 * x2
    x3, it3, _ = ins.cg(op, b, x0, 1e-9, 500)
    x4, it4, _ = ins.cg_py(op, b, x0, 1e-9, 500)
    assert it3 == it4 and np.linalg.norm(x3 - x4) <= 1e-12 * np.linalg.norm(x4)


def test_bicgstab_c_equals_numpy():
    n_nodes, bs = 1200, 3
    A, b = _spd(n_nodes * bs, 2)
    A = (A + sp.random(A.shape[0], A.shape[0], density=2.0 / A.shape[0], random_state=5, format="csr")).tocsr()  # nonsymmetric
    blocks = np.stack([A[i * bs:(i + 1) * bs, i * bs:(i + 1) * bs].toarray() for i in range(n_nodes)])
    prec = ins.BlockJacobi(np.linalg.inv(blocks))
    op = ins.CsrOp(A)
    tol = 1e-8 * np.linalg.norm(b)
    x1, it1, _ = ins.bicgstab(op, prec, b, tol, 300)
    x2, it2, _ = ins.bicgstab_py(op, prec, b, tol, 300)
    assert it1 == it2
    assert np.linalg.norm(x1 - x2) <= 1e-10 * np.linalg.norm(x2)
    assert np.linalg.norm(A @ x1 - b) <= 10 * tol


def test_csr_split_blocks():
    A, _ = _spd(500, 3)
    nu = 380
    uu, up, pu, pp = ins.csr_split(A, nu)
    assert abs(uu - A[:nu, :nu]).max() == 0 and abs(up - A[:nu, nu:]).max() == 0
    assert abs(pu - A[nu:, :nu]).max() == 0 and abs(pp - A[nu:, nu:]).max() == 0


def test_thread_team_is_settable():
    assert ins.set_threads(2) == 2
    assert ins.set_threads(None) >= 1
