"""ORACLE (test infrastructure, NOT product code).

CPU restatement of Fluid::MPI::SUPGFluidSolver<dim>::BlockIncompSchurPreconditioner (reference source/mpi_supg_solver.cpp:35-192)
and of SUPGFluidSolver::solve with it (:297-328): FGMRES to 1e-6 |rhs| with
    P^-1 = [Pvv^-1  -Pvv^-1 Avp Tpp^-1; 0  Tpp^-1] [I 0; -Apv Pvv^-1  I],   Tpp = App - Apv Pvv^-1 Avp (matrix-free, :20-32),
Pvv^-1 = ILU(0) of Avv (:51), Tpp^-1 by an inner GMRES(200) to 1e-3 |ptmp| preconditioned by an ILU(0) of
B2pp = App - Apv diag(rowsum|Avv|)^-1 Avp (:68-133) from the one-step line-search initial guess (:163-169).

The reference's ILU(0) factors are Hypre Euclid's (source/preconditioner_pilut.cpp:124-138; Hypre is not vendored, "parity
unpinned"; on one rank Euclid's ILU(0) is the textbook IKJ ILU(0) restated here). Used to compare ITERATION COUNTS of the device
preconditioner with the reference's algorithm - the converged fields are compared against the sparse direct solve as before."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def ilu0(A):
    """textbook IKJ ILU(0) on the pattern of A; returns (L unit lower, U upper) as CSR"""
    A = A.tocsr().copy()
    A.sort_indices()
    rp, ci, v = A.indptr, A.indices, A.data.copy()
    n = A.shape[0]
    diag = np.array([rp[i] + int(np.searchsorted(ci[rp[i]:rp[i + 1]], i)) for i in range(n)])
    for i in range(n):
        pos = {int(ci[p]): p for p in range(rp[i], rp[i + 1])}
        for p in range(rp[i], diag[i]):
            k = int(ci[p])
            v[p] /= v[diag[k]]
            for q in range(diag[k] + 1, rp[k + 1]):
                t = pos.get(int(ci[q]))
                if t is not None:
                    v[t] -= v[p] * v[q]
    LU = sp.csr_matrix((v, ci, rp), shape=A.shape)
    L = (sp.tril(LU, -1) + sp.identity(n)).tocsr()
    U = sp.triu(LU, 0).tocsr()
    return L, U


class Ilu0Inverse:
    def __init__(self, A):
        self.L, self.U = ilu0(A)

    def __call__(self, b):
        return spla.spsolve_triangular(self.U, spla.spsolve_triangular(self.L, b, lower=True), lower=False)


def gmres_left(A, M, b, x0, tol_abs, max_it, restart=200):
    """deal.II SolverGMRES with a left preconditioner (the default): minimises and tests |M (b - A x)|. Returns (x, iterations)."""
    x = x0.copy()
    its = 0
    while True:
        r = M(b - A(x))
        beta = np.linalg.norm(r)
        if beta <= tol_abs or its >= max_it:
            return x, its
        V = [r / beta]
        H = np.zeros((restart + 1, restart))
        y = np.zeros(0)
        done = False
        for j in range(restart):
            w = M(A(V[j]))
            for i in range(j + 1):
                H[i, j] = w @ V[i]
                w = w - H[i, j] * V[i]
            H[j + 1, j] = np.linalg.norm(w)
            rhs = np.zeros(j + 2)
            rhs[0] = beta
            y, *_ = np.linalg.lstsq(H[: j + 2, : j + 1], rhs, rcond=None)
            res = np.linalg.norm(rhs - H[: j + 2, : j + 1] @ y)
            its += 1
            if res <= tol_abs or its >= max_it or H[j + 1, j] == 0.0:
                done = True
                break
            V.append(w / H[j + 1, j])
        for i in range(y.size):
            x = x + y[i] * V[i]
        if done:
            return x, its


class BlockIncompSchurPreconditioner:
    def __init__(self, system_matrix, n_u):
        A = system_matrix.tocsr()
        self.n_u = n_u
        self.Avv, self.Avp = A[:n_u, :n_u].tocsr(), A[:n_u, n_u:].tocsr()
        self.Apv, self.App = A[n_u:, :n_u].tocsr(), A[n_u:, n_u:].tocsr()
        self.Pvv = Ilu0Inverse(self.Avv)
        rowsum = np.asarray(abs(self.Avv).sum(axis=1)).ravel()
        B2pp = (self.App - self.Apv @ sp.diags(1.0 / rowsum) @ self.Avp).tocsr()
        self.B2pp = Ilu0Inverse(B2pp)
        self.tpp_its = 0

    def Tpp(self, x):
        return self.App @ x - self.Apv @ self.Pvv(self.Avp @ x)

    def __call__(self, src):
        su, sp_ = src[: self.n_u], src[self.n_u:]
        ptmp = sp_ - self.Apv @ self.Pvv(su)
        c = ptmp.copy()
        Sc = self.Tpp(c)
        den = Sc @ c
        x0 = c * ((ptmp @ c) / den) if den != 0 else np.zeros_like(c)
        dp, its = gmres_left(self.Tpp, self.B2pp, ptmp, x0, 1e-3 * np.linalg.norm(ptmp), ptmp.size, 200)
        self.tpp_its += its
        du = self.Pvv(su) - self.Pvv(self.Avp @ dp)
        return np.concatenate([du, dp])


def solve(o, use_nonzero_constraints, rel_tol=1e-6):
    """SUPGFluidSolver::solve (:297-328) for the oracle solver `o` (system_matrix / system_rhs assembled): returns
    (newton_update, FGMRES iterations, inner GMRES iterations)"""
    from . import fem
    from .ins import CsrOp, fgmres

    P = BlockIncompSchurPreconditioner(o.system_matrix, o.n_u)
    x, its, _ = fgmres(CsrOp(o.system_matrix), P, o.system_rhs, rel_tol * np.linalg.norm(o.system_rhs), o.n)
    x[o.con != 0] = o.nonzero_val[o.con != 0] if use_nonzero_constraints else 0.0
    fem.distribute(o.dofs, x)
    return x, its, P.tpp_its
