"""Probe (development aid, CPU, oracle): does an additive aggregation coarse space (piecewise constants over G^3 boxes, Galerkin
coarse matrix, exact coarse solve) on top of the node-block Jacobi preconditioner cut the iterations of (a) the inner A~^-1
BiCGStab solve and (b) "CG for Sm"?  dt is matched to config 3's mass / stiffness ratio.
    python scripts/probe_coarse_space.py [cells] [G]"""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import ins as O
from util import cavity_prm, make_oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
G = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dt = (1.0 / n) ** 2 / ((1.0 / 128) ** 2 / 1e-2)
o = make_oracle(cavity_prm(3, dt=dt), (n, n, n), (0, 0, 0), (1, 1, 1), a_inv=("bicgstab", 1e-1, 4000))
o.run_one_step(True)  # a developed-enough state
S, Mm, rhs = o.assemble(False)
nu = o.n_u
Auu, Bt, B, _ = O.csr_split(S, nu)
Auu = Auu.tocsr()
free = o.con[:nu] == 0
rng = np.random.default_rng(0)


def aggregates(coords, ncomp, mask):
    box = np.minimum((coords * G).astype(int), G - 1)
    agg = (box[:, 0] + G * (box[:, 1] + G * box[:, 2]))
    rows = np.arange(coords.shape[0] * ncomp)
    cols = np.repeat(agg, ncomp) * ncomp + np.tile(np.arange(ncomp), coords.shape[0])
    Z = sp.csr_matrix((mask.astype(float), (rows, cols)), shape=(rows.size, G ** 3 * ncomp))
    keep = np.asarray(Z.sum(axis=0)).ravel() > 0
    return Z[:, keep]


def krylov(A, prec, b, rel, symmetric):
    it = [0]

    def cb(*a):
        it[0] += 1

    M = spla.LinearOperator(A.shape, matvec=prec)
    if symmetric:
        x, info = spla.cg(A, b, rtol=rel, M=M, callback=cb, maxiter=5000)
    else:
        x, info = spla.bicgstab(A, b, rtol=rel, M=M, callback=cb, maxiter=5000)
    return it[0]


# (a) A_uu
nn = nu // 3
blocks = np.zeros((nn, 3, 3))
for c in range(3):
    for e in range(3):
        blocks[:, c, e] = np.asarray(Auu[np.arange(nn) * 3 + c, np.arange(nn) * 3 + e]).ravel()
binv = np.linalg.inv(blocks)
jac = lambda v: np.einsum("nce,ne->nc", binv, v.reshape(nn, 3)).ravel()
Z = aggregates(o.dofs.ucoords, 3, free)
E = (Z.T @ Auu @ Z).toarray()
Einv = np.linalg.inv(E)
two = lambda v: jac(v) + Z @ (Einv @ (Z.T @ v))
for trial in range(2):
    b = rng.uniform(-1, 1, nu) * free if trial == 0 else (Bt @ rng.uniform(-1, 1, o.n - nu)) * free
    print(f"A_uu ({'random' if trial == 0 else 'B^T p'} rhs), {nu} dofs, coarse {E.shape[0]}: BiCGStab its to 1e-1: Jacobi {krylov(Auu, jac, b, 1e-1, False)}, "
          f"+coarse {krylov(Auu, two, b, 1e-1, False)}; to 1e-3: {krylov(Auu, jac, b, 1e-3, False)} / {krylov(Auu, two, b, 1e-3, False)}", flush=True)

# (b) S_m
inv_diag = 1.0 / Mm.diagonal()[:nu]
Sm = (B @ sp.diags(inv_diag) @ Bt).tocsr()
npn = Sm.shape[0]
Zp = aggregates(o.dofs.pcoords, 1, np.ones(npn, bool))
Ep = (Zp.T @ Sm @ Zp).toarray()
Ep += np.mean(np.diag(Ep)) * np.ones_like(Ep) / Ep.shape[0]  # the constant null space
Epinv = np.linalg.inv(Ep)
d = 1.0 / Sm.diagonal()
for omega in (1.0,):
    twop = lambda v: omega * d * v + Zp @ (Epinv @ (Zp.T @ v))
    b = B @ (rng.uniform(-1, 1, nu) * free)
    b -= b.mean()
    print(f"S_m, {npn} dofs, coarse {Ep.shape[0]}: CG its to 1e-3: none {krylov(Sm, lambda v: v, b, 1e-3, True)}, Jacobi {krylov(Sm, lambda v: d * v, b, 1e-3, True)}, "
          f"Jacobi+coarse {krylov(Sm, twop, b, 1e-3, True)}", flush=True)
