"""Probe (development aid, CPU, oracle): the Jacobi variant of the SUPG block preconditioner (several ranks / large meshes) on the
mass-dominated fsi-wall-3D system (dt = 1e-6): how do the FGMRES and inner T_pp iteration counts change when P_vv^-1 is k steps of
Chebyshev-accelerated block-Jacobi on A_vv instead of one block-Jacobi step?  Cost unit: one pass over A_vv = 9 units, over
A_vp / A_pv = 3 units each, A_pp = 1 (entries of the Q1 blocks).   python scripts/probe_supg_pvv.py"""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import scipy.sparse as sp

import bench
from oracle import grid, ins as O, prm, scns

ftria, _ = bench.fsi_meshes(5, 1, 1, half=True)
v, c, b = ftria.get_mesh()
o = scns.SCnsIM(grid.HexMesh(v, c, b), prm.Params(bench.fsi_prm_path(5)))
o.run_one_step(True)
o.evaluation_point = o.present.copy()
o.assemble(False)
A = o.system_matrix.tocsr()
rhs = o.system_rhs
nu = o.n_u
Avv, Avp, Apv, App = A[:nu, :nu].tocsr(), A[:nu, nu:].tocsr(), A[nu:, :nu].tocsr(), A[nu:, nu:].tocsr()
dim, nn = 3, nu // 3
blocks = np.zeros((nn, 3, 3))
for a in range(3):
    for e in range(3):
        blocks[:, a, e] = np.asarray(Avv[np.arange(nn) * 3 + a, np.arange(nn) * 3 + e]).ravel()
binv = np.linalg.inv(blocks)
jac = lambda x: np.einsum("nce,ne->nc", binv, x.reshape(nn, 3)).ravel()
cost = {"units": 0}


def Avv_mul(x):
    cost["units"] += 9
    return Avv @ x


def cheb(k, lmin, lmax):
    """k steps of Chebyshev iteration on D^-1 A_vv x = D^-1 b from x = 0"""
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)

    def apply(bv):
        if k == 1:
            return jac(bv) / theta
        x = np.zeros_like(bv)
        r = bv.copy()
        sigma = theta / delta
        rho = 1.0 / sigma
        d = jac(r) / theta
        for it in range(k):
            x = x + d
            if it == k - 1:
                break
            r = r - Avv_mul(d)
            rho_new = 1.0 / (2 * sigma - rho)
            d = rho_new * rho * d + 2 * rho_new / delta * jac(r)
            rho = rho_new
        return x

    return apply


# eigenvalue range of D^-1 A_vv (power iteration / a few Lanczos steps would do on the device; here exact-ish)
from scipy.sparse.linalg import LinearOperator, eigs

op = LinearOperator((nu, nu), matvec=lambda x: jac(Avv @ x))
lmax = float(np.real(eigs(op, k=1, which="LM", return_eigenvectors=False, tol=1e-3)[0]))
lmin = float(np.real(eigs(op, k=1, which="SM", return_eigenvectors=False, tol=1e-2, maxiter=5000)[0])) if nu < 20000 else lmax / 30
print(f"D^-1 A_vv spectrum ~ [{lmin:.3g}, {lmax:.3g}], {nu} velocity dofs", flush=True)

rowsum_inv = 1.0 / np.asarray(abs(Avv).sum(axis=1)).ravel()
b2diag = App.diagonal() - np.asarray((Apv.multiply(rowsum_inv[None, :])).multiply(Avp.T).sum(axis=1)).ravel()


def run(Pvv, label):
    cost["units"] = 0
    inner = {"its": 0}

    def Tpp(x):
        cost["units"] += 3 + 3 + 1
        return App @ x - Apv @ Pvv(Avp @ x)

    def prec(src):
        su, sp_ = src[:nu], src[nu:]
        cost["units"] += 3
        ptmp = sp_ - Apv @ Pvv(su)
        tol = 1e-3 * np.linalg.norm(ptmp)
        if tol > 0:
            dp, its, _ = O.fgmres(Tpp, lambda x: x / b2diag, ptmp, tol, ptmp.size, 50)
        else:
            dp, its = np.zeros_like(ptmp), 0
        inner["its"] += its
        cost["units"] += 3
        du = Pvv(su - Avp @ dp)
        return np.concatenate([du, dp])

    def Aop(x):
        cost["units"] += 16
        return A @ x

    x, its, res = O.fgmres(Aop, prec, rhs, 1e-6 * np.linalg.norm(rhs), A.shape[0], 30)
    print(f"{label:28s}: FGMRES {its:3d}, inner T_pp {inner['its']:5d}, cost {cost['units'] / 1e3:8.1f} k units", flush=True)


run(jac, "block-Jacobi (current)")
for k in (2, 3, 4):
    run(cheb(k, lmax / 20, 1.05 * lmax), f"Chebyshev k={k} [lmax/20, lmax]")
