"""CPU experiment (oracle only): geometric multigrid for the INNER A_uu solve of the Schur preconditioner.

Question: A~^-1 of BlockSchurPreconditioner (mpi_insim.cpp:124-127; here an inner Krylov solve to 1e-1) costs about 65
node-block-Jacobi BiCGStab iterations = 130 products with A_uu per application at config 3 (128^3 cells) and grows like
1/h. Does a V-cycle on rediscretised coarse operators (Q2 nested spaces, evaluation point injected, Chebyshev /
node-block-Jacobi smoother) bring that down to a mesh-independent handful of products despite the grad-div term?

dt is chosen per fine mesh so that the mass/stiffness ratio h^2 / (dt (mu + gamma)) equals config 3's (h = 1/128, dt = 1e-2).
Work is counted in fine-level products (a product on level l costs 8^-l)."""
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from util import cavity_prm, make_oracle

from oracle import ins

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n_levels = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dt = (1.0 / n) ** 2 / ((1.0 / 128) ** 2 / 1e-2)
amp = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0


def swirl(pts):
    """a smooth cavity-like velocity (vanishes on the walls except the lid z = 1), |u| <= amp"""
    x, y, z = pts.T
    u = np.zeros_like(pts)
    u[:, 0] = amp * np.sin(np.pi * x) ** 2 * z * z * (1 + 0.5 * np.sin(2 * np.pi * y))
    u[:, 2] = -amp * np.sin(2 * np.pi * x) * z * (1 - z) * np.sin(np.pi * y)
    u[:, 1] = 0.3 * amp * np.sin(np.pi * y) ** 2 * np.sin(2 * np.pi * z) * np.sin(np.pi * x)
    return u


class Level:
    def __init__(self, n):
        t0 = time.perf_counter()
        o = make_oracle(cavity_prm(3, dt=dt), (n,) * 3, (0, 0, 0), (1, 1, 1))
        d = o.dofs
        ev = np.zeros(o.n)
        ev[: o.n_u] = swirl(d.ucoords).ravel()
        o.evaluation_point[:] = ev
        o.present[:] = ev
        A, _, _ = o.assemble(False, with_mass=False)
        nu = o.n_u
        self.n, self.nu, self.nn = n, nu, nu // 3
        self.A = A[:nu, :nu].tocsr()
        self.con = o.con[:nu] != 0
        self.coords = d.ucoords
        # node-block Jacobi
        nn = self.nn
        blocks = np.zeros((nn, 3, 3))
        idx = np.arange(nn) * 3
        for c in range(3):
            for e in range(3):
                blocks[:, c, e] = np.asarray(self.A[idx + c, idx + e]).ravel()
        self.binv = np.linalg.inv(blocks)
        self.op = ins.CsrOp(self.A)
        # lambda_max of D^-1 A by a few power iterations
        rng = np.random.default_rng(0)
        v = rng.standard_normal(nu)
        lam = 1.0
        for _ in range(15):
            w = self.jac(self.op(v))
            lam = np.linalg.norm(w) / np.linalg.norm(v)
            v = w / np.linalg.norm(w)
        self.lmax = lam
        print(f"level n={n}: {nu} dofs, nnz {self.A.nnz}, lambda_max(D^-1 A) ~ {lam:.3f}, {time.perf_counter() - t0:.1f} s", flush=True)

    def jac(self, v):
        return np.einsum("nce,ne->nc", self.binv, v.reshape(self.nn, 3)).ravel()


def prolongation(coarse: Level, fine: Level):
    """Q2 nested interpolation, node lattices 2n+1 per direction (lexicographic by position), kron I3"""
    nc, nf = coarse.n, fine.n
    # 1-D: fine lattice index i in [0, 4 nc], coarse lattice j in [0, 2 nc]; coarse cell k spans coarse lattice 2k..2k+2
    rows, cols, vals = [], [], []
    for i in range(2 * nf + 1):
        k = min(i // 4, nc - 1)
        xi = (i - 4 * k) / 4.0
        w = [(1 - xi) * (1 - 2 * xi), 4 * xi * (1 - xi), xi * (2 * xi - 1)]
        for a in range(3):
            if abs(w[a]) > 1e-14:
                rows.append(i), cols.append(2 * k + a), vals.append(w[a])
    P1 = sp.csr_matrix((vals, (rows, cols)), shape=(2 * nf + 1, 2 * nc + 1))

    def lattice(lv):
        h = 1.0 / (2 * lv.n)
        ijk = np.rint(lv.coords / h).astype(np.int64)
        m = 2 * lv.n + 1
        return ijk[:, 0] + m * (ijk[:, 1] + m * ijk[:, 2])

    P3 = sp.kron(P1, sp.kron(P1, P1)).tocsr()  # lattice index = x + m (y + m z): kron order (z, y, x)
    lf, lc = lattice(fine), lattice(coarse)
    P = P3[lf][:, lc]
    P = sp.kron(P, sp.identity(3)).tocsr()
    # constrained dofs carry no correction
    keep_f = sp.diags((~fine.con).astype(float))
    keep_c = sp.diags((~coarse.con).astype(float))
    return (keep_f @ P @ keep_c).tocsr()


levels = []
m = n
for l in range(n_levels):
    levels.append(Level(m))
    m //= 2
P = [prolongation(levels[l + 1], levels[l]) for l in range(n_levels - 1)]
coarse_lu = spla.splu(levels[-1].A.tocsc())
work = [0.0]


def cheb(lv, l, b, x, degree, alpha):
    """Chebyshev iteration on D^-1 A, eigenvalue window [lmax/alpha, 1.1 lmax]; x = None means zero initial guess"""
    lo, hi = lv.lmax / alpha, 1.1 * lv.lmax
    theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo)
    sigma = theta / delta
    rho = 1.0 / sigma
    if x is None:
        r = b.copy()
        x = np.zeros_like(b)
    else:
        r = b - lv.op(x)
        work[0] += 8.0 ** -l
    d = lv.jac(r) / theta
    for k in range(degree):
        x = x + d
        if k == degree - 1:
            break
        r = r - lv.op(d)
        work[0] += 8.0 ** -l
        rho_new = 1.0 / (2 * sigma - rho)
        d = rho_new * rho * d + 2 * rho_new / delta * lv.jac(r)
        rho = rho_new
    return x


def vcycle(l, b, degree, alpha, coarse_solver="lu"):
    lv = levels[l]
    if l == n_levels - 1:
        if coarse_solver == "lu":
            return coarse_lu.solve(b)
        return cheb(lv, l, b, None, 8, 30.0)
    x = cheb(lv, l, b, None, degree, alpha)
    r = b - lv.op(x)
    work[0] += 8.0 ** -l
    xc = vcycle(l + 1, P[l].T @ r, degree, alpha, coarse_solver)
    x = x + P[l] @ xc
    return cheb(lv, l, b, x, degree, alpha)


fine = levels[0]
rng = np.random.default_rng(1)
# right-hand sides like the ones the preconditioner sees: a smooth field and a rough one (both zero on constrained dofs)
b_smooth = swirl(fine.coords).ravel() * (~fine.con)
b_rough = rng.uniform(-1, 1, fine.nu) * (~fine.con)
for name, b in [("smooth", b_smooth), ("rough", b_rough)]:
    nb = np.linalg.norm(b)
    for rel in (1e-1, 1e-2):
        x, it, res = ins.bicgstab(fine.op, fine.jac, b, rel * nb, 5000)
        print(f"[{name}] rel {rel:g}: Jacobi-BiCGStab {it} its = {2 * it} products", flush=True)
        for degree, alpha in [(2, 4.0), (3, 6.0), (3, 10.0), (4, 10.0), (4, 20.0), (5, 30.0)]:
            for cs in ("lu", "cheb"):
                work[0] = 0.0
                fine.op.n_apply = 0
                x, its, res = ins.fgmres(fine.op, lambda v: vcycle(0, v, degree, alpha, cs), b, rel * nb, 200)
                true = np.linalg.norm(b - fine.A @ x) / nb
                print(f"    V(cheb {degree}, window 1/{alpha:g}, coarse {cs}) + FGMRES: {its} its, {work[0] + its:.1f} fine-product units, true rel res {true:.2e}",
                      flush=True)
