"""Probe (development aid, CPU, oracle): does a LAGGED SYMMETRIC stand-in for A_uu inside the preconditioner - mass + viscous +
grad-div terms only, i.e. the matrix assembled at zero velocity, solved with block-Jacobi CG instead of BiCGStab on the current
A_uu - change the FGMRES / Newton iteration counts?  python scripts/probe_lagged_inner.py [cells] [steps] [developed 0|1]"""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import scipy.sparse as sp

from oracle import ins as O
from util import cavity_prm, make_oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
developed = int(sys.argv[3]) if len(sys.argv) > 3 else 0


def pcg(A, prec, b, tol, max_it):
    x = np.zeros_like(b)
    r = b.copy()
    z = prec(r)
    p = z.copy()
    rz = r @ z
    it = 0
    while np.linalg.norm(r) > tol and it < max_it:
        Ap = A(p)
        a = rz / (p @ Ap)
        x += a * p
        r -= a * Ap
        z = prec(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
        it += 1
    return x, it


def run(mode):
    o = make_oracle(cavity_prm(3), (n, n, n), (0, 0, 0), (1, 1, 1), a_inv=("bicgstab", 1e-1, 4000))
    if developed:  # a swirl of magnitude ~1 everywhere instead of a start from rest
        x = o.dofs.ucoords
        u = np.zeros((o.dofs.n_unodes, 3))
        u[:, 0] = np.sin(np.pi * x[:, 0]) * np.cos(np.pi * x[:, 2])
        u[:, 2] = -np.cos(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 2])
        o.present[: o.n_u] = u.ravel()
        o.present[o.con != 0] = 0
    lin = {}
    if mode != "current":
        # A_lin: the velocity block assembled at zero velocity (no convection terms), same constraints
        keep = o.evaluation_point.copy(), o.present.copy()
        o.evaluation_point[:] = 0
        o.present[:] = 0
        S, _, _ = o.assemble(False)
        lin["A"] = O.csr_split(S, o.n_u)[0].tocsr()
        o.evaluation_point[:], o.present[:] = keep
        if mode == "lagged_bicgstab":
            o.a_inv_filter = lambda A: lin["A"]
        else:
            orig = o._make_preconditioner

            def patched():
                vm = orig()
                A = lin["A"]
                A_op = O.CsrOp(A)
                dim, nn = 3, o.n_u // 3
                blocks = np.zeros((nn, 3, 3))
                for c in range(3):
                    for e in range(3):
                        blocks[:, c, e] = np.asarray(A[np.arange(nn) * 3 + c, np.arange(nn) * 3 + e]).ravel()
                prec = O.BlockJacobi(np.linalg.inv(blocks))
                # replace a_inverse inside vmult's closure
                for cell in vm.__closure__:
                    pass
                stats = o.precond_stats
                import types

                freevars = vm.__code__.co_freevars
                idx = freevars.index("a_inverse")

                def a_inverse(v):
                    xx, it = pcg(A_op, prec, v, 1e-1 * np.linalg.norm(v), 4000)
                    stats["a_inv"] += it
                    return xx

                vm.__closure__[idx].cell_contents = a_inverse
                return vm

            o._make_preconditioner = patched
    tot = {"fgmres": 0, "a_inv": 0, "newton": 0}
    for k in range(steps):
        h0 = len(o.history)
        o.run_one_step(k == 0)
        for h in o.history[h0:]:
            tot["fgmres"] += h[4]
            tot["newton"] += 1
    return tot, o


for mode in ["current", "lagged_bicgstab", "lagged_cg"]:
    stats_total = 0
    # count inner iterations by wrapping
    import oracle.ins as M

    calls = {"its": 0}
    orig_b = M.bicgstab

    def counting(*a, **k):
        r = orig_b(*a, **k)
        calls["its"] += r[1]
        return r

    M.bicgstab = counting
    tot, o = run(mode)
    M.bicgstab = orig_b
    print(mode, tot, "bicgstab its", calls["its"], "cg its (lagged_cg)", o.precond_stats if mode == "lagged_cg" else "", "|u|", np.linalg.norm(o.velocity()), flush=True)
