"""ILU(0) factors of the SUPG block preconditioner on the device (csrc/ilu0.cu) - the stand-in for the reference's Hypre Euclid
factors (source/mpi_supg_solver.cpp:51, 130-133, source/preconditioner_pilut.cpp:124-138; Hypre is not vendored: "parity unpinned"
for its arithmetic, ILU(0) itself is a fixed algorithm): factors and the two triangular sweeps against a plain Python IKJ ILU(0)
(1e-13), and what they buy - the inner T_pp iteration counts of a viscous SUPG case against the Jacobi factors of round 1, with the
converged fields unchanged (1e-6 against the oracle's sparse direct solves)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu


def ilu0_reference(A):
    """textbook IKJ ILU(0) restricted to the pattern of A (Saad, Iterative Methods, alg. 10.4); returns the factors in A's pattern"""
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
    return v, diag


@pytest.mark.parametrize("n,stencil", [(14, 1), (11, 2)])
def test_ilu0_factors_and_sweeps_match_the_textbook_algorithm(n, stencil):
    import openifem_b200 as ifem

    # a nonsymmetric convection-diffusion-like matrix on an n x n grid with a (2 stencil + 1)^2 pattern (stencil 2 = the pattern of B2pp)
    rng = np.random.default_rng(2)
    idx = np.arange(n * n).reshape(n, n)
    rows, cols, vals = [], [], []
    for i in range(n):
        for j in range(n):
            for di in range(-stencil, stencil + 1):
                for dj in range(-stencil, stencil + 1):
                    if 0 <= i + di < n and 0 <= j + dj < n:
                        rows.append(idx[i, j])
                        cols.append(idx[i + di, j + dj])
                        vals.append((8.0 * (2 * stencil + 1) ** 2 if di == 0 and dj == 0 else -1.0) + 0.3 * rng.uniform(-1, 1))
    A = sp.csr_matrix((vals, (rows, cols)), shape=(n * n, n * n))
    A.sort_indices()
    b = rng.uniform(-1, 1, n * n)
    f, x, (nl, nu) = ifem.ilu0_apply(A, b)
    f_ref, diag = ilu0_reference(A)
    assert np.abs(f - f_ref).max() <= 1e-13 * np.abs(f_ref).max()
    LU = sp.csr_matrix((f_ref, A.indices, A.indptr), shape=A.shape)
    L = sp.tril(LU, -1).tocsr() + sp.identity(n * n, format="csr")
    Um = sp.triu(LU, 0).tocsr()
    x_ref = spla.spsolve_triangular(Um, spla.spsolve_triangular(L, b, lower=True), lower=False)
    assert np.abs(x - x_ref).max() <= 1e-12 * np.abs(x_ref).max()
    assert 1 < nl < n * n and 1 < nu < n * n  # the sweeps are level scheduled, not sequential


def test_ilu_factors_cut_the_inner_iterations_of_a_viscous_supg_case():
    import openifem_b200 as ifem
    from oracle import fem, prm, scns
    from test_scns_gpu import scns_prm

    # viscous, large time step: the regime of the reference's SUPG goldens where diagonal factors stall
    text = scns_prm(2, dt=5e-2, mu=0.5, rho=1.0, newton_tol=1e-8)
    reps, hi = (30, 10), (3.0, 1.0)
    o = scns.SCnsIM(fem.BoxMesh(reps, (0, 0), hi), prm.Params(text, is_text=True))
    counts = {}
    for mode in (0, 1):
        tria = ifem.Triangulation(2)
        ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0, 0), hi, True)
        g = ifem.Fluid.MPI.SCnsIM(tria, ifem.Parameters.AllParameters(text=text))
        g.setup()
        g.set_control(fgmres_rel=1e-10, supg_ilu=mode)
        for k in range(2):
            g.run_one_step(k == 0)
        h = g.history()
        counts[mode] = (sum(r["gmres_its"] for r in h), sum(r["a_inv_its"] for r in h))
        if mode == 0:
            for k in range(2):
                o.run_one_step(k == 0)
        sol = g.get_current_solution()
        rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
        assert rel(sol[: o.n_u], o.velocity()) < 1e-6 and rel(sol[o.n_u:], o.pressure()) < 1e-6
    (outer_j, inner_j), (outer_i, inner_i) = counts[0], counts[1]
    print(f"Jacobi factors: {outer_j} FGMRES / {inner_j} inner T_pp iterations; ILU(0): {outer_i} / {inner_i}")
    assert outer_i <= outer_j and inner_i * 3 <= inner_j


def test_iteration_counts_against_the_references_preconditioner():
    """the device's BlockIncompSchurPreconditioner with ILU(0) factors against the oracle's restatement of the reference's
    (oracle/supg_precond.py: ILU(0) for P_vv and B2pp, left-preconditioned GMRES(200) from the line-search guess) on the same Newton
    systems of a viscous case: FGMRES and inner T_pp iteration counts within a factor 2 of each other (the device restarts the inner
    solve after 50 vectors, preconditions it from the right and starts from zero)"""
    import openifem_b200 as ifem
    from oracle import fem, prm, scns, supg_precond
    from test_scns_gpu import scns_prm

    text = scns_prm(2, dt=5e-2, mu=0.5, rho=1.0, newton_tol=1e-8)
    reps, hi = (40, 12), (4.0, 1.2)
    o = scns.SCnsIM(fem.BoxMesh(reps, (0, 0), hi), prm.Params(text, is_text=True))
    tria = ifem.Triangulation(2)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0, 0), hi, True)
    g = ifem.Fluid.MPI.SCnsIM(tria, ifem.Parameters.AllParameters(text=text))
    g.setup()
    g.set_control(supg_ilu=1)  # default tolerance: 1e-6 |rhs| as in the reference
    g.run_one_step(True)
    dev = [(r["gmres_its"], r["a_inv_its"]) for r in g.history()]
    # the oracle walks the same Newton iteration with the reference's preconditioner
    o.timestep, o.time = 1, o.dt
    o.evaluation_point = o.present.copy()
    ref = []
    for it in range(len(dev)):
        o.assemble(it == 0)
        x, outer, inner = supg_precond.solve(o, it == 0)
        o.evaluation_point = o.evaluation_point + x
        ref.append((outer, inner))
    print("device (FGMRES, inner):", dev, " reference algorithm:", ref)
    for (do, di), (ro, ri) in zip(dev, ref):
        assert do <= 2 * ro + 2 and ro <= 2 * do + 2, (dev, ref)
        assert di <= 2 * ri + 10 and ri <= 2 * di + 10, (dev, ref)
    sol = g.get_current_solution()
    assert np.linalg.norm(sol - o.evaluation_point) / np.linalg.norm(sol) < 1e-4  # both solved to 1e-6 |rhs| per Newton step
