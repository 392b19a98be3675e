"""Hanging-node constraints of locally refined meshes in the oracle (reference: DoFTools::make_hanging_node_constraints at
source/mpi_fluid_solver.cpp:182-184, the band refinement of tests/fsi_leaflet_mpi/fsi_leaflet_mpi.cpp:66-76 and
tests/fsi-wall-3D/fsi-wall-3D.cpp:47-53, AffineConstraints::distribute_local_to_global at source/mpi_scnsim.cpp:548-560).

The reference pins none of this directly ("parity unpinned": its FSI cases are smoke tests); the oracle is pinned here on
three properties that any correct implementation of the constraints has:
  * the hanging vertices the oracle finds from the geometry are the ones the product's mesh class records;
  * the cell-wise constrained scatter equals the algebraic condensation C^T A C, C^T (b - A g) of the unconstrained assembly;
  * patch tests: the lines make the Q1 space conforming (Laplace stiffness matrix condensed through them: a globally linear
    field has zero residual at the free interior nodes), and a uniform state leaves a zero SCnsIM Newton residual on a
    band-refined mesh - neither holds when the lines are dropped."""
import numpy as np
import pytest
import scipy.sparse as sp

from test_scns_gpu import scns_prm


def refined_mesh(dim):
    """box with a refined band in the middle (the shape of the reference's FSI meshes)"""
    import openifem_b200 as ifem
    from oracle import grid

    tria = ifem.Triangulation(dim)
    reps, hi = ((8, 4), (2.0, 1.0)) if dim == 2 else ((3, 3, 6), (1.0, 1.0, 1.5))
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0,) * dim, hi, True)
    v, c, _ = tria.get_mesh()
    cen = v[c].mean(axis=1)
    flags = ((cen[:, 0] > 0.5) & (cen[:, 0] < 1.25)) if dim == 2 else ((cen[:, 2] > 0.5) & (cen[:, 2] < 1.0))
    tria.execute_refinement(flags.astype(np.uint8))
    v, c, b = tria.get_mesh()
    return tria, (grid.QuadMesh(v, c, b) if dim == 2 else grid.HexMesh(v, c, b))


@pytest.mark.parametrize("dim", [2, 3])
def test_hanging_vertices_match_mesh_class(dim):
    from oracle import fem

    tria, mesh = refined_mesh(dim)
    d = fem.FluidDofs(mesh, 1, 1)
    hv, hk, hm = tria.hanging()
    assert hv.size == len(d.hanging_u) > 0
    v = mesh.vertices
    mine = {tuple(np.round(d.ucoords[h], 9)): sorted(tuple(np.round(d.ucoords[m], 9)) for m in ms) for h, (ms, _) in d.hanging_u.items()}
    theirs = {tuple(np.round(v[h], 9)): sorted(tuple(np.round(v[m], 9)) for m in hm[i, :hk[i]]) for i, h in enumerate(hv)}
    assert mine == theirs
    if dim == 3:
        assert {len(ms) for ms, _ in d.hanging_u.values()} == {2, 4}


def _scns(dim, **kw):
    from oracle import prm, scns

    _, mesh = refined_mesh(dim)
    return scns.SCnsIM(mesh, prm.Params(scns_prm(dim, **kw), is_text=True))


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("nonzero", [True, False])
def test_constrained_scatter_equals_condensation(dim, nonzero):
    from oracle import fem

    full = 3 if dim == 2 else 7
    dirichlet = {0: (full, [1.0, 0.5, -0.25][:dim]), 2: (full, [0.0] * dim), 3: (full, [0.0] * dim)}
    o = _scns(dim, dirichlet=dirichlet, gravity=[1.0, -9.8, 0.5][:dim])
    rng = np.random.default_rng(5)
    o.evaluation_point[:] = rng.uniform(-1, 1, o.n)
    o.present[:] = rng.uniform(-1, 1, o.n)
    o.indicator[:] = (rng.uniform(size=o.mesh.n_cells) < 0.3).astype(np.int32)
    o.fsi_acceleration[:] = rng.uniform(-1, 1, o.n)
    A, b = o.assemble(nonzero)
    A, b = A.copy(), b.copy()
    # unconstrained assembly on the same pattern: no Dirichlet lines, hanging dofs treated as free
    d = o.dofs
    saved, saved_con = d.hanging_dofs, o.con
    d.hanging_dofs, o.con = {}, np.zeros_like(o.con)
    A0, b0 = o.assemble(False)
    d.hanging_dofs, o.con = saved, saved_con
    con, g, ptr, master, weight = fem.resolve_constraints(d, o.con, o.nonzero_val if nonzero else np.zeros(o.n))
    free = np.nonzero(con == 0)[0]
    C = sp.lil_matrix((o.n, o.n))
    for i in free:
        C[i, i] = 1.0
    for i in np.nonzero(con == 2)[0]:
        for k in range(ptr[i], ptr[i + 1]):
            C[i, master[k]] = weight[k]
    C = C.tocsr()
    Ac = (C.T @ A0 @ C).tocsr()
    bc = C.T @ (b0 - A0 @ g)
    sel = sp.diags((con == 0).astype(float))
    err_A = sp.linalg.norm(sel @ A @ sel - Ac) / sp.linalg.norm(Ac)
    err_b = np.linalg.norm(b[free] - bc[free]) / np.linalg.norm(bc[free])
    assert err_A < 1e-13 and err_b < 1e-13, (err_A, err_b)
    # constrained rows: nothing but a positive diagonal, rhs = diagonal * inhomogeneity
    rows = np.nonzero(con)[0]
    Ad = A.tocsr()[rows]
    assert abs(Ad - sp.diags(A.diagonal()).tocsr()[rows]).max() == 0.0 and (A.diagonal()[rows] > 0).all()
    assert np.allclose(b[rows], A.diagonal()[rows] * (g[rows] if nonzero else 0.0), rtol=1e-13, atol=0)
    # and nothing is left in the columns of constrained dofs
    assert abs((A @ sp.diags((con != 0).astype(float))) - sp.diags(A.diagonal() * (con != 0))).max() == 0.0


@pytest.mark.parametrize("dim", [2, 3])
def test_patch_test_laplace(dim):
    """the hanging-node lines make the Q1 space conforming: with the stiffness matrix of the Laplacian condensed through them,
    a globally linear field has zero residual at every free interior node (and not without the lines)"""
    from oracle import fem

    _, mesh = refined_mesh(dim)
    d = fem.FluidDofs(mesh, 1, 1)
    tab, n, x = d.pnodes, d.n_pnodes, d.pcoords
    fe = fem.FEQ(dim, 1)
    qp, qw = fem.qgauss(dim, 2)
    _, dN = fe.eval(qp)
    K = sp.lil_matrix((n, n))
    for cn in tab:
        J = np.einsum("vi,qvj->qij", x[cn], dN)
        G = np.einsum("qaj,qjk->qak", dN, np.linalg.inv(J))
        K[np.ix_(cn, cn)] += np.einsum("qak,qbk,q->ab", G, G, qw * np.abs(np.linalg.det(J)))
    K = K.tocsr()
    C = sp.lil_matrix((n, n))
    for i in range(n):
        if i in d.hanging_p:
            ms, w = d.hanging_p[i]
            for m in ms:
                C[i, m] = w
        else:
            C[i, i] = 1.0
    C = C.tocsr()
    lin = 0.3 + x @ np.array([0.8, -0.4, 0.55][:dim])
    assert np.allclose(C @ lin, lin, rtol=0, atol=1e-14)  # the lines reproduce a linear field at the hanging nodes
    lo, hi = x.min(axis=0), x.max(axis=0)
    interior = np.all((x > lo + 1e-12) & (x < hi - 1e-12), axis=1)
    free = interior & ~np.isin(np.arange(n), list(d.hanging_p))
    r = (C.T @ K @ C @ lin)[free]
    assert np.abs(r).max() < 1e-13, np.abs(r).max()
    assert np.abs((K @ lin)[interior]).max() > 1e-3  # the unconstrained (non-conforming) assembly does not pass


@pytest.mark.parametrize("dim", [2, 3])
def test_patch_test_uniform_state(dim):
    """SCnsIM on the band-refined mesh: uniform velocity and pressure are an exact steady state (every term of
    mpi_scnsim.cpp:429-512 carries a gradient or a time difference except p div(phi_i), which only cancels between the cells
    around a node when the scatter through the hanging-node lines is consistent)"""
    full = 3 if dim == 2 else 7
    o = _scns(dim, dirichlet={i: (full, [0.0] * dim) for i in range(2 * dim)}, mu=0.7, rho=1.1)
    d = o.dofs
    o.present[: o.n_u] = np.tile([0.3, -0.2, 0.45][:dim], d.n_unodes)
    o.present[o.n_u:] = 2.5
    o.evaluation_point[:] = o.present
    _, rhs = o.assemble(False)
    assert np.abs(rhs).max() < 1e-13, np.abs(rhs).max()
    saved = d.hanging_dofs
    d.hanging_dofs = {}
    _, rhs_free = o.assemble(False)
    d.hanging_dofs = saved
    assert np.abs(rhs_free).max() > 1e-3
