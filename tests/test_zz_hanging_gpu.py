"""GPU parity on locally refined meshes (hanging nodes): the band-refined channels of the reference's FSI cases
(tests/fsi_leaflet_mpi/fsi_leaflet_mpi.cpp:66-76, tests/fsi-wall-3D/fsi-wall-3D.cpp:47-53; BASELINE configs 4 and 5).

Product: cell kernels + post-assembly condensation through the hanging-node lines (openifem_b200/csrc/hanging.cu), called through
the C ABI. Oracle: cell-wise constrained scatter as AffineConstraints::distribute_local_to_global does it
(oracle/csrc/oracle_common.h) - two different routes to C^T A C, pinned in tests/test_hanging_oracle_cpu.py.

Tolerances: assembled matrix / rhs 1e-12 relative (the diagonal of a hanging row: |sum of the local diagonals| here, sum of
their absolute values in deal.II - equal for these operators, whose local diagonals are positive); fields after time steps
1e-6 relative (device FGMRES to 1e-10 |rhs|, oracle sparse direct); the uniform-state patch test 1e-13 absolute."""
import numpy as np
import pytest
import scipy.sparse as sp

from test_hanging_oracle_cpu import refined_mesh
from test_scns_gpu import scns_prm

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def _make(dim, cls="SCnsIM", **kw):
    import openifem_b200 as ifem
    from oracle import prm, scns

    tria, mesh = refined_mesh(dim)
    text = scns_prm(dim, **kw)
    o = getattr(scns, cls)(mesh, prm.Params(text, is_text=True))
    g = getattr(ifem.Fluid.MPI, cls)(tria, ifem.Parameters.AllParameters(text=text))
    g.setup()
    assert o.dofs.hanging_u and g.n_dofs == o.n
    return o, g


def _dirichlet(dim):
    full = 3 if dim == 2 else 7
    # inflow on id 0 (a hanging node of the 3-D band sits on it only through its masters), walls, open outflow
    return {0: (full, [1.0, 0.5, -0.25][:dim]), 2: (full, [0.0] * dim), 3: (full, [0.0] * dim)}


@pytest.mark.parametrize("cls", ["SCnsIM", "SUPGInsIM"])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("nonzero", [True, False])
def test_assembly_on_refined_mesh_matches_oracle(cls, dim, nonzero):
    o, g = _make(dim, cls, dirichlet=_dirichlet(dim), gravity=[1.0, -9.8, 0.5][:dim], neumann={1: 3.5})
    rng = np.random.default_rng(11)
    ev, pr = rng.uniform(-1, 1, o.n), rng.uniform(-1, 1, o.n)
    o.evaluation_point[:], o.present[:] = ev, pr
    g.set_vector(g.EVALUATION_POINT, ev)
    g.set_vector(g.PRESENT, pr)
    if cls == "SCnsIM":
        acc = rng.uniform(-1, 1, o.n)
        ind = (rng.uniform(size=o.mesh.n_cells) < 0.4).astype(np.int32)
        o.fsi_acceleration[:], o.indicator[:] = acc, ind
        g.set_vector(g.FSI_ACCELERATION, acc)
        g.set_indicator(ind)
    A_ref, b_ref = o.assemble(nonzero)
    g.assemble(nonzero)
    A, b = g.get_matrix(0), g.get_vector(g.SYSTEM_RHS)
    err_A = sp.linalg.norm(A - A_ref) / sp.linalg.norm(A_ref)
    assert err_A < 1e-12 and _rel(b, b_ref) < 1e-12, (err_A, _rel(b, b_ref))
    # rows and columns of the hanging dofs hold the diagonal only
    h = np.nonzero(o.dofs.is_hanging)[0]
    D = sp.diags(A.diagonal())
    assert abs(A.tocsr()[h] - D.tocsr()[h]).max() == 0.0 and abs(A.tocsc()[:, h] - D.tocsc()[:, h]).max() == 0.0


@pytest.mark.parametrize("dim", [2, 3])
def test_uniform_state_patch_test_on_device(dim):
    full = 3 if dim == 2 else 7
    o, g = _make(dim, dirichlet={i: (full, [0.0] * dim) for i in range(2 * dim)}, mu=0.7, rho=1.1)
    x = np.zeros(o.n)
    x[: o.n_u] = np.tile([0.3, -0.2, 0.45][:dim], o.dofs.n_unodes)
    x[o.n_u:] = 2.5
    g.set_vector(g.EVALUATION_POINT, x)
    g.set_vector(g.PRESENT, x)
    g.assemble(False)
    assert np.abs(g.get_vector(g.SYSTEM_RHS)).max() < 1e-13


@pytest.mark.parametrize("dim", [2, 3])
def test_time_steps_on_refined_mesh_match_oracle(dim):
    o, g = _make(dim, dirichlet=_dirichlet(dim), dt=1e-3, newton_tol=1e-9)
    g.set_control(fgmres_rel=1e-10)
    for k in range(3):
        o.run_one_step(k == 0)
        g.run_one_step(k == 0)
    sol = g.get_current_solution()
    assert _rel(sol[: o.n_u], o.velocity()) < 1e-6 and _rel(sol[o.n_u:], o.pressure()) < 1e-6
    # the solution satisfies the hanging-node lines
    from oracle import fem

    assert np.allclose(fem.distribute(o.dofs, sol.copy()), sol, rtol=0, atol=1e-12 * np.abs(sol).max())
    ho, hg = o.history, g.history()
    assert len(ho) == len(hg)
    for a, b in zip(ho, hg):
        assert abs(a[2] - b["abs_res"]) <= 1e-6 * max(a[2], 1e-12) + 1e-13


def test_hanging_node_properties_at_scale():
    """size-independent properties on a config-5-shaped mesh far beyond what the oracle handles in seconds (fsi-wall-3D box at three
    times the reference's resolution: 183 600 cells, 0.8 M dofs, 3 960 hanging nodes): a uniform state leaves a zero residual, the
    rows and columns of hanging dofs are diagonal after the condensation, and a time step returns a field that satisfies every
    hanging-node line"""
    import openifem_b200 as ifem

    scale = 3
    tria = ifem.Triangulation(3)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, (10 * scale, 10 * scale, 40 * scale), (0, 0, 0), (1, 1, 4), True)
    v, c, _ = tria.get_mesh()
    cz = v[c].mean(axis=1)[:, 2]
    tria.execute_refinement(((cz >= 2) & (cz <= 2.4)).astype(np.uint8))
    v, c, _ = tria.get_mesh()
    hv, hk, hm = tria.hanging()
    assert tria.n_active_cells() == 6800 * scale ** 3 and hv.size > 3000
    full = 7
    text = scns_prm(3, dirichlet={i: (full, [0.0] * 3) for i in range(6)}, mu=0.7, rho=1.1, dt=1e-3)
    g = ifem.Fluid.MPI.SCnsIM(tria, ifem.Parameters.AllParameters(text=text))
    g.setup()
    n_nodes = v.shape[0]
    assert g.n_dofs == 4 * n_nodes
    x = np.zeros(g.n_dofs)
    x[: 3 * n_nodes] = np.tile([0.3, -0.2, 0.45], n_nodes)
    x[3 * n_nodes:] = 2.5
    g.set_vector(g.EVALUATION_POINT, x)
    g.set_vector(g.PRESENT, x)
    g.assemble(False)
    assert np.abs(g.get_vector(g.SYSTEM_RHS)).max() < 1e-11
    # node of a vertex: lexicographic (z, y, x) order of the quantised positions (csrc/mesh.cpp spatial_renumber)
    lo, hi = v.min(axis=0), v.max(axis=0)
    q = np.rint((v - lo) / (hi - lo) * float((1 << 21) - 1)).astype(np.int64)
    order = np.lexsort((q[:, 0], q[:, 1], q[:, 2]))
    node = np.empty(n_nodes, dtype=np.int64)
    node[order] = np.arange(n_nodes)
    A = g.get_matrix(0).tocsr()
    hd = np.concatenate([3 * node[hv] + k for k in range(3)] + [3 * n_nodes + node[hv]])
    D = sp.diags(A.diagonal()).tocsr()
    assert abs(A[hd] - D[hd]).max() == 0.0 and abs(A.tocsc()[:, hd] - D.tocsc()[:, hd]).max() == 0.0 and (A.diagonal()[hd] > 0).all()
    # a time step from a non-trivial state: the update is distributed through the lines
    x0 = np.zeros(g.n_dofs)
    interior = np.all((v > lo + 1e-9) & (v < hi - 1e-9), axis=1)
    bump = np.where(interior, np.sin(np.pi * v[:, 0]) * np.sin(np.pi * v[:, 1]) * np.sin(np.pi * v[:, 2] / 4), 0.0)
    vel = np.zeros((n_nodes, 3))
    for k in (2, 4):  # the initial state has to satisfy the lines itself: the Newton updates are distributed, the state is not
        sel = hk == k
        bump[hv[sel]] = bump[hm[sel, :k]].mean(axis=1)
    vel[node, 0] = bump
    x0[: 3 * n_nodes] = vel.ravel()
    g.set_vector(g.PRESENT, x0)
    g.run_one_step(False)
    sol = g.get_current_solution()
    u = sol[: 3 * n_nodes].reshape(-1, 3)
    p = sol[3 * n_nodes:]
    for k in (2, 4):
        sel = hk == k
        m = node[hm[sel, :k]]
        assert np.abs(u[node[hv[sel]]] - u[m].mean(axis=1)).max() < 1e-12 * max(np.abs(u).max(), 1e-300) + 1e-15
        assert np.abs(p[node[hv[sel]]] - p[m].mean(axis=1)).max() < 1e-12 * max(np.abs(p).max(), 1e-300) + 1e-15
    assert np.abs(u).max() > 0.1
