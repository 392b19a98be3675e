"""Self-consistency of the FSI oracle (oracle/fsi.py) on CPU: the two point_in_solid algorithms of the
reference (2-D crossing number, mpi_fsi.cpp:154-215; cell-wise point_inside, :216-222) agree away from the
boundary, indicator counts match the analytic overlap, and Q1 interpolation reproduces linear fields."""
import numpy as np

from oracle import fem, fsi


def _solid(dim, reps, lo, hi, disp_fn=None):
    mesh = fem.BoxMesh(reps, lo, hi)
    u = np.zeros(mesh.vertices.size)
    if disp_fn is not None:
        u = disp_fn(mesh.vertices).ravel()
    return mesh, u, fsi.SolidGeometry(mesh, u)


def test_crossing_number_agrees_with_cell_search_2d():
    shear = lambda X: np.stack([0.15 * X[:, 1] ** 2, 0.05 * X[:, 0]], axis=1)
    mesh, u, s = _solid(2, (5, 7), (0.3, 0.2), (0.6, 0.9), shear)
    rng = np.random.default_rng(3)
    pts = rng.uniform(0.0, 1.1, size=(400, 2))
    n_in = 0
    for p in pts:
        a = s.in_box(p) and fsi.point_in_solid_2d(p, s.box, s.segments)
        b = s.locate(p)[0] is not None
        assert a == b
        n_in += a
    assert 40 < n_in < 200
    # vertices of the solid itself are "inside" (on-vertex rule)
    for v in s.x[::5]:
        assert s.point_in_solid(v)


def test_indicator_counts_box_overlap():
    # binary-representable coordinates: the reference's box / crossing tests are exact comparisons
    fluid = fem.BoxMesh((8, 8), (0, 0), (1, 1))
    mesh, u, s = _solid(2, (3, 3), (0.25, 0.25), (0.75, 0.75))
    ind = fsi.update_indicator(fluid, s)
    # fluid cells inside [0.25,0.75]^2 on a 1/8 grid: 4 x 4 cells fully covered (closed-set test on the boundary)
    assert ind.sum() == 16
    fluid3 = fem.BoxMesh((8, 8, 8), (0, 0, 0), (1, 1, 1))
    mesh3, u3, s3 = _solid(3, (2, 2, 2), (0.25, 0.25, 0.25), (0.75, 0.75, 0.5))
    ind3 = fsi.update_indicator(fluid3, s3)
    assert ind3.sum() == 4 * 4 * 2


def test_q1_interpolation_reproduces_linear_fields():
    stretch = lambda X: np.stack([0.1 * X[:, 0], -0.05 * X[:, 1], 0.02 * X[:, 2]], axis=1)
    mesh, u, s = _solid(3, (3, 2, 2), (0, 0, 0), (1, 1, 1), stretch)
    A = np.array([[1.0, 2.0, -1.0], [0.5, 0.0, 3.0], [-2.0, 1.0, 0.25]])
    field = (s.x @ A.T + np.array([0.3, -0.2, 0.1])).ravel()
    rng = np.random.default_rng(5)
    for p in rng.uniform(0.05, 0.95, size=(20, 3)) * np.array([1.1, 0.95, 1.02]):
        val = s.interpolate(field, p)
        assert val is not None
        assert np.allclose(val, A @ p + np.array([0.3, -0.2, 0.1]), atol=1e-11)


def test_refine_mesh_flags_and_transfer_of_the_oracle():
    """oracle/fsi.py FSI.refine_mesh (mpi_fsi.cpp:1024-1117) on its own: cells within one diameter of the solid boundary reach
    level 2 after the two calls FSI::run makes, cells far away stay coarse, the levels are capped, and the transfer (old FE field
    evaluated at the new support points) reproduces a bilinear field exactly"""
    import openifem_b200 as ifem  # host-side mesh class only (no device work)
    from oracle import grid, prm, scns, solid
    from test_fsi_gpu import _fsi_text

    P = prm.Params(_fsi_text(2), is_text=True)
    tria = ifem.Triangulation(2)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, (12, 12), (0.0, 0.0), (1.0, 1.0), True)
    v, c, b = tria.get_mesh()
    o_fluid = scns.SCnsIM(grid.QuadMesh(v, c, b), P)
    o_solid = solid.HyperElasticity(fem.BoxMesh((4, 6), (0.3125, 0.0), (0.5625, 0.6875)), P)
    loop = fsi.FSI(o_fluid, o_solid, True)
    f = lambda X: 0.3 + 0.8 * X[:, 0] - 0.4 * X[:, 1] + 0.25 * X[:, 0] * X[:, 1]  # bilinear: in the Q1 space of every mesh
    d = o_fluid.dofs
    o_fluid.present[0: o_fluid.n_u: 2] = f(d.ucoords)
    o_fluid.present[o_fluid.n_u:] = -f(d.pcoords)
    for _ in range(2):
        new = loop.refine_mesh(tria, 0, 3)
    lv = tria.levels()
    assert lv.max() == 2 and lv.min() == 0 and new.mesh.n_cells == tria.n_active_cells() > 144
    X = new.mesh.vertices[new.mesh.cells].mean(axis=1)
    far = np.hypot(X[:, 0] - 0.9, X[:, 1] - 0.9) < 0.08
    assert far.any() and (lv[far] == 0).all()
    near = (np.abs(X[:, 0] - 0.3125) < 0.02) & (X[:, 1] < 0.6)
    assert near.any() and (lv[near] == 2).all()
    nd = new.dofs
    assert np.abs(new.present[0: new.n_u: 2] - f(nd.ucoords)).max() < 1e-13 and np.abs(new.present[new.n_u:] + f(nd.pcoords)).max() < 1e-13
    assert len(nd.hanging_u) > 0
    for _ in range(3):  # never beyond Global refinements + 3
        new = loop.refine_mesh(tria, 0, 3)
    assert tria.levels().max() == 3
