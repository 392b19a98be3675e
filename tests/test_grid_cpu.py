"""Host-side mesh generation of the product (no GPU needed): Utils::GridCreator<dim>::flow_around_cylinder
(reference source/utilities.cpp:343-574) and the manifold-aware refine_global against the oracle's restatement
(oracle/grid.py), which reproduces the reference goldens of tests/fluid_cylinder_mpi* on that mesh."""
import numpy as np
import pytest

import openifem_b200 as ifem
from oracle import grid


def _sorted(a):
    return a[np.lexsort(a.T[::-1])]


@pytest.mark.parametrize("level", [0, 1, 3])
def test_cylinder_mesh_2d_matches_oracle(level):
    t = ifem.Triangulation(2)
    ifem.GridCreator.flow_around_cylinder(t)
    t.refine_global(level)
    v, c, b = t.get_mesh()
    m = grid.flow_around_cylinder_2d(True).refine_global(level)
    assert c.shape == m.cells.shape == (92 * 4 ** level, 4)
    assert np.abs(_sorted(v) - _sorted(m.vertices)).max() < 1e-14
    assert np.abs(_sorted(v[c].mean(axis=1)) - _sorted(m.vertices[m.cells].mean(axis=1))).max() < 1e-14
    assert sorted(zip(*np.unique(b[:, 2], return_counts=True))) == sorted(zip(*np.unique(m.boundary_faces[:, 2], return_counts=True)))
    # positively oriented cells, area of the channel minus the polygonal hole, vertices of the cylinder ON the circle
    p = v[c]
    area = 0.5 * ((p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 2, 0] - p[:, 0, 0]) * (p[:, 1, 1] - p[:, 0, 1]))
    assert area.min() > 0
    on_cyl = np.unique(np.concatenate([c[cell, [0, 1]] for cell, face, bid in b if bid == 4]))
    assert len(on_cyl) == 8 * 2 ** level
    assert np.abs(np.hypot(*(v[on_cyl] - 0.2).T) - 0.05).max() < 1e-14


def test_cylinder_mesh_3d_extruded():
    """GridCreator<3>: the 2-D mesh with left = -0.3 (25 x 4 bulk cells) extruded in 8 layers, boundary ids 0..6"""
    t = ifem.Triangulation(3)
    ifem.GridCreator.flow_around_cylinder(t)
    v, c, b = t.get_mesh()
    assert c.shape == ((25 * 4 - 4 + 8) * 8, 8)
    assert v[:, 0].min() == -0.3 and v[:, 0].max() == 2.2 and v[:, 2].max() == 0.41
    ids = dict(zip(*np.unique(b[:, 2], return_counts=True)))
    assert ids == {0: 4 * 8, 1: 4 * 8, 2: 25 * 8, 3: 25 * 8, 4: 104, 5: 104, 6: 8 * 8}
    # positive Jacobian at the first corner of every cell
    p = v[c]
    e1, e2, e3 = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0], p[:, 4] - p[:, 0]
    assert np.einsum("ij,ij->i", np.cross(e1, e2), e3).min() > 0
    t.refine_global(1)
    assert t.n_active_cells() == 832 * 8


def test_mesh_import_through_the_abi_round_trips_and_finds_hanging_vertices():
    """ifem_tria_set_mesh: a host that owns its triangulation (deal.II in the reference) hands over plain arrays; a locally refined
    mesh keeps its hanging vertices; inverted cells are refused"""
    import numpy as np

    import openifem_b200 as ifem

    a = ifem.Triangulation(3)
    ifem.GridGenerator.subdivided_hyper_rectangle(a, (3, 2, 4), (0, 0, 0), (1.5, 1.0, 2.0), True)
    v, c, _ = a.get_mesh()
    a.execute_refinement((v[c].mean(axis=1)[:, 2] > 1.0).astype(np.uint8))
    v, c, f = a.get_mesh()
    b = ifem.Triangulation(3)
    b.set_mesh(v, c, f)
    v2, c2, f2 = b.get_mesh()
    assert np.array_equal(v, v2) and np.array_equal(c, c2) and np.array_equal(f, f2)
    ha, hb = a.hanging(), b.hanging()
    assert ha[0].size > 0 and all(np.array_equal(x, y) for x, y in zip(ha, hb))
    b.refine_global(1)
    assert b.n_active_cells() == 8 * a.n_active_cells()
    bad = c.copy()
    bad[0, [0, 1]] = bad[0, [1, 0]]
    import pytest

    with pytest.raises(RuntimeError, match="inverted"):
        ifem.Triangulation(3).set_mesh(v, bad, f)


def _cell_volumes(v, c):
    dim = v.shape[1]
    e = [v[c[:, 1 << d]] - v[c[:, 0]] for d in range(dim)]
    return np.abs(np.linalg.det(np.stack(e, axis=1)))


import numpy as np  # noqa: E402
import pytest  # noqa: E402


@pytest.mark.parametrize("dim", [2, 3])
def test_coarsening_and_refinement_keeps_a_balanced_mesh_and_transfers_linear_fields(dim):
    """execute_coarsening_and_refinement (FSI::refine_mesh's mesh operation): random refine / coarsen flags over several rounds -
    the mesh stays a partition of the box, neighbours differ by one level at most (find_hanging_vertices would throw), the
    transfer plan reproduces a linear field, and coarsening everything leads back to the coarse mesh"""
    import openifem_b200 as ifem

    rng = np.random.default_rng(4)
    t = ifem.Triangulation(dim)
    reps = (5, 4) if dim == 2 else (3, 2, 3)
    hi = (2.5, 2.0) if dim == 2 else (1.5, 1.0, 1.5)
    ifem.GridGenerator.subdivided_hyper_rectangle(t, reps, (0,) * dim, hi, True)
    vol = float(np.prod(hi))
    n0 = t.n_active_cells()
    lin = lambda x: 0.3 + x @ np.array([0.8, -0.4, 0.55][:dim])
    v, c, b = t.get_mesh()
    area0 = {}
    for k in range(6):
        nc = t.n_active_cells()
        refine = rng.uniform(size=nc) < (0.25 if k < 4 else 0.05)
        coarsen = (rng.uniform(size=nc) < 0.5) & ~refine
        field = lin(v)
        ptr, old, w = t.execute_coarsening_and_refinement(refine, coarsen)
        v, c, b = t.get_mesh()
        assert abs(_cell_volumes(v, c).sum() - vol) < 1e-12 * vol
        new_field = np.array([np.dot(w[ptr[i]:ptr[i + 1]], field[old[ptr[i]:ptr[i + 1]]]) for i in range(v.shape[0])])
        assert np.abs(new_field - lin(v)).max() < 1e-13
        lv = t.levels()
        assert lv.min() >= 0 and lv.max() <= k + 1
        # boundary faces still cover the boundary: total boundary measure is that of the box
        meas = 0.0
        for (cell, face, _id) in b:
            axis, side = face // 2, face % 2
            cv = v[c[cell]]
            ext = cv.max(axis=0) - cv.min(axis=0)
            meas += np.prod(np.delete(ext, axis))
        expect = 2 * sum(np.prod(np.delete(np.array(hi), a)) for a in range(dim))
        assert abs(meas - expect) < 1e-12 * expect
        hv, hk, hm = t.hanging()
        assert (lv.max() == lv.min()) == (hv.size == 0)
    assert t.levels().max() >= 2
    for _ in range(8):  # coarsen everything: back to the coarse mesh, cell for cell
        nc = t.n_active_cells()
        t.execute_coarsening_and_refinement(np.zeros(nc, bool), np.ones(nc, bool))
    v2, c2, b2 = t.get_mesh()
    assert t.n_active_cells() == n0 and t.levels().max() == 0 and t.hanging()[0].size == 0
    assert abs(_cell_volumes(v2, c2).sum() - vol) < 1e-12 * vol and v2.shape[0] == int(np.prod([r + 1 for r in reps]))
