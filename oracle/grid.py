"""ORACLE (test infrastructure, NOT product code).

General 2-D quadrilateral meshes for the oracle and the restatement of
Utils::GridCreator<2>::flow_around_cylinder (reference source/utilities.cpp:343-526),
the mesh of the reference's golden tests tests/fluid_cylinder_mpi (0.374235 / 46.5226)
and tests/fluid_cylinder_mpi_scnsim.

What is restated from deal.II (absent from this image, documented behaviour):
  * GridGenerator::hyper_cube_with_cylindrical_hole(tria, r_in, r_out) in 2-D: an 8-cell
    hyper_shell whose outer vertices at 45, 135, 225, 315 degrees are moved to the corners
    of the square [-r_out, r_out]^2 (the ones on the axes stay);
  * merge_triangulations with a vertex tolerance: coincident vertices keep the coordinates
    of the FIRST triangulation (here the Cartesian bulk mesh);
  * PolarManifold on the hole boundary: new points are linear in (r, theta) about the centre;
  * TransfiniteInterpolationManifold on the 8 cells around the hole: the vertices created by
    uniform refinement are the images of the dyadic points of the coarse cell's chart under
        x(s,t) = (1-s) L(t) + s R(t) + (1-t) B(s) + t T(s)
                 - [(1-s)(1-t) v0 + s(1-t) v1 + (1-s) t v2 + s t v3]
    with straight lines for the edges that are not on the hole and the polar arc for the one
    that is;
  * everything else is the flat manifold (midpoints; cell centres = mean of the corners),
    and FEValues uses MappingQ1, i.e. refined cells are straight-sided.
"""
from __future__ import annotations

import numpy as np

# local vertex order of a cell: lexicographic, v = x + 2*y (deal.II); face_no = 2*axis + side
_FACE_VERTS = {0: (0, 2), 1: (1, 3), 2: (0, 1), 3: (2, 3)}


class PolarArcChart:
    """Chart of one coarse cell next to the hole: corners v[0..3] (lexicographic), the edge t = 0 (v0 -> v1) is
    an arc about `centre`, the other three edges are straight."""

    def __init__(self, corners, centre):
        self.v = np.asarray(corners, dtype=np.float64)
        self.c = np.asarray(centre, dtype=np.float64)
        d0, d1 = self.v[0] - self.c, self.v[1] - self.c
        self.r0, self.r1 = np.hypot(*d0), np.hypot(*d1)
        self.a0, a1 = np.arctan2(d0[1], d0[0]), np.arctan2(d1[1], d1[0])
        # shortest way round (PolarManifold is periodic in theta)
        da = a1 - self.a0
        while da > np.pi:
            da -= 2 * np.pi
        while da < -np.pi:
            da += 2 * np.pi
        self.da = da

    def arc(self, s):
        r = (1 - s) * self.r0 + s * self.r1
        a = self.a0 + s * self.da
        return self.c + r * np.array([np.cos(a), np.sin(a)])

    def __call__(self, s, t):
        v = self.v
        B = self.arc(s)
        T = (1 - s) * v[2] + s * v[3]
        L = (1 - t) * v[0] + t * v[2]
        R = (1 - t) * v[1] + t * v[3]
        return ((1 - s) * L + s * R + (1 - t) * B + t * T
                - ((1 - s) * (1 - t) * v[0] + s * (1 - t) * v[1] + (1 - s) * t * v[2] + s * t * v[3]))


class QuadMesh:
    """vertices [nv][2], cells [nc][4] (lexicographic corner order, positively oriented), boundary_faces
    [(cell, face_no, boundary id)]. charts: optional (chart_of_cell [nc] int, -1 = flat; box [nc][4] =
    (s0, t0, s1, t1) of the cell inside its chart; list of chart callables)."""

    dim = 2

    def __init__(self, vertices, cells, boundary_faces, charts=None):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.boundary_faces = np.ascontiguousarray(boundary_faces, dtype=np.int32).reshape(-1, 3)
        self.n_cells = self.cells.shape[0]
        if charts is None:
            charts = (np.full(self.n_cells, -1, dtype=np.int32), np.zeros((self.n_cells, 4)), [])
        self.chart_of_cell, self.chart_box, self.charts = charts

    # ---- FE_Q(p) nodes: geometric entities, numbered in lexicographic (y, x) order of quantised position ----
    def _entities(self):
        """Q2 entity table: [nc][9] ids into (vertices | edges | cell centres) and their positions (MappingQ1
        support points: means of the corners they interpolate)."""
        nv, nc = self.vertices.shape[0], self.n_cells
        edge_id = {}
        tab = np.zeros((nc, 9), dtype=np.int64)
        # local Q2 node a = i + 3*j on the lattice {0,1,2}^2; corners it averages
        loc_corners = {}
        for j in range(3):
            for i in range(3):
                xs = (0,) if i == 0 else (1,) if i == 2 else (0, 1)
                ys = (0,) if j == 0 else (1,) if j == 2 else (0, 1)
                loc_corners[i + 3 * j] = [x + 2 * y for y in ys for x in xs]
        pos = [p for p in self.vertices]
        edge_pos = []
        for c in range(nc):
            cv = self.cells[c]
            for a in range(9):
                cs = loc_corners[a]
                if len(cs) == 1:
                    tab[c, a] = cv[cs[0]]
                elif len(cs) == 2:
                    key = tuple(sorted((int(cv[cs[0]]), int(cv[cs[1]]))))
                    if key not in edge_id:
                        edge_id[key] = len(edge_id)
                        edge_pos.append(0.5 * (self.vertices[key[0]] + self.vertices[key[1]]))
                    tab[c, a] = nv + edge_id[key]
                else:
                    tab[c, a] = -1 - c
        ne = len(edge_id)
        tab[tab < 0] = nv + ne + (-1 - tab[tab < 0])
        centres = self.vertices[self.cells].mean(axis=1)
        coords = np.concatenate([self.vertices, np.asarray(edge_pos).reshape(-1, 2), centres])
        return tab, coords

    @staticmethod
    def _spatial_renumber(tab, coords):
        lo, hi = coords.min(axis=0), coords.max(axis=0)
        ext = np.where(hi > lo, hi - lo, 1.0)
        Q = float((1 << 21) - 1)
        q = np.rint((coords - lo) / ext * Q).astype(np.int64)
        key = (q[:, 1] << 21) | q[:, 0]
        order = np.argsort(key, kind="stable")
        new_id = np.empty_like(order)
        new_id[order] = np.arange(order.size)
        return new_id[tab].astype(np.int32), coords[order]

    def node_table(self, p: int):
        if p == 1:
            tab, coords = self._spatial_renumber(self.cells.astype(np.int64), self.vertices)
        elif p == 2:
            tab, coords = self._spatial_renumber(*self._entities())
        else:
            raise ValueError("FE_Q(1) and FE_Q(2) only")
        return tab, coords.shape[0], coords

    # ---- uniform refinement; cells with a chart place their new vertices on it ----
    def refine_global(self, times: int = 1):
        m = self
        for _ in range(times):
            m = m._refine_once()
        return m

    def _refine_once(self):
        tab, coords = self._entities()
        coords = coords.copy()
        nc = self.n_cells
        # curved placement: local lattice point (i, j) of a chart cell -> chart((s0 + i/2 ds), (t0 + j/2 dt))
        for c in np.nonzero(self.chart_of_cell >= 0)[0]:
            ch = self.charts[self.chart_of_cell[c]]
            s0, t0, s1, t1 = self.chart_box[c]
            for j in range(3):
                for i in range(3):
                    if i != 1 and j != 1:
                        continue  # corners exist already
                    coords[tab[c, i + 3 * j]] = ch(s0 + 0.5 * i * (s1 - s0), t0 + 0.5 * j * (t1 - t0))
        new_cells = np.zeros((nc * 4, 4), dtype=np.int32)
        chart_of = np.repeat(self.chart_of_cell, 4)
        box = np.zeros((nc * 4, 4))
        for child in range(4):
            cx, cy = child & 1, child >> 1
            for v in range(4):
                vx, vy = v & 1, v >> 1
                new_cells[child::4, v] = tab[:, (cx + vx) + 3 * (cy + vy)]
            s0, t0, s1, t1 = self.chart_box.T
            ds, dt = 0.5 * (s1 - s0), 0.5 * (t1 - t0)
            box[child::4] = np.stack([s0 + cx * ds, t0 + cy * dt, s0 + (cx + 1) * ds, t0 + (cy + 1) * dt], axis=1)
        bf = []
        for (c, face, bid) in self.boundary_faces:
            axis, side = face // 2, face % 2
            for child in range(4):
                if ((child >> axis) & 1) == side:
                    bf.append((4 * c + child, face, bid))
        return QuadMesh(coords, new_cells, np.asarray(bf, dtype=np.int32), (chart_of, box, self.charts))


def _orient(vertices, quad):
    """corner ids of a quad in any cyclic order -> lexicographic order (v0, v1, v2, v3) with positive area"""
    p = vertices[list(quad)]
    area = 0.0
    for k in range(4):
        a, b = p[k], p[(k + 1) % 4]
        area += a[0] * b[1] - a[1] * b[0]
    q = list(quad) if area > 0 else list(quad)[::-1]
    return [q[0], q[1], q[3], q[2]]


def flow_around_cylinder_2d(compute_in_2d=True):
    """Utils::GridCreator<dim>::flow_around_cylinder_2d + the boundary ids of GridCreator<2>::flow_around_cylinder
    (utilities.cpp:343-526): channel [left, 2.2] x [0, 0.41] with a hole of diameter 0.1 about (0.2, 0.2).
    Boundary ids: 0 inflow (x = left), 1 outflow (x = 2.2), 2 y = 0, 3 y = 0.41, 4 cylinder."""
    left = 0.0 if compute_in_2d else -0.3
    nx, ny = (22 if compute_in_2d else 25), 4
    xs = np.linspace(left, 2.2, nx + 1)
    ys = np.linspace(0.0, 0.41, ny + 1)
    vid = {}
    verts = []

    def vertex(i, j):
        if (i, j) not in vid:
            vid[(i, j)] = len(verts)
            verts.append((xs[i], ys[j]))
        return vid[(i, j)]

    cells = []
    removed = []
    for j in range(ny):
        for i in range(nx):
            centre = np.array([0.5 * (xs[i] + xs[i + 1]), 0.5 * (ys[j] + ys[j + 1])])
            if np.linalg.norm(centre - np.array([0.2, 0.2])) < 0.15:
                removed.append((i, j))
                continue
            cells.append([vertex(i, j), vertex(i + 1, j), vertex(i, j + 1), vertex(i + 1, j + 1)])
    # the hole left by the removed cells is the square spanned by lattice points i0..i0+2, j0..j0+2
    i0, j0 = min(i for i, _ in removed), min(j for _, j in removed)
    assert sorted(removed) == sorted((i0 + a, j0 + b) for a in range(2) for b in range(2))
    n_bulk = len(cells)
    # outer ring of hyper_cube_with_cylindrical_hole, merged onto the bulk vertices: angle k*45 degrees -> lattice point
    ring = [(2, 1), (2, 2), (1, 2), (0, 2), (0, 1), (0, 0), (1, 0), (2, 0)]
    outer = [vertex(i0 + a, j0 + b) for a, b in ring]
    # inner ring: radius 0.05 about the centre of the cylinder triangulation, then re-centred at (0.2, 0.2)
    # (utilities.cpp:452-480: the mean of the circle's vertices is moved to (0.2, 0.2))
    centre = np.array([0.2, 0.2])
    inner = []
    for k in range(8):
        a = 2 * np.pi * k / 8
        inner.append(len(verts))
        verts.append((centre[0] + 0.05 * np.cos(a), centre[1] + 0.05 * np.sin(a)))
    verts = np.asarray(verts, dtype=np.float64)
    charts = []
    chart_of = [-1] * n_bulk
    box = [(0.0, 0.0, 0.0, 0.0)] * n_bulk
    bfaces = []
    for k in range(8):
        k1 = (k + 1) % 8
        # local order chosen so that the arc is the edge t = 0 (v0 -> v1) and the cell is positively oriented:
        # going round anticlockwise: inner_k1 -> inner_k is clockwise on the circle, so v0 = inner_k1, v1 = inner_k,
        # v2 = outer_k1, v3 = outer_k
        cell = [inner[k1], inner[k], outer[k1], outer[k]]
        p = verts[cell]
        area = 0.5 * ((p[1][0] - p[0][0]) * (p[2][1] - p[0][1]) - (p[2][0] - p[0][0]) * (p[1][1] - p[0][1]))
        if area < 0:
            cell = [inner[k], inner[k1], outer[k], outer[k1]]
        cells.append(cell)
        charts.append(PolarArcChart(verts[cell], centre))
        chart_of.append(k)
        box.append((0.0, 0.0, 1.0, 1.0))
        bfaces.append((len(cells) - 1, 2, 4))  # the arc is local face 2 (y = 0 side of the chart)
    cells = np.asarray(cells, dtype=np.int32)
    # boundary faces of the bulk: faces that belong to exactly one cell
    count = {}
    for c, cv in enumerate(cells):
        for f, (a, b) in _FACE_VERTS.items():
            count.setdefault(tuple(sorted((int(cv[a]), int(cv[b])))), []).append((c, f))
    for key, owners in count.items():
        if len(owners) != 1:
            continue
        c, f = owners[0]
        if c >= n_bulk:
            continue  # arcs were added above (the other faces of the ring cells are interior)
        mid = 0.5 * (verts[key[0]] + verts[key[1]])
        if abs(mid[0] - 2.2) < 1e-12:
            bid = 1
        elif abs(mid[0] - left) < 1e-12:
            bid = 0
        elif abs(mid[1] - 0.41) < 1e-12:
            bid = 3
        elif abs(mid[1]) < 1e-12:
            bid = 2
        else:
            bid = 4
        bfaces.append((c, f, bid))
    return QuadMesh(verts, cells, np.asarray(bfaces, dtype=np.int32),
                    (np.asarray(chart_of, dtype=np.int32), np.asarray(box, dtype=np.float64), charts))


class HexMesh:
    """Unstructured hexahedral mesh from plain arrays (test infrastructure): vertices [nv][3], cells [nc][8] in lexicographic corner
    order, boundary_faces [(cell, face_no = 2 * axis + side, boundary id)]. FE_Q(p) nodes are the geometric entities of the cells
    (vertices, edge / face / cell midpoints = means of the corners they interpolate: the support points of the default Q1 mapping,
    reference source/mpi_insim.cpp:167), numbered in lexicographic (z, y, x) order of their quantised position like QuadMesh."""

    dim = 3

    def __init__(self, vertices, cells, boundary_faces):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.boundary_faces = np.ascontiguousarray(boundary_faces, dtype=np.int32).reshape(-1, 3)
        self.n_cells = self.cells.shape[0]

    def _entities(self):
        nc = self.n_cells
        # local Q2 node (i, j, k) on {0,1,2}^3 -> the corners it averages
        loc = []
        for k in range(3):
            for j in range(3):
                for i in range(3):
                    sel = [(0,) if t == 0 else (1,) if t == 2 else (0, 1) for t in (i, j, k)]
                    loc.append([x + 2 * y + 4 * z for z in sel[2] for y in sel[1] for x in sel[0]])
        ids, pos = {}, []
        tab = np.zeros((nc, 27), dtype=np.int64)
        for c in range(nc):
            cv = self.cells[c]
            for a in range(27):
                key = tuple(sorted(int(cv[v]) for v in loc[a]))
                n = ids.get(key)
                if n is None:
                    n = ids[key] = len(pos)
                    pos.append(self.vertices[list(key)].mean(axis=0))
                tab[c, a] = n
        return tab, np.asarray(pos)

    @staticmethod
    def _spatial_renumber(tab, coords):
        lo, hi = coords.min(axis=0), coords.max(axis=0)
        ext = np.where(hi > lo, hi - lo, 1.0)
        Q = float((1 << 21) - 1)
        q = np.rint((coords - lo) / ext * Q).astype(np.int64)
        key = (q[:, 2] << 42) | (q[:, 1] << 21) | q[:, 0]
        order = np.argsort(key, kind="stable")
        new_id = np.empty_like(order)
        new_id[order] = np.arange(order.size)
        return new_id[tab].astype(np.int32), coords[order]

    def node_table(self, p: int):
        if p == 1:
            tab, coords = self._spatial_renumber(self.cells.astype(np.int64), self.vertices)
        elif p == 2:
            tab, coords = self._spatial_renumber(*self._entities())
        else:
            raise ValueError("FE_Q(1) and FE_Q(2) only")
        return tab, coords.shape[0], coords
