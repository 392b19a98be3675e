"""ORACLE (test infrastructure, NOT product code).

CPU restatement of the deal.II pieces the OpenIFEM hot path relies on, for
box meshes: Lagrange Q_p tensor-product shape functions on [0,1]^d with
equidistant support points (FE_Q, p<=2), tensor Gauss-Legendre quadrature
(QGauss), Q1 geometry mapping (FEValues without mapping argument,
reference mpi_insim.cpp:167), colorized boundary ids (2*axis+side,
GridGenerator::subdivided_hyper_rectangle(..., colorize=true)), node / DoF
numbering into [u | p] blocks (reference mpi_fluid_solver.cpp:116-162; the
permutation differs from Cuthill-McKee, results compared are permutation
invariant) and Dirichlet constraints with "first boundary id wins"
(reference mpi_fluid_solver.cpp:165-280, VectorTools::interpolate_boundary_values
never overwrites an existing constraint line).

deal.II itself is not vendored in /root/reference and is absent from this
image ("parity unpinned" for the third-party semantics restated here; they are
pinned indirectly through the reference's golden values, see
tests/test_oracle_goldens.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this package.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------
# 1-D building blocks
# ----------------------------------------------------------------------------
def gauss_legendre_01(n: int):
    """QGauss<1>(n) on [0,1]."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange_1d(p: int, x):
    """Values and derivatives of the p+1 equidistant Lagrange polynomials."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    nodes = np.linspace(0.0, 1.0, p + 1)
    V = np.ones((x.size, p + 1))
    D = np.zeros((x.size, p + 1))
    for i in range(p + 1):
        for j in range(p + 1):
            if j != i:
                V[:, i] *= (x - nodes[j]) / (nodes[i] - nodes[j])
        for k in range(p + 1):
            if k == i:
                continue
            t = np.ones_like(x) / (nodes[i] - nodes[k])
            for j in range(p + 1):
                if j != i and j != k:
                    t *= (x - nodes[j]) / (nodes[i] - nodes[j])
            D[:, i] += t
    return V, D


class FEQ:
    """Scalar FE_Q(p) on [0,1]^dim, local nodes lexicographic (x fastest)."""

    def __init__(self, dim: int, p: int):
        self.dim, self.p = dim, p
        self.n1 = p + 1
        self.n = self.n1 ** dim
        idx = np.indices((self.n1,) * dim).reshape(dim, -1)[::-1].T  # x fastest
        self.lattice = idx  # [n][dim] integer lattice coordinate of each node
        self.unit_points = idx / float(p)

    def eval(self, pts):
        """pts [m][dim] -> N [m][n], dN [m][n][dim] (reference gradients)."""
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, self.dim)
        m = pts.shape[0]
        V1 = []
        D1 = []
        for d in range(self.dim):
            v, dv = lagrange_1d(self.p, pts[:, d])
            V1.append(v)
            D1.append(dv)
        N = np.ones((m, self.n))
        dN = np.ones((m, self.n, self.dim))
        for a in range(self.n):
            for d in range(self.dim):
                ia = self.lattice[a, d]
                N[:, a] *= V1[d][:, ia]
                for e in range(self.dim):
                    dN[:, a, e] *= D1[d][:, ia] if e == d else V1[d][:, ia]
        return N, dN


def qgauss(dim: int, n: int):
    """Tensor QGauss<dim>(n): points [nq][dim] (x fastest), weights [nq]."""
    x, w = gauss_legendre_01(n)
    idx = np.indices((n,) * dim).reshape(dim, -1)[::-1].T
    pts = x[idx]
    wts = np.prod(w[idx], axis=1)
    return pts, wts


# ----------------------------------------------------------------------------
# Box mesh (GridGenerator::subdivided_hyper_rectangle, colorize = true)
# ----------------------------------------------------------------------------
class BoxMesh:
    def __init__(self, subdivisions, lo, hi):
        self.n = tuple(int(s) for s in subdivisions)
        self.dim = len(self.n)
        self.lo = np.asarray(lo, dtype=np.float64)
        self.hi = np.asarray(hi, dtype=np.float64)
        dim, n = self.dim, self.n
        nv = tuple(k + 1 for k in n)
        grid = np.indices(nv[::-1]).reshape(dim, -1)[::-1].T  # x fastest
        self.vertices = self.lo + grid * ((self.hi - self.lo) / np.asarray(n))
        cgrid = np.indices(n[::-1]).reshape(dim, -1)[::-1].T
        self.cell_ijk = cgrid
        self.n_cells = cgrid.shape[0]
        corner = np.indices((2,) * dim).reshape(dim, -1)[::-1].T  # [2^d][dim]
        vstr = np.cumprod((1,) + nv[:-1])
        self.cells = ((cgrid[:, None, :] + corner[None, :, :]) * vstr).sum(-1).astype(np.int32)
        # boundary faces: (cell, face_no = 2*axis+side, boundary id = face_no)
        faces = []
        for axis in range(dim):
            for side in (0, 1):
                sel = np.nonzero(cgrid[:, axis] == (0 if side == 0 else n[axis] - 1))[0]
                for c in sel:
                    faces.append((c, 2 * axis + side, 2 * axis + side))
        self.boundary_faces = np.asarray(faces, dtype=np.int32).reshape(-1, 3)

    def refine_global(self, times: int):
        m = BoxMesh(tuple(k * 2 ** times for k in self.n), self.lo, self.hi)
        return m

    def node_table(self, p: int):
        """cell -> global node ids for FE_Q(p); nodes lexicographic on the
        (p*n+1)^dim lattice. Returns (table [nc][(p+1)^d] int32, n_nodes,
        node coordinates [n_nodes][dim])."""
        dim, n = self.dim, self.n
        nn = tuple(p * k + 1 for k in n)
        nstr = np.cumprod((1,) + nn[:-1])
        loc = np.indices((p + 1,) * dim).reshape(dim, -1)[::-1].T
        tab = ((self.cell_ijk[:, None, :] * p + loc[None, :, :]) * nstr).sum(-1).astype(np.int32)
        grid = np.indices(nn[::-1]).reshape(dim, -1)[::-1].T
        coords = self.lo + grid * ((self.hi - self.lo) / (np.asarray(n) * p))
        return tab, int(np.prod(nn)), coords


def face_local_nodes(dim: int, p: int, face_no: int):
    """Local FE_Q(p) node indices lying on face 2*axis+side."""
    fe = FEQ(dim, p)
    axis, side = face_no // 2, face_no % 2
    return np.nonzero(fe.lattice[:, axis] == (0 if side == 0 else p))[0]


# ----------------------------------------------------------------------------
# Fluid DoF layout and Dirichlet constraints
# ----------------------------------------------------------------------------
class FluidDofs:
    """[u | p] block layout: u dof = dim*node + c, p dof = dim*Nu + pnode.
    Local dof order in a cell: velocity (a*dim + c), then pressure nodes."""

    def __init__(self, mesh: BoxMesh, pu: int, pp: int):
        self.mesh, self.pu, self.pp = mesh, pu, pp
        dim = mesh.dim
        self.dim = dim
        self.unodes, self.n_unodes, self.ucoords = mesh.node_table(pu)
        self.pnodes, self.n_pnodes, self.pcoords = mesh.node_table(pp)
        self.n_u = dim * self.n_unodes
        self.n_p = self.n_pnodes
        self.n_dofs = self.n_u + self.n_p
        nu_loc = self.unodes.shape[1]
        udofs = (self.unodes[:, :, None] * dim + np.arange(dim)[None, None, :]).reshape(mesh.n_cells, nu_loc * dim)
        pdofs = self.n_u + self.pnodes
        self.cell_dofs = np.concatenate([udofs, pdofs], axis=1).astype(np.int32)
        self.dofs_per_cell = self.cell_dofs.shape[1]
        # hanging nodes of a locally refined mesh (DoFTools::make_hanging_node_constraints, mpi_fluid_solver.cpp:182-184)
        self.hanging_u, self.hanging_p = {}, {}
        if not isinstance(mesh, BoxMesh):  # a box mesh is uniform by construction
            if pu == 1 and pp == 1:
                self.hanging_u = hanging_nodes(self.unodes, self.ucoords)
                self.hanging_p = hanging_nodes(self.pnodes, self.pcoords)
            elif hanging_nodes(*mesh.node_table(1)[::2]):
                raise NotImplementedError("hanging-node constraints are restated for FE_Q(1) velocity and pressure only")
        self.hanging_dofs = {}  # dof -> (master dofs, weights), same component as the slave
        for h, (ms, w) in self.hanging_u.items():
            for c in range(dim):
                self.hanging_dofs[dim * h + c] = (dim * ms + c, w)
        for h, (ms, w) in self.hanging_p.items():
            self.hanging_dofs[self.n_u + h] = (self.n_u + ms, w)
        self.is_hanging = np.zeros(self.n_dofs, dtype=bool)
        self.is_hanging[list(self.hanging_dofs)] = True

    def support_points(self):
        pts = np.zeros((self.n_dofs, self.dim))
        pts[: self.n_u] = np.repeat(self.ucoords, self.dim, axis=0)
        pts[self.n_u:] = self.pcoords
        return pts


def hanging_nodes(tab, coords):
    """Hanging nodes of FE_Q(1) on a mesh with at most one level of difference between neighbouring cells, found from the
    geometry alone: a node that sits at the midpoint of another cell's edge is constrained to the mean of the two edge ends,
    one at the centre of another cell's face (3-D) to the mean of the four face corners - the constraint lines
    DoFTools::make_hanging_node_constraints writes for Q1 (reference source/mpi_fluid_solver.cpp:182-184). Returns
    {node: (masters, weight)}; tab is the FE_Q(1) node table. FE_Q(2) on locally refined meshes is not restated."""
    dim = coords.shape[1]
    lo, hi = coords.min(axis=0), coords.max(axis=0)
    ext = np.where(hi > lo, hi - lo, 1.0)

    def keys(x):  # half steps of a 2^20 lattice: a midpoint never rounds onto an end of its edge
        q = np.rint((x - lo) / ext * float((1 << 21) - 2)).astype(np.int64)
        k = np.zeros(q.shape[0], dtype=np.int64)
        for d in range(dim - 1, -1, -1):
            k = (k << 21) | q[:, d]
        return k

    node_key = keys(coords)
    order = np.argsort(node_key)
    sorted_key = node_key[order]
    out = {}
    groups = []
    for v in range(1 << dim):
        for d in range(dim):
            if not (v >> d) & 1:
                groups.append((v, v | (1 << d)))
    if dim == 3:
        for axis in range(3):
            for side in (0, 1):
                groups.append(tuple(v for v in range(8) if ((v >> axis) & 1) == side))
    for g in groups:
        ids = tab[:, list(g)]  # [nc][2 or 4]
        k = keys(coords[ids].mean(axis=1))
        pos = np.minimum(np.searchsorted(sorted_key, k), sorted_key.size - 1)
        hit = sorted_key[pos] == k
        for c in np.nonzero(hit)[0]:
            h = int(order[pos[c]])
            if h in ids[c]:
                continue
            if h not in out or len(out[h][0]) < ids.shape[1]:
                out[h] = (np.sort(ids[c]).astype(np.int64), 1.0 / ids.shape[1])
    for h, (ms, _) in out.items():
        assert not any(int(m) in out for m in ms), "a hanging node depends on another hanging node"
    return out


def resolve_constraints(dofs, con, val):
    """AffineConstraints::close() for the lines of one constraint object: hanging-node lines whose masters carry a Dirichlet
    value have that master replaced by its value (it adds weight * value to the line's inhomogeneity). Returns
    (flag [n_dofs]: 0 free, 1 line without masters, 2 line with masters; inhomogeneity [n_dofs]; ptr, master, weight of
    the lines in CSR form over all dofs)."""
    con2, val2 = con.copy(), val.copy()
    ptr = np.zeros(dofs.n_dofs + 1, dtype=np.int64)
    masters, weights = {}, {}
    for g, (ms, w) in dofs.hanging_dofs.items():
        assert not con[g], "a hanging dof carries its hanging-node line only (interpolate_boundary_values skips constrained dofs)"
        keep = [int(m) for m in ms if not con[m]]
        val2[g] = sum(w * val[m] for m in ms if con[m])
        con2[g] = 2 if keep else 1
        masters[g], weights[g] = keep, [w] * len(keep)
        ptr[g + 1] = len(keep)
    ptr = np.cumsum(ptr)
    master = np.zeros(max(1, int(ptr[-1])), dtype=np.int32)
    weight = np.zeros(max(1, int(ptr[-1])))
    for g in masters:
        master[ptr[g]:ptr[g + 1]] = masters[g]
        weight[ptr[g]:ptr[g + 1]] = weights[g]
    return con2, val2, ptr, master, weight


def distribute(dofs, x):
    """AffineConstraints::distribute for the hanging-node lines, after the Dirichlet entries of x were set"""
    for g, (ms, w) in dofs.hanging_dofs.items():
        x[g] = w * x[ms].sum()
    return x


def component_mask(flag: int, dim: int):
    """1-x 2-y 3-xy 4-z 5-xz 6-yz 7-xyz (mpi_fluid_solver.cpp:199-243)."""
    return [c for c in range(dim) if flag & (1 << c)]


def make_dirichlet_constraints(dofs: FluidDofs, dirichlet_bcs: dict, hard_coded=None, time=0.0):
    """dirichlet_bcs: {boundary id: (flag, [values])} visited in ascending id
    order; first constraint on a dof wins. Returns (flag[n_dofs] uint8,
    nonzero values[n_dofs]); zero_constraints share the flags with value 0.
    hard_coded: {id: f(point, component, time)} overrides the constant values."""
    mesh, dim, pu = dofs.mesh, dofs.dim, dofs.pu
    con = np.zeros(dofs.n_dofs, dtype=np.uint8)
    val = np.zeros(dofs.n_dofs)
    for bid in sorted(dirichlet_bcs):
        flag, values = dirichlet_bcs[bid]
        comps = component_mask(flag, dim)
        aug = np.zeros(dim)
        for k, c in enumerate(comps):
            aug[c] = values[k]
        for (cell, face_no, fid) in mesh.boundary_faces:
            if fid != bid:
                continue
            for a in face_local_nodes(dim, pu, face_no):
                node = dofs.unodes[cell, a]
                for c in comps:
                    g = dim * node + c
                    if con[g] or dofs.is_hanging[g]:  # an existing line (also a hanging-node one) is never overwritten
                        continue
                    con[g] = 1
                    if hard_coded is not None and bid in hard_coded:
                        val[g] = hard_coded[bid](dofs.ucoords[node], c, time)
                    else:
                        val[g] = aug[c]
    return con, val


def full_pattern(cell_dofs, n_dofs, hanging_dofs=None):
    """DoFTools::make_sparsity_pattern without coupling table: every dof of a
    cell couples with every other (mpi_fluid_solver.cpp:311-312); with hanging-node lines the masters of a cell's
    hanging dofs couple with the cell's dofs and with each other as well. CSR, sorted."""
    import scipy.sparse as sp

    nc, k = cell_dofs.shape
    rows = np.repeat(cell_dofs, k, axis=1).ravel()
    cols = np.tile(cell_dofs, (1, k)).ravel()
    if hanging_dofs:
        er, ec = [], []
        for cd in cell_dofs:
            extra = [int(m) for g in cd if int(g) in hanging_dofs for m in hanging_dofs[int(g)][0]]
            if not extra:
                continue
            full = np.unique(np.concatenate([cd, extra]))
            er.append(np.repeat(full, full.size))
            ec.append(np.tile(full, full.size))
        if er:
            rows = np.concatenate([rows] + er)
            cols = np.concatenate([cols] + ec)
    A = sp.coo_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(n_dofs, n_dofs)).tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32)
