"""ORACLE (test infrastructure, NOT product code).

CPU restatement of the immersed-coupling hot path of MPI::FSI (reference source/mpi_fsi.cpp):
  move_solid_mesh      :40-75    deformed solid vertices x = X + u (Q1 solid: vertices are the nodes)
  update_solid_box     :95-119   bounding box of the deformed solid
  point_in_solid       :143-224  bbox reject; 2-D crossing number over boundary segments with the
                                 on-edge / on-vertex special cases; 3-D "any solid cell contains it"
  update_indicator     :292-319  indicator = 1 iff all 2^dim vertices of a fluid cell are in the solid
  find_fluid_bc        :324-663  fsi_acceleration at velocity support points inside the solid,
                                 (v_s - v_f)/dt + (grad v_f) v_f - a_s, first touching cell wins; and the
                                 Dirichlet variant (use_dirichlet_bc) on non-cell-interior support points
together with Utils::GridInterpolator / CellLocator (source/utilities.cpp:193-341): locate the solid cell,
invert the Q1 map by Newton (MappingQ1::transform_real_to_unit_cell), evaluate the Q1 solid field.

deal.II pieces restated from their documented behaviour ("parity unpinned": the reference has no test that
checks an indicator or an interpolated value): CellAccessor::point_inside = unit-cell test after the inverse
Q1 map; find_active_cell_around_point accepts points within 1e-10 of the unit cell - the same tolerance TOL
is used for both here and in the CUDA kernels.
Plain Python loops: small cases only.
"""
from __future__ import annotations

import numpy as np

from . import fem

TOL = 1e-10


def deformed(X, u, dim):
    return X + u.reshape(-1, dim)


def solid_box(x):
    """[min0, max0, min1, max1, ...] (update_solid_box)"""
    box = np.empty(2 * x.shape[1])
    box[0::2] = x.min(axis=0)
    box[1::2] = x.max(axis=0)
    return box


def q1_shape(dim, xi):
    """N[2^dim], dN[2^dim][dim] of the Q1 element at xi (lexicographic vertices, x fastest)."""
    nv = 1 << dim
    N = np.ones(nv)
    dN = np.ones((nv, dim))
    for v in range(nv):
        for d in range(dim):
            bit = (v >> d) & 1
            f = xi[d] if bit else 1.0 - xi[d]
            df = 1.0 if bit else -1.0
            N[v] *= f
            for e in range(dim):
                dN[v, e] *= df if e == d else f
    return N, dN


def inverse_q1(verts, p, max_it=30):
    """transform_real_to_unit_cell for a Q1 cell: Newton from the cell centre. Returns (xi, converged)."""
    dim = verts.shape[1]
    xi = np.full(dim, 0.5)
    for _ in range(max_it):
        N, dN = q1_shape(dim, xi)
        r = N @ verts - p
        J = verts.T @ dN  # J[i][j] = dx_i / dxi_j
        try:
            dx = np.linalg.solve(J, r)
        except np.linalg.LinAlgError:
            return xi, False
        xi = xi - dx
        if np.linalg.norm(dx) < 1e-13:
            return xi, True
    return xi, False


def point_in_cell(verts, p, tol=TOL):
    lo, hi = verts.min(axis=0), verts.max(axis=0)
    if np.any(p < lo - 1e-12) or np.any(p > hi + 1e-12):
        return False, None
    xi, ok = inverse_q1(verts, p)
    if not ok:
        return False, None
    return bool(np.all(xi >= -tol) and np.all(xi <= 1.0 + tol)), xi


def point_in_solid_2d(point, box, segments):
    """mpi_fsi.cpp:154-215, statement by statement. segments: [(p1, p2)] of the deformed boundary faces."""
    cross_number = 0
    half_cross_number = 0
    for p1, p2 in segments:
        y_diff1 = p1[1] - point[1]
        y_diff2 = p2[1] - point[1]
        x_diff1 = p1[0] - point[0]
        x_diff2 = p2[0] - point[0]
        r1 = p1 - p2
        r2 = np.zeros(2)
        if r1[1] != 0.0:
            r2 = r1 * (point[1] - p2[1]) / r1[1]
        if y_diff1 * y_diff2 < 0:
            if r2[0] + p2[0] > point[0]:
                cross_number += 1
            elif r2[0] + p2[0] == point[0]:
                return True
        elif y_diff1 * y_diff2 == 0:
            if y_diff1 == 0 and y_diff2 == 0:
                if x_diff1 * x_diff2 < 0:
                    return True
                else:
                    continue
            elif r2[0] + p2[0] > point[0]:
                if point[1] != box[2] and point[1] != box[3]:
                    half_cross_number += 1
            elif np.array_equal(point, p1) or np.array_equal(point, p2):
                return True
    cross_number += half_cross_number // 2
    return cross_number % 2 == 1


class SolidGeometry:
    """Deformed solid mesh + what point_in_solid / the interpolator need."""

    def __init__(self, mesh: fem.BoxMesh, displacement):
        self.dim = mesh.dim
        self.cells = mesh.cells
        self.x = deformed(mesh.vertices, displacement, mesh.dim)
        self.box = solid_box(self.x)
        # bounding boxes of the cells: the same reject point_in_cell applies first, for all cells at once (speed only - the cells
        # that pass are visited in index order exactly as before)
        X = self.x[self.cells]
        self.cell_lo, self.cell_hi = X.min(axis=1) - 1e-12, X.max(axis=1) + 1e-12
        self.segments = []
        if self.dim == 2:
            # face f = 2*axis+side of a quad: its two vertices in lexicographic local numbering
            fv = {0: (0, 2), 1: (1, 3), 2: (0, 1), 3: (2, 3)}
            for (cell, face, _id) in mesh.boundary_faces:
                a, b = fv[int(face)]
                self.segments.append((self.x[self.cells[cell, a]], self.x[self.cells[cell, b]]))

    def in_box(self, p):
        return not (np.any(p < self.box[0::2]) or np.any(p > self.box[1::2]))

    def point_in_solid(self, p):
        if not self.in_box(p):
            return False
        if self.dim == 2:
            return point_in_solid_2d(p, self.box, self.segments)
        for c in self._candidates(p):
            if point_in_cell(self.x[self.cells[c]], p)[0]:
                return True
        return False

    def _candidates(self, p):
        return np.nonzero(np.all((p >= self.cell_lo) & (p <= self.cell_hi), axis=1))[0]

    def locate(self, p):
        """lowest-index solid cell containing p (within TOL) and the unit coordinates, or (None, None)"""
        for c in self._candidates(p):
            ok, xi = point_in_cell(self.x[self.cells[c]], p)
            if ok:
                return int(c), np.clip(xi, 0.0, 1.0)  # GeometryInfo::project_to_unit_cell
        return None, None

    def interpolate(self, field, p):
        """GridInterpolator::point_value of a Q1 vector field [n_nodes*dim]; zeros if not found"""
        c, xi = self.locate(p)
        if c is None:
            return None
        N, _ = q1_shape(self.dim, xi)
        vals = field.reshape(-1, self.dim)[self.cells[c]]
        return N @ vals


def update_indicator(fluid_mesh: fem.BoxMesh, solid: SolidGeometry):
    ind = np.zeros(fluid_mesh.n_cells, dtype=np.int32)
    for c in range(fluid_mesh.n_cells):
        inside = 0
        for v in fluid_mesh.cells[c]:
            if not solid.point_in_solid(fluid_mesh.vertices[v]):
                break
            inside += 1
        ind[c] = 1 if inside == fluid_mesh.cells.shape[1] else 0
    return ind


def interpolate_scalar(solid: SolidGeometry, nodal, p):
    c, xi = solid.locate(p)
    if c is None:
        return 0.0  # GridInterpolator::point_value returns 0 when the point is not found
    N, _ = q1_shape(solid.dim, xi)
    return float(N @ nodal[solid.cells[c]])


def find_fluid_bc_stress(fluid, solid: SolidGeometry, indicator, solid_stress, fsi_stress):
    """First part of find_fluid_bc (mpi_fsi.cpp:411-476): fluid has .stress [dim*dim][n_unodes] and the scalar
    FE_Q(pu) support points = velocity nodes; fsi_stress [dim(dim+1)/2][n_unodes] is updated IN PLACE (entries not
    touched keep their values, as in the reference)."""
    dim, d = fluid.dim, fluid.dofs
    touched = np.zeros(d.n_unodes, dtype=bool)
    for c in range(fluid.mesh.n_cells):
        if indicator[c] == 0:
            continue
        for node in d.unodes[c]:
            if touched[node]:
                continue
            touched[node] = True
            p = d.ucoords[node]
            if not solid.point_in_solid(p):
                continue
            k = 0
            for i in range(dim):
                for j in range(i + 1):
                    fsi_stress[k, node] = fluid.stress[i * dim + j, node] - interpolate_scalar(solid, solid_stress[i * dim + j], p)
                    k += 1
    return fsi_stress


def find_fluid_bc(fluid, solid: SolidGeometry, indicator, solid_velocity, solid_acceleration, dt, use_dirichlet_bc=False):
    """fluid: oracle.ins.InsIM (Q2/Q1). Returns fsi_acceleration [n_dofs] and, for the Dirichlet variant,
    (flags [n_dofs], inhomogeneity [n_dofs]) of the inner constraints BEFORE the merge."""
    dim = fluid.dim
    d = fluid.dofs
    nu_loc = d.unodes.shape[1]
    feu, feg = fluid.feu, fem.FEQ(dim, 1)
    unit = feu.unit_points  # support points of the scalar Q2 element, local node order
    Ng_all, dNg_all = feg.eval(unit)
    _, dNu_all = feu.eval(unit)
    fsi_acc = np.zeros(fluid.n)
    con = np.zeros(fluid.n, dtype=np.uint8)
    inhom = np.zeros(fluid.n)
    touched = np.zeros(fluid.n, dtype=bool)
    present = fluid.present
    for c in range(fluid.mesh.n_cells):
        X = fluid.mesh.vertices[fluid.mesh.cells[c]]
        nodes = d.unodes[c]
        if not use_dirichlet_bc:
            if indicator[c] == 0:
                continue
            U = present[: fluid.n_u].reshape(-1, dim)[nodes]  # [nu_loc][dim]
            for a in range(nu_loc):
                for comp in range(dim):
                    g = dim * nodes[a] + comp
                    if touched[g]:
                        continue
                    touched[g] = True
                    p = d.ucoords[nodes[a]]
                    if not solid.point_in_solid(p):
                        continue
                    vs = solid.interpolate(solid_velocity, p)
                    a_s = solid.interpolate(solid_acceleration, p)
                    if vs is None:
                        raise RuntimeError(f"Cannot find point in solid: {p}")
                    J = X.T @ dNg_all[a]
                    G = dNu_all[a] @ np.linalg.inv(J)  # physical gradients of the Q2 shapes at support point a
                    grad_v = U.T @ G  # grad_v[i][k] = d v_i / d x_k
                    v = U[a]
                    fluid_acc = (vs - v) / dt + grad_v @ v
                    fsi_acc[g] = fluid_acc[comp] - a_s[comp]
        else:
            for a in range(nu_loc):
                inside_dim = sum(1 for k in range(dim) if 0 < abs(unit[a][k]) < 1)
                if inside_dim == dim:
                    continue
                for comp in range(dim):
                    g = dim * nodes[a] + comp
                    if touched[g]:
                        continue
                    touched[g] = True
                    p = d.ucoords[nodes[a]]
                    if not solid.point_in_solid(p):
                        continue
                    vs = solid.interpolate(solid_velocity, p)
                    if vs is None:
                        raise RuntimeError(f"Cannot find point in solid: {p}")
                    con[g] = 1
                    inhom[g] = vs[comp] - present[g]
    return fsi_acc, con, inhom


# ----------------------------------------------------------------------------------------------------------
# find_solid_bc (mpi_fsi.cpp:666-867) and the coupled time loop (mpi_fsi.cpp:1172-1214)
# ----------------------------------------------------------------------------------------------------------
def locate_in_fluid(fluid_mesh: fem.BoxMesh, p):
    """lowest-index fluid cell containing p (GridInterpolator on the fluid DoFHandler) and unit coordinates"""
    boxes = getattr(fluid_mesh, "_cell_boxes", None)
    if boxes is None:  # cached on the mesh object (speed only: the same reject point_in_cell applies first)
        X = fluid_mesh.vertices[fluid_mesh.cells]
        boxes = fluid_mesh._cell_boxes = (X.min(axis=1) - 1e-12, X.max(axis=1) + 1e-12)
    for c in np.nonzero(np.all((p >= boxes[0]) & (p <= boxes[1]), axis=1))[0]:
        ok, xi = point_in_cell(fluid_mesh.vertices[fluid_mesh.cells[c]], p)
        if ok:
            return int(c), np.clip(xi, 0.0, 1.0)
    return None, None


def find_solid_bc(fluid, solid_mesh: fem.BoxMesh, displacement, solid_dirichlet_bcs, fluid_stress=None):
    """fluid: oracle fluid solver (dofs, feu, fep, present). Returns fsi_stress_rows [dim][n_sdofs],
    fluid_velocity [n_sdofs], fluid_pressure [n_snodes]."""
    dim = fluid.dim
    d = fluid.dofs
    x = deformed(solid_mesh.vertices, displacement, dim)
    n_sn = solid_mesh.vertices.shape[0]
    rows = np.zeros((dim, n_sn * dim))
    fvel = np.zeros(n_sn * dim)
    fpre = np.zeros(n_sn)
    fixed = (1 << dim) - 1
    done = set()
    for (cell, face, fid) in solid_mesh.boundary_faces:
        if solid_dirichlet_bcs.get(int(fid)) == fixed:
            continue
        for a in fem.face_local_nodes(dim, 1, int(face)):
            node = int(solid_mesh.cells[cell, a])
            if node in done:
                continue
            done.add(node)
            c, xi = locate_in_fluid(fluid.mesh, x[node])
            if c is None:
                continue
            Nu = fluid.feu.eval(xi[None, :])[0][0]
            Np = fluid.fep.eval(xi[None, :])[0][0]
            U = fluid.present[: fluid.n_u].reshape(-1, dim)[d.unodes[c]]
            v = Nu @ U
            pr = float(Np @ fluid.present[fluid.n_u + d.pnodes[c]])
            visc = np.zeros((dim, dim))
            if fluid_stress is not None:
                for i in range(dim):
                    for j in range(i, dim):
                        visc[i, j] = visc[j, i] = float(Nu @ fluid_stress[i * dim + j, d.unodes[c]])
            sigma = -pr * np.eye(dim) + visc
            for d1 in range(dim):
                rows[d1, dim * node: dim * node + dim] = sigma[d1]
            fvel[dim * node: dim * node + dim] = v
            fpre[node] = pr
    return rows, fvel, fpre


class FSI:
    """MPI::FSI<dim>::run loop with an oracle fluid (oracle.scns.SCnsIM) and solid (oracle.solid.HyperElasticity)."""

    def __init__(self, fluid, solid, use_dirichlet_bc=False):
        self.fluid, self.solid, self.use_dirichlet_bc = fluid, solid, use_dirichlet_bc
        self.base_con, self.base_val = fluid.con.copy(), fluid.nonzero_val.copy()
        self.penetration_criterion, self.penetration_direction = None, None
        self.contact_iterations = 0

    def set_penetration_criterion(self, criterion, direction):
        """FSI::set_penetration_criterion (mpi_fsi.cpp:1229-1237): criterion(point) > 1e-5 means the point penetrates"""
        self.penetration_criterion = criterion
        self.penetration_direction = np.asarray(direction, dtype=float)

    def apply_contact_model(self, first_step):
        """FSI::apply_contact_model (mpi_fsi.cpp:869-970): repeat the solid step, each time adding a stress
        multiplier * penetration along the penetration direction to fsi_stress_rows at the penetrating boundary
        vertices (once per (cell, boundary face, face vertex) visit), until no boundary vertex penetrates"""
        s = self.solid
        cached = [v.copy() for v in (s.cur_a, s.cur_v, s.cur_u, s.prev_a, s.prev_v, s.prev_u)]
        still = True
        while still:
            s.run_one_step(first_step)
            self.contact_iterations += 1
            still = self.contact_scan()
            if still:
                s.cur_a, s.cur_v, s.cur_u, s.prev_a, s.prev_v, s.prev_u = [v.copy() for v in cached]
                s.time -= s.dt
                s.timestep -= 1

    def contact_scan(self):
        """the penetration scan of apply_contact_model (mpi_fsi.cpp:897-956) on the solid's current displacement; adds the
        extra stresses to solid.fsi_stress_rows and returns still_penetrate"""
        s = self.solid
        dim = s.dim
        mult = s.prm.contact_force_multiplier
        dirn = self.penetration_direction
        still = False
        x = deformed(s.mesh.vertices, s.cur_u, dim)
        for (cell, face, fid) in s.mesh.boundary_faces:
            axis, side = int(face) // 2, int(face) % 2
            cn = s.mesh.cells[cell]
            J = np.einsum("vi,vj->ij", x[cn], s.face_dG[face][0])
            nds = np.linalg.det(J) * np.linalg.inv(J)[axis, :] * (1.0 if side else -1.0)
            normal = nds / np.linalg.norm(nds)  # fe_face_values.normal_vector(0) on the moved mesh
            for a in fem.face_local_nodes(dim, 1, int(face)):
                node = int(cn[a])
                pen = self.penetration_criterion(x[node])
                if not pen > 1e-5:
                    continue
                still = True
                traction = mult * pen / np.linalg.norm(dirn) * dirn
                for d1 in range(dim):
                    extra = traction[d1] / normal[d1] if normal[d1] > 1e-5 else 0.0
                    s.fsi_stress_rows[d1, dim * node + dim - 1] += extra
        return still

    def refine_mesh(self, tria, min_grid_level, max_grid_level):
        """FSI::refine_mesh (mpi_fsi.cpp:1024-1117). `tria` is the host-side triangulation object that owns the fluid mesh (the
        product's mesh class: flagged coarsening / refinement with 2:1 balancing is host infrastructure, like the mesh generators);
        the flags (:1030-1080) and the solution transfer (:1082-1110, parallel::distributed::SolutionTransfer = the old FE field
        evaluated at the new support points) are restated here. The fluid solver is rebuilt on the new mesh as setup_dofs /
        make_constraints / initialize_system do, with present_solution interpolated."""
        from . import grid, scns

        f, s = self.fluid, self.solid
        dim = f.dim
        x = deformed(s.mesh.vertices, s.cur_u, dim)
        first_face = {}
        for (cell, face, _fid) in s.mesh.boundary_faces:
            first_face[int(cell)] = min(first_face.get(int(cell), 99), int(face))
        pts = np.array([x[s.mesh.cells[c, fem.face_local_nodes(dim, 1, face)]].mean(axis=0) for c, face in sorted(first_face.items())])
        X = f.mesh.vertices[f.mesh.cells]  # [nc][2^d][dim]
        centre = X.mean(axis=1)
        diam = np.sqrt(((X[:, :, None, :] - X[:, None, :, :]) ** 2).sum(-1)).max(axis=(1, 2))
        dist = np.sqrt(((centre[:, None, :] - pts[None, :, :]) ** 2).sum(-1)).min(axis=1)
        refine = dist < diam
        coarsen = ~refine
        level = tria.levels()
        if level.max() + 1 > max_grid_level:
            refine &= level < max_grid_level
        coarsen &= level != min_grid_level
        old_mesh, old = f.mesh, f.present.copy()
        old_d = f.dofs
        tria.execute_coarsening_and_refinement(refine, coarsen)
        v, c, b = tria.get_mesh()
        new = type(f)((grid.QuadMesh if dim == 2 else grid.HexMesh)(v, c, b), f.prm, hard_coded=f.hard_coded)
        # solution transfer: the old Q1 field at the new support points
        feu, fep = f.feu, f.fep
        for coords, nodes_old, n_comp, off_old, off_new in ((new.dofs.ucoords, old_d.unodes, dim, 0, 0), (new.dofs.pcoords, old_d.pnodes, 1, f.n_u, new.n_u)):
            fe = feu if n_comp == dim else fep
            for n, p in enumerate(coords):
                cell, xi = locate_in_fluid(old_mesh, p)
                N = fe.eval(xi[None, :])[0][0]
                for comp in range(n_comp):
                    new.present[off_new + n_comp * n + comp] = N @ old[off_old + n_comp * nodes_old[cell] + comp]
        new.time, new.timestep, new.history, new.bc_time = f.time, f.timestep, f.history, f.bc_time
        if getattr(f, "turbulence_model", None) is not None:
            # pre_refine_mesh / post_refine_mesh of the turbulence model (mpi_fsi.cpp:1093-1096, 1113-1116): nu~ travels like the
            # fluid's solution; the model is re-initialised on the new mesh first (mpi_supg_solver.cpp:290-293)
            tm_old, tm = f.turbulence_model, new.attach_turbulence_model("Spalart-Allmaras")
            for n, p in enumerate(new.dofs.ucoords):
                cell, xi = locate_in_fluid(old_mesh, p)
                tm.present[n] = feu.eval(xi[None, :])[0][0] @ tm_old.present[old_d.unodes[cell]]
            tm.history = tm_old.history
        self.fluid = new
        self.base_con, self.base_val = new.con.copy(), new.nonzero_val.copy()
        return new

    def run(self):
        """the time loop of FSI::run (mpi_fsi.cpp:1172-1226) without refinement / checkpoints"""
        p = self.fluid.prm
        t, first = 0.0, True
        while p.end_time - t > 1e-12:
            self.run_one_step(first)
            first = False
            t += p.time_step

    def run_one_step(self, first_step):
        f, s = self.fluid, self.solid
        dim = f.dim
        s.fsi_stress_rows, s.fluid_velocity, s.fluid_pressure = find_solid_bc(
            f, s.mesh, s.cur_u, s.prm.solid_dirichlet_bcs, getattr(f, "stress", None))
        if self.penetration_criterion is not None:
            self.apply_contact_model(first_step)
        else:
            s.run_one_step(first_step)
        geo = SolidGeometry(s.mesh, s.cur_u)
        f.indicator[:] = update_indicator(f.mesh, geo)
        # make_constraints(); after the first step the nonzero constraints become the zero ones
        f.con = self.base_con.copy()
        f.nonzero_val = self.base_val.copy() if first_step else np.zeros_like(self.base_val)
        if hasattr(f, "fsi_stress"):
            find_fluid_bc_stress(f, geo, f.indicator, s.stress, f.fsi_stress)
        acc, icon, iinh = find_fluid_bc(f, geo, f.indicator, s.cur_v, s.cur_a, f.dt, self.use_dirichlet_bc)
        f.fsi_acceleration[:] = acc
        if self.use_dirichlet_bc:
            new = (icon != 0) & (f.con == 0) & ~f.dofs.is_hanging  # merge(..., left_object_wins): existing lines stay
            f.con[new] = 1
            f.nonzero_val[new] = iinh[new]
        tm = getattr(f, "turbulence_model", None)
        if tm is not None:  # mpi_fsi.cpp:1199-1210 (update_boundary_condition comes before find_fluid_bc there; they are independent)
            tm.make_constraints()
            tm.update_boundary_condition(first_step)
            tm.run_one_step(True)
        f.run_one_step(True)
