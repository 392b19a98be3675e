"""ORACLE (test infrastructure, NOT product code).

CPU restatement of Fluid::MPI::SCnsIM (reference source/mpi_scnsim.cpp: assemble, C++ loop in
oracle/csrc/oracle_scns.cpp) on top of Fluid::MPI::SUPGFluidSolver (source/mpi_supg_solver.cpp: solve :297-328,
run_one_step :331-425, run :428-486) and the FluidSolver base pieces it uses: update_stress
(source/mpi_fluid_solver.cpp:716-811), apply_initial_condition (:368-414), set_body_force /
set_sigma_pml_field (include/mpi_fluid_solver.h:120-143).

The reference solves the Newton systems with FGMRES to 1e-6 |rhs| preconditioned by ILU(0)-based Schur
factors (Hypre Euclid, rank-count dependent, not vendored); the oracle uses a sparse direct solve, which
agrees with any converged FGMRES solve to that tolerance ("parity unpinned" for the iterates, pinned on the
reference goldens tests/fluid_body_force_mpi and tests/fluid_initial_condition_mpi at the field level).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem
from .ins import _p, lib


def h_shape_functions(dim, pu, pp, dofs_per_cell):
    """The first dofs_per_cell / dofs_per_vertex shape functions of FESystem(FE_Q(pu)^dim, FE_Q(pp)) in deal.II's
    cell-local numbering (vertex dofs first; per vertex: dim velocity components, then the pressure):
    (type, local node) with type 0 = velocity shape of that node, 1 = pressure shape (mpi_scnsim.cpp:251-257)."""
    n_h = dofs_per_cell // (dim + 1)
    feu, fep = fem.FEQ(dim, pu), fem.FEQ(dim, pp)
    out = []
    for v in range(1 << dim):
        bits = [(v >> d) & 1 for d in range(dim)]
        un = int(np.nonzero(np.all(feu.lattice == np.array(bits) * pu, axis=1))[0][0])
        pn = int(np.nonzero(np.all(fep.lattice == np.array(bits) * pp, axis=1))[0][0])
        out += [(0, un)] * dim + [(1, pn)]
    out = out[:n_h]
    return np.array([t for t, _ in out], dtype=np.int32), np.array([n for _, n in out], dtype=np.int32)


class SCnsIM:
    def __init__(self, mesh: fem.BoxMesh, params, body_force=None, sigma_pml_field=None, initial_condition=None, hard_coded=None):
        self.mesh, self.prm = mesh, params
        dim = mesh.dim
        self.dim = dim
        pu, pp = params.fluid_velocity_degree, params.fluid_pressure_degree
        self.dofs = fem.FluidDofs(mesh, pu, pp)
        d = self.dofs
        self.n_u, self.n_p, self.n = d.n_u, d.n_p, d.n_dofs
        feu, fep, feg = fem.FEQ(dim, pu), fem.FEQ(dim, pp), fem.FEQ(dim, 1)
        self.feu, self.fep = feu, fep
        qp, qw = fem.qgauss(dim, pu + 1)
        self.qp, self.qw, self.nq = qp, np.ascontiguousarray(qw), qw.size
        self.Nu, self.dNu = [np.ascontiguousarray(a) for a in feu.eval(qp)]
        self.Np, self.dNp = [np.ascontiguousarray(a) for a in fep.eval(qp)]
        self.Ngeo, self.dNgeo = [np.ascontiguousarray(a) for a in feg.eval(qp)]
        fq, fw = fem.qgauss(dim - 1, pu + 1)
        self.nqf, self.qwf = fw.size, np.ascontiguousarray(fw)
        Nuf, dGf = [], []
        for face in range(2 * dim):
            pts = np.insert(fq, face // 2, float(face % 2), axis=1)
            Nuf.append(feu.eval(pts)[0])
            dGf.append(feg.eval(pts)[1])
        self.Nu_face = np.ascontiguousarray(np.stack(Nuf))
        self.dNgeo_face = np.ascontiguousarray(np.stack(dGf))
        self.hard_coded, self.bc_time = hard_coded, 0.0
        self.con, self.nonzero_val = fem.make_dirichlet_constraints(d, params.fluid_dirichlet_bcs, hard_coded)
        self.rowptr, self.col = fem.full_pattern(d.cell_dofs, self.n, d.hanging_dofs)
        self.h_type, self.h_node = h_shape_functions(dim, pu, pp, d.dofs_per_cell)
        self.present = np.zeros(self.n)
        self.evaluation_point = np.zeros(self.n)
        self.fsi_acceleration = np.zeros(self.n)
        self.indicator = np.zeros(mesh.n_cells, dtype=np.int32)
        self.stress = np.zeros((dim * dim, d.n_unodes))
        self.fsi_stress = np.zeros((dim * (dim + 1) // 2, d.n_unodes))
        self.time, self.timestep, self.dt = 0.0, 0, params.time_step
        self.history = []
        self.turbulence_model = None
        self.vertices = np.ascontiguousarray(mesh.vertices)
        self.cells = np.ascontiguousarray(mesh.cells)
        self.cell_dofs = np.ascontiguousarray(d.cell_dofs)
        self.cell_unodes = np.ascontiguousarray(d.unodes)
        self.bfaces = np.ascontiguousarray(mesh.boundary_faces)
        # quadrature point coordinates (Q1 map) for the user fields
        X = mesh.vertices[mesh.cells]
        self.xq = np.einsum("qv,cvd->cqd", self.Ngeo, X)
        self.sigma_pml = None
        self.body_force = None
        if sigma_pml_field is not None:
            self.sigma_pml = np.ascontiguousarray([[sigma_pml_field(x, 0) for x in cq] for cq in self.xq], dtype=np.float64)
        if body_force is not None:
            self.body_force = np.ascontiguousarray([[[body_force(x, c) for c in range(dim)] for x in cq] for cq in self.xq],
                                                   dtype=np.float64)
        if initial_condition is not None:  # apply_initial_condition (:368-414)
            pts = d.support_points()
            for g in range(self.n_u):
                self.present[g] = initial_condition(pts[g], g % dim)
            for g in range(self.n_u, self.n):
                self.present[g] = initial_condition(pts[g], dim)
        # projection from quadrature points to the dofs of scalar FE_Q(pu) on the reference cell
        Mref = np.einsum("qi,qj,q->ij", self.Nu, self.Nu, self.qw)
        self.qpt_to_dof = np.linalg.solve(Mref, (self.Nu * self.qw[:, None]).T)

    def attach_turbulence_model(self, model_name: str):
        """FluidSolver::attach_turbulence_model (mpi_fluid_solver.cpp:53-63); the model's make_constraints / initialize_system
        (:276-279, mpi_supg_solver.cpp:290-293) run here because this class sets itself up in the constructor"""
        if model_name != "Spalart-Allmaras":
            raise NotImplementedError(model_name)  # TurbulenceModelFactory::create: ExcNotImplemented
        from .spalart_allmaras import SpalartAllmaras

        self.turbulence_model = SpalartAllmaras(self)
        return self.turbulence_model

    def _closed_constraints(self, use_nonzero_constraints: bool):
        """(flags, inhomogeneities or None) of nonzero_constraints / zero_constraints after close(); hanging-node lines
        (locally refined meshes) are handed to the C cell loops through oracle_set_constraint_lines"""
        if not self.dofs.hanging_dofs:
            return self.con, (self.nonzero_val if use_nonzero_constraints else None)
        val = self.nonzero_val if use_nonzero_constraints else np.zeros(self.n)
        con, inhom, ptr, master, weight = fem.resolve_constraints(self.dofs, self.con, val)
        lib().oracle_set_constraint_lines(C.c_int64(self.n), _p(ptr, C.c_int64), _p(master, C.c_int), _p(weight))
        return con, (np.ascontiguousarray(inhom) if use_nonzero_constraints else None)

    def _release_constraints(self):
        if self.dofs.hanging_dofs:
            lib().oracle_set_constraint_lines(C.c_int64(0), None, None, None)

    def assemble(self, use_nonzero_constraints: bool):
        p = self.prm
        A = np.zeros(self.col.size)
        rhs = np.zeros(self.n)
        con, inhom = self._closed_constraints(use_nonzero_constraints)
        nids = np.asarray(sorted(p.fluid_neumann_bcs), dtype=np.int32)
        nvals = np.asarray([p.fluid_neumann_bcs[i] for i in nids], dtype=np.float64)
        grav = np.asarray(p.gravity, dtype=np.float64)
        stress = np.ascontiguousarray(self.stress)
        fsis = np.ascontiguousarray(self.fsi_stress)
        eddy = np.ascontiguousarray(self.turbulence_model.eddy_viscosity) if self.turbulence_model is not None else None
        lib().oracle_scns_set_eddy_viscosity(_p(eddy))
        rc = lib().oracle_scns_assemble(
            C.c_int(self.dim), C.c_int(self.feu.n), C.c_int(self.fep.n), C.c_int(self.mesh.n_cells), _p(self.vertices),
            _p(self.cells, C.c_int), _p(self.cell_dofs, C.c_int), _p(self.cell_unodes, C.c_int), C.c_int(self.nq), _p(self.qw),
            _p(self.Nu), _p(self.dNu), _p(self.Np), _p(self.dNp), _p(self.dNgeo), C.c_int(self.nqf), _p(self.qwf),
            _p(self.Nu_face), _p(self.dNgeo_face), _p(self.evaluation_point), _p(self.present), _p(self.fsi_acceleration),
            _p(self.indicator, C.c_int), _p(stress), _p(fsis), C.c_int(self.dofs.n_unodes), _p(self.sigma_pml), _p(self.body_force),
            C.c_int(self.h_type.size), _p(self.h_type, C.c_int), _p(self.h_node, C.c_int), C.c_double(p.viscosity),
            C.c_double(p.fluid_rho), C.c_double(p.solid_rho), C.c_double(self.dt), _p(grav), C.c_int(self.bfaces.shape[0]),
            _p(self.bfaces, C.c_int), C.c_int(nids.size), _p(nids, C.c_int), _p(nvals), _p(con, C.c_ubyte), _p(inhom),
            _p(self.rowptr, C.c_int64), _p(self.col, C.c_int), _p(A), _p(rhs))
        self._release_constraints()
        lib().oracle_scns_set_eddy_viscosity(None)
        assert rc == 0
        self.system_matrix = sp.csr_matrix((A, self.col, self.rowptr), shape=(self.n, self.n))
        self.system_rhs = rhs
        return self.system_matrix, rhs

    def solve(self, use_nonzero_constraints: bool):
        x = spla.spsolve(self.system_matrix.tocsc(), self.system_rhs)
        x[self.con != 0] = self.nonzero_val[self.con != 0] if use_nonzero_constraints else 0.0
        fem.distribute(self.dofs, x)  # constraints.distribute(newton_update), mpi_supg_solver.cpp:323-325
        self.newton_update = x
        return 0, 0.0

    def update_stress(self):
        """FluidSolver::update_stress (:716-811): 2 mu sym grad v at q -> qpt_to_dof -> nodal average"""
        dim, d = self.dim, self.dofs
        X = self.mesh.vertices[self.mesh.cells]
        Jm = np.einsum("cvi,qvj->cqij", X, self.dNgeo)
        G = np.einsum("qaj,cqjk->cqak", self.dNu, np.linalg.inv(Jm))
        U = self.present[: self.n_u].reshape(-1, dim)[d.unodes]  # [c][a][comp]
        grad = np.einsum("cai,cqak->cqik", U, G)
        tau = self.prm.viscosity * (grad + np.swapaxes(grad, -1, -2))  # 2 mu sym grad
        cell_stress = np.einsum("aq,cqij->cija", self.qpt_to_dof, tau)
        stress = np.zeros((dim * dim, d.n_unodes))
        count = np.zeros(d.n_unodes)
        np.add.at(count, d.unodes.ravel(), 1.0)
        for i in range(dim):
            for j in range(dim):
                np.add.at(stress[i * dim + j], d.unodes.ravel(), cell_stress[:, i, j, :].ravel())
        self.stress = stress / count[None, :]
        return self.stress

    def run_one_step(self, apply_nonzero_constraints: bool):
        p = self.prm
        self.timestep += 1
        self.time += self.dt
        current_residual = initial_residual = relative_residual = 1.0
        outer = 0
        self.evaluation_point = self.present.copy()
        while relative_residual > p.fluid_tolerance and current_residual > 1e-14:
            if outer >= p.fluid_max_iterations:
                raise RuntimeError("Too many Newton iterations!")
            nz = apply_nonzero_constraints and outer == 0
            self.assemble(nz)
            self.solve(nz)
            current_residual = np.linalg.norm(self.system_rhs)
            self.evaluation_point = self.evaluation_point + self.newton_update
            if outer == 0:
                initial_residual = current_residual
            relative_residual = current_residual / initial_residual
            self.history.append((self.timestep, outer, current_residual, relative_residual))
            outer += 1
        self.present = self.evaluation_point.copy()
        self.update_stress()

    def make_constraints(self):
        self.con, self.nonzero_val = fem.make_dirichlet_constraints(self.dofs, self.prm.fluid_dirichlet_bcs, self.hard_coded,
                                                                    self.bc_time)
        if self.turbulence_model is not None:  # mpi_fluid_solver.cpp:276-279
            self.turbulence_model.make_constraints()

    def run(self, max_steps=None):
        """SUPGFluidSolver::run (mpi_supg_solver.cpp:427-486): with hard-coded boundary values the functions' clock is
        advanced by dt before every make_constraints() (:438-444, :470-478) and every step applies the nonzero
        constraints - the functions return the *increment* of the boundary value over the step."""
        if self.hard_coded:
            self.bc_time += self.dt
            self.make_constraints()
        if self.turbulence_model is not None:  # :456-461
            self.turbulence_model.run_one_step(True)
        self.run_one_step(True)
        k = 1
        while self.prm.end_time - self.time > 1e-12 and (max_steps is None or k < max_steps):
            if self.turbulence_model is not None:  # :464-469
                self.turbulence_model.run_one_step(False)
            if self.hard_coded:
                self.bc_time += self.dt
                self.make_constraints()
                self.run_one_step(True)
            else:
                self.run_one_step(False)
            k += 1

    def velocity(self):
        return self.present[: self.n_u]

    def pressure(self):
        return self.present[self.n_u:]


class SUPGInsIM(SCnsIM):
    """Fluid::MPI::SUPGInsIM<dim> (reference source/mpi_insim_supg.cpp; cell loop in oracle/csrc/oracle_insim_supg.cpp):
    incompressible Navier-Stokes with SUPG / PSPG / LSIC stabilisation on SUPGFluidSolver - everything but assemble() is
    shared with SCnsIM (Newton loop, solve, update_stress, time loop)."""

    def assemble(self, use_nonzero_constraints: bool):
        p = self.prm
        A = np.zeros(self.col.size)
        rhs = np.zeros(self.n)
        con, inhom = self._closed_constraints(use_nonzero_constraints)
        nids = np.asarray(sorted(p.fluid_neumann_bcs), dtype=np.int32)
        nvals = np.asarray([p.fluid_neumann_bcs[i] for i in nids], dtype=np.float64)
        grav = np.asarray(p.gravity, dtype=np.float64)
        rc = lib().oracle_insim_supg_assemble(
            C.c_int(self.dim), C.c_int(self.feu.n), C.c_int(self.fep.n), C.c_int(self.mesh.n_cells), _p(self.vertices),
            _p(self.cells, C.c_int), _p(self.cell_dofs, C.c_int), C.c_int(self.nq), _p(self.qw), _p(self.Nu), _p(self.dNu),
            _p(self.Np), _p(self.dNp), _p(self.dNgeo), C.c_int(self.nqf), _p(self.qwf), _p(self.Nu_face), _p(self.dNgeo_face),
            _p(self.evaluation_point), _p(self.present), _p(self.body_force), C.c_int(self.h_type.size), _p(self.h_type, C.c_int),
            _p(self.h_node, C.c_int), C.c_double(p.viscosity), C.c_double(p.fluid_rho), C.c_double(self.dt), _p(grav),
            C.c_int(self.bfaces.shape[0]), _p(self.bfaces, C.c_int), C.c_int(nids.size), _p(nids, C.c_int), _p(nvals),
            _p(con, C.c_ubyte), _p(inhom), _p(self.rowptr, C.c_int64), _p(self.col, C.c_int), _p(A), _p(rhs))
        self._release_constraints()
        assert rc == 0
        self.system_matrix = sp.csr_matrix((A, self.col, self.rowptr), shape=(self.n, self.n))
        self.system_rhs = rhs
        return self.system_matrix, rhs
