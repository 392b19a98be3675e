"""ORACLE (test infrastructure, NOT product code).

CPU restatement of Solid::MPI::HyperElasticity (reference source/mpi_hyper_elasticity.cpp) with the
NeoHookean material (source/hyper_elastic_material.cpp:8-39, include/hyper_elastic_material.h:50-70,
include/neo_hookean.h:26-34) and the solver base Solid::MPI::SolidSolver
(source/mpi_solid_solver.cpp:43-161, 316-328) on box meshes, FE_Q(1)^dim, QGauss(2).

Third-party pieces restated from their documented behaviour (deal.II is not vendored):
  * Physics::Elasticity::Kinematics::F / F_iso / b : F = I + Grad u, F_iso = J^(-1/dim) F, b = F F^T
  * StandardTensors: I, IxI, S (symmetric 4th-order identity), dev_P = S - IxI/dim
  * PETSc CG + PCBJACOBI(ILU0) to 1e-8 |b| (mpi_solid_solver.cpp:151-157) -> sparse direct solve here
    (converged results agree to the solver tolerance; the reference's iterates depend on the rank count).
Pinned on the reference golden tests/solid_beam_bending_mpi_NeoHookean (2-D: u_min -0.0616287,
u_max 0.00867069, rel 1e-3) in tests/test_oracle_goldens.py.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem


def standard_tensors(dim):
    I = np.eye(dim)
    IxI = np.einsum("ij,kl->ijkl", I, I)
    S = 0.5 * (np.einsum("ik,jl->ijkl", I, I) + np.einsum("il,jk->ijkl", I, I))
    dev_P = S - IxI / dim
    return I, IxI, S, dev_P


def neo_hookean_update(grad_u, c1, kappa):
    """PointHistory::update for NeoHookean (mpi_hyper_elasticity.cpp:37-65): grad_u [...,dim,dim] ->
    F_inv [...,dim,dim], tau [...,dim,dim], Jc [...,dim,dim,dim,dim], det_F [...]."""
    dim = grad_u.shape[-1]
    I, IxI, S, dev_P = standard_tensors(dim)
    F = I + grad_u
    J = np.linalg.det(F)
    F_inv = np.linalg.inv(F)
    Fb = F * (J ** (-1.0 / dim))[..., None, None]
    b_bar = np.einsum("...ik,...jk->...ij", Fb, Fb)
    tau_bar = 2.0 * c1 * b_bar
    tau_iso = np.einsum("ijkl,...kl->...ij", dev_P, tau_bar)
    p = kappa * (J - 1.0)
    tau_vol = (J * p)[..., None, None] * I
    tau = tau_iso + tau_vol
    p_tilde = p + J * kappa
    Jc_vol = J[..., None, None, None, None] * (p_tilde[..., None, None, None, None] * IxI - 2.0 * p[..., None, None, None, None] * S)
    tr = np.einsum("...ii->...", tau_bar)
    Jc_iso = (2.0 / dim) * tr[..., None, None, None, None] * dev_P - (2.0 / dim) * (
        np.einsum("...ij,kl->...ijkl", tau_iso, I) + np.einsum("ij,...kl->...ijkl", I, tau_iso))
    return F_inv, tau, Jc_iso + Jc_vol, J


def kirchhoff_update(grad_u, young, poisson):
    """PointHistory::update for solid_type = Kirchhoff (mpi_hyper_elasticity.cpp:50-58, include/kirchhoff_elastic_material.h:
    36-76): E = (F^T F - I) / 2, S = lambda tr(E) I + 2 mu E, tau = F S F^T; Jc = lambda IxI + 2 mu S4 (constant)."""
    dim = grad_u.shape[-1]
    I, IxI, S4, _ = standard_tensors(dim)
    lam = young * poisson / ((1 + poisson) * (1 - 2 * poisson))
    mu = young / (2 * (1 + poisson))
    F = I + grad_u
    J = np.linalg.det(F)
    F_inv = np.linalg.inv(F)
    E = 0.5 * (np.einsum("...ki,...kj->...ij", F, F) - I)
    pk2 = lam * np.einsum("...ii->...", E)[..., None, None] * I + 2 * mu * E
    tau = np.einsum("...ik,...kl,...jl->...ij", F, pk2, F)
    Jc = np.broadcast_to(lam * IxI + 2 * mu * S4, grad_u.shape[:-2] + (dim,) * 4).copy()
    return F_inv, tau, Jc, J


class SolidDofs:
    def __init__(self, mesh: fem.BoxMesh, degree: int):
        self.mesh, self.dim = mesh, mesh.dim
        self.nodes, self.n_nodes, self.coords = mesh.node_table(degree)
        dim = mesh.dim
        self.n_dofs = dim * self.n_nodes
        npc = self.nodes.shape[1]
        self.cell_dofs = (self.nodes[:, :, None] * dim + np.arange(dim)[None, None, :]).reshape(mesh.n_cells, npc * dim).astype(np.int32)


class SolidBase:
    """what Solid::MPI::SolidSolver / SharedSolidSolver set up for every material (mpi_solid_solver.cpp:43-161,
    mpi_shared_solid_solver.cpp:45-234): FE_Q(1)^dim, QGauss(2), homogeneous Dirichlet constraints, Newmark vectors"""

    def __init__(self, mesh: fem.BoxMesh, params, verbose=False):
        self.mesh, self.prm, self.verbose = mesh, params, verbose
        dim = mesh.dim
        self.dim = dim
        deg = params.solid_degree
        assert deg == 1
        self.dofs = SolidDofs(mesh, deg)
        self.n = self.dofs.n_dofs
        fe, feg = fem.FEQ(dim, deg), fem.FEQ(dim, 1)
        qp, qw = fem.qgauss(dim, deg + 1)
        self.qw = qw
        self.N, self.dN = fe.eval(qp)  # [nq][n], [nq][n][dim]
        self.nq, self.npc = qw.size, fe.n
        # geometry per cell / q (Q1 map): J[c,q,i,j] = sum_v X[c,v,i] dN[q,v,j]
        X = mesh.vertices[mesh.cells]  # [nc][nv][dim]
        Jm = np.einsum("cvi,qvj->cqij", X, feg.eval(qp)[1])
        self.JxW = np.linalg.det(Jm) * qw[None, :]
        Jinv = np.linalg.inv(Jm)
        # physical gradients of the scalar shape functions G[c,q,a,k] = dN[q,a,j] Jinv[c,q,j,k]
        self.G = np.einsum("qaj,cqjk->cqak", self.dN, Jinv)
        # face quadrature tables for Neumann faces
        fq, fw = fem.qgauss(dim - 1, deg + 1)
        self.fw = fw
        self.face_N, self.face_dG = [], []
        for face in range(2 * dim):
            axis, side = face // 2, face % 2
            pts = np.insert(fq, axis, float(side), axis=1)
            self.face_N.append(fe.eval(pts)[0])
            self.face_dG.append(feg.eval(pts)[1])
        # homogeneous Dirichlet constraints (mpi_solid_solver.cpp:67-94)
        self.con = np.zeros(self.n, dtype=np.uint8)
        for bid, flag in sorted(params.solid_dirichlet_bcs.items()):
            comps = fem.component_mask(flag, dim)
            for (cell, face_no, fid) in mesh.boundary_faces:
                if fid != bid:
                    continue
                for a in fem.face_local_nodes(dim, deg, face_no):
                    for c in comps:
                        self.con[dim * self.dofs.nodes[cell, a] + c] = 1
        self.rowptr, self.col = fem.full_pattern(self.dofs.cell_dofs, self.n)
        self.rho = params.solid_rho
        self.dt = params.time_step
        self.time, self.timestep = 0.0, 0
        z = lambda: np.zeros(self.n)
        self.cur_u, self.cur_v, self.cur_a = z(), z(), z()
        self.prev_u, self.prev_v, self.prev_a = z(), z(), z()
        self.history = []
        self.fsi_stress_rows = np.zeros((dim, self.n))
        self.fluid_velocity = np.zeros(self.n)
        self.fluid_pressure = np.zeros(self.dofs.n_nodes)
        # initial velocity (mpi_shared_solid_solver.cpp:152-196, mpi_solid_solver.cpp:116-137), constraints distributed
        iv = np.asarray(list(params.initial_velocity)[:dim], dtype=float)
        if np.any(iv != 0):
            self.prev_v = np.tile(iv, self.dofs.n_nodes)
            self.prev_v[self.con != 0] = 0.0
            self.cur_v = self.prev_v.copy()

    # -- Neumann faces: traction / pressure w.r.t. the reference configuration, or (FSI) sigma_f n on the deformed face --
    def face_rhs(self, rhs):
        dim, npc, p = self.dim, self.npc, self.prm
        if p.simulation_type == "FSI":
            # FSI traction sigma_f n on the deformed face (mpi_shared_hyper_elasticity.cpp:495-554,
            # mpi_shared_linear_elasticity.cpp:193-271)
            rows = self.fsi_stress_rows.reshape(dim, -1, dim)  # [d1][node][d2]
            xdef = self.mesh.vertices + self.cur_u.reshape(-1, dim)
            for (cell, face, fid) in self.mesh.boundary_faces:
                axis, side = face // 2, face % 2
                cn = self.mesh.cells[cell]
                X = xdef[cn]
                for q in range(self.fw.size):
                    J = np.einsum("vi,vj->ij", X, self.face_dG[face][q])
                    nds = np.linalg.det(J) * np.linalg.inv(J)[axis, :] * (1.0 if side else -1.0)
                    dS = np.linalg.norm(nds)
                    sigma = np.einsum("b,ibj->ij", self.face_N[face][q], rows[:, self.dofs.nodes[cell], :])
                    traction = sigma @ (nds / dS)
                    for a in range(npc):
                        for c in range(dim):
                            rhs[cell, a * dim + c] += self.face_N[face][q, a] * traction[c] * dS * self.fw[q]
        if p.simulation_type != "FSI" and p.solid_neumann_bcs:
            for (cell, face, fid) in self.mesh.boundary_faces:
                if (self.skip_dirichlet_faces and fid in p.solid_dirichlet_bcs) or fid not in p.solid_neumann_bcs:
                    continue
                axis, side = face // 2, face % 2
                X = self.mesh.vertices[self.mesh.cells[cell]]
                val = p.solid_neumann_bcs[fid]
                for q in range(self.fw.size):
                    J = np.einsum("vi,vj->ij", X, self.face_dG[face][q])
                    nds = np.linalg.det(J) * np.linalg.inv(J)[axis, :] * (1.0 if side else -1.0)  # n dS / w
                    dS = np.linalg.norm(nds)
                    if p.solid_neumann_bc_type == "Traction":
                        traction = np.asarray(val[:dim], dtype=float)
                    else:
                        traction = (nds / dS) * val[0]
                    for a in range(npc):
                        for c in range(dim):
                            rhs[cell, a * dim + c] += self.face_N[face][q, a] * traction[c] * dS * self.fw[q]

    skip_dirichlet_faces = True  # mpi_hyper_elasticity.cpp:452-456 skips faces with a Dirichlet id; the linear solvers do not

    def scatter(self, K, f):
        """constraints.distribute_local_to_global with homogeneous constraints: constrained rows keep |K_ii| on the
        diagonal, constrained columns are dropped; K [nc][n][n] or None, f [nc][n] or None"""
        cd = self.dofs.cell_dofs
        n = cd.shape[1]
        con = self.con
        A = rhs = None
        if K is not None:
            rows = np.repeat(cd, n, axis=1).ravel()
            cols = np.tile(cd, (1, n)).ravel()
            vals = K.reshape(-1).copy()
            rc, cc = con[rows] != 0, con[cols] != 0
            diag = rows == cols
            loc_diag = np.tile(np.eye(n, dtype=bool).ravel(), cd.shape[0])
            keep = (~rc & ~cc) | (rc & diag & loc_diag)
            vals = np.where(rc & diag & loc_diag, np.abs(vals), vals)
            A = sp.coo_matrix((vals[keep], (rows[keep], cols[keep])), shape=(self.n, self.n)).tocsr()
        if f is not None:
            rhs = np.zeros(self.n)
            fr = f.ravel().copy()
            fr[con[cd.ravel()] != 0] = 0.0
            np.add.at(rhs, cd.ravel(), fr)
        return A, rhs

    def solve(self, A, b):
        x = spla.spsolve(A.tocsc(), b)
        x[self.con != 0] = 0.0  # constraints.distribute
        return x

    def get_error(self, v):
        t = v.copy()
        t[self.con != 0] = 0.0
        return np.linalg.norm(t)

    def run(self):
        self.run_one_step(True)
        while self.prm.end_time - self.time > 1e-12:
            self.run_one_step(False)


class HyperElasticity(SolidBase):
    def __init__(self, mesh: fem.BoxMesh, params, verbose=False, material_id=None):
        """material_id: cell->material_id() of every cell (1-based part number); None = part 1 everywhere. With
        `Number of solid parts = 1` every cell uses part 1 whatever its id (mpi_hyper_elasticity.cpp:226-228)."""
        super().__init__(mesh, params, verbose)
        self.kirchhoff = params.solid_type == "Kirchhoff"  # PointHistory::setup (mpi_hyper_elasticity.cpp:8-35)
        ids = np.ones(mesh.n_cells, dtype=int) if material_id is None or params.n_solid_parts == 1 else np.asarray(material_id, dtype=int)
        self.material_id = ids
        self.update_qph(self.cur_u)

    # -- update_qph (:241-275) ------------------------------------------------
    def update_qph(self, u):
        dim = self.dim
        ue = u[self.dofs.cell_dofs].reshape(self.mesh.n_cells, self.npc, dim)  # [c][a][comp]
        grad_u = np.einsum("cai,cqak->cqik", ue, self.G)
        nc, nq = grad_u.shape[:2]
        self.F_inv, self.tau = np.empty((nc, nq, dim, dim)), np.empty((nc, nq, dim, dim))
        self.Jc, self.detF = np.empty((nc, nq) + (dim,) * 4), np.empty((nc, nq))
        for m in np.unique(self.material_id):  # the material of a point is that of its cell's part
            sel = self.material_id == m
            if self.kirchhoff:
                out = kirchhoff_update(grad_u[sel], self.prm.E[m - 1], self.prm.nu[m - 1])
            else:
                out = neo_hookean_update(grad_u[sel], self.prm.C[m - 1][0], self.prm.C[m - 1][1])
            self.F_inv[sel], self.tau[sel], self.Jc[sel], self.detF[sel] = out

    # -- SharedHyperElasticity::update_strain_and_stress (mpi_shared_hyper_elasticity.cpp:599-714) ------------
    def update_strain_and_stress(self):
        dim, nodes = self.dim, self.dofs.nodes
        Mref = np.einsum("qi,qj,q->ij", self.N, self.N, self.qw)
        qpt_to_dof = np.linalg.solve(Mref, (self.N * self.qw[:, None]).T)
        quad_stress = self.tau / self.detF[..., None, None]
        quad_strain = np.linalg.inv(self.F_inv)
        cs = np.einsum("aq,cqij->cija", qpt_to_dof, quad_stress)
        ce = np.einsum("aq,cqij->cija", qpt_to_dof, quad_strain)
        n = self.dofs.n_nodes
        stress, strain, count = np.zeros((dim * dim, n)), np.zeros((dim * dim, n)), np.zeros(n)
        np.add.at(count, nodes.ravel(), 1.0)
        for i in range(dim):
            for j in range(dim):
                np.add.at(stress[i * dim + j], nodes.ravel(), cs[:, i, j, :].ravel())
                np.add.at(strain[i * dim + j], nodes.ravel(), ce[:, i, j, :].ravel())
        self.stress, self.strain = stress / count, strain / count
        return self.stress, self.strain

    # -- assemble_system (:317-535) ---------------------------------------------
    def local_matrices(self, initial_step):
        dim, nc, npc, nq = self.dim, self.mesh.n_cells, self.npc, self.nq
        gamma = 0.5 + self.prm.damping
        beta = gamma / 2
        n = npc * dim
        # shape data per dof i = (a, c)
        # g[c,q,a,k] = G[c,q,a,m] F_inv[c,q,m,k]   (row c_i of grad_phi_i)
        g = np.einsum("cqam,cqmk->cqak", self.G, self.F_inv)
        I = np.eye(dim)
        # grad_phi[c,q,a,ci,r,k] = delta(r,ci) g[c,q,a,k]
        gp = np.einsum("ir,cqak->cqairk", I, g)
        sgp = 0.5 * (gp + np.swapaxes(gp, -1, -2))
        sgp = sgp.reshape(nc, nq, n, dim, dim)
        phi = np.einsum("qa,ir->qair", self.N, I).reshape(nq, n, dim)  # phi[q,i,r]
        w = self.JxW
        grav = np.asarray(self.prm.gravity[:dim], dtype=float)
        rhs = -np.einsum("cqirs,cqrs,cq->ci", sgp, self.tau, w) + self.rho * np.einsum("qir,r,cq->ci", phi, grav, w)
        if initial_step:
            K = self.rho * np.einsum("qir,qjr,cq->cij", phi, phi, w)
        else:
            K = (self.rho / (beta * self.dt ** 2)) * np.einsum("qir,qjr,cq->cij", phi, phi, w)
            K += np.einsum("cqirs,cqrstu,cqjtu,cq->cij", sgp, self.Jc, sgp, w, optimize=True)
            # geometric term for equal components: g_a . tau . g_b
            geo = np.einsum("cqak,cqkl,cqbl,cq->cab", g, self.tau, g, w, optimize=True)
            K += np.einsum("cab,ij->caibj", geo, I).reshape(nc, n, n)
        self.face_rhs(rhs)  # Neumann faces (:445-505)
        return K, rhs

    def assemble_system(self, initial_step):
        K, f = self.local_matrices(initial_step)
        A, rhs = self.scatter(K, f)
        if initial_step:
            self.mass_matrix = A
        else:
            self.system_matrix = A
        self.system_rhs = rhs
        return A, rhs

    # -- run_one_step (:83-207) --------------------------------------------------
    def run_one_step(self, first_step):
        p = self.prm
        gamma = 0.5 + p.damping
        beta = gamma / 2
        dt = self.dt
        if first_step:
            self.assemble_system(True)
            self.prev_a = self.solve(self.mass_matrix, self.system_rhs)
        self.time += dt
        self.timestep += 1
        pred = self.prev_u + dt * self.prev_v + (0.5 - beta) * dt * dt * self.prev_a
        nerr_u = nerr_f = 1.0
        err_u0 = err_f0 = 1.0
        it = 0
        shared_twin = p.simulation_type == "FSI"  # mpi_shared_hyper_elasticity.cpp:125-127
        err_u = 1.0
        while (nerr_u > p.tol_d or nerr_f > p.tol_f) and (not shared_twin or err_u > 1e-12):
            if it >= p.solid_max_iterations:
                raise RuntimeError("Too many Newton iterations!")
            self.cur_a = (self.cur_u - pred) / (beta * dt * dt)
            self.cur_v = self.prev_v + dt * (1 - gamma) * self.prev_a + dt * gamma * self.cur_a
            self.assemble_system(False)
            self.system_rhs = self.system_rhs - self.mass_matrix @ self.cur_a
            du = self.solve(self.system_matrix, self.system_rhs)
            err_f = self.get_error(self.system_rhs)
            if it == 0:
                err_f0 = err_f
            err_u = self.get_error(du)
            if it == 0:
                err_u0 = err_u
            # a solid at rest with no load has err_0 = 0: the reference divides all the same (mpi_hyper_elasticity.cpp:150-166), the
            # IEEE NaN compares false against the tolerances and the loop ends after this iteration - same here, without the warning
            with np.errstate(invalid="ignore", divide="ignore"):
                nerr_f = np.float64(err_f) / np.float64(err_f0)
                nerr_u = np.float64(err_u) / np.float64(err_u0)
            self.cur_u = self.cur_u + du
            self.update_qph(self.cur_u)
            self.history.append((self.timestep, it, err_f, err_u))
            if self.verbose:
                print(f"step {self.timestep} it {it} res_F {err_f:.3e} res_U {err_u:.3e}")
            it += 1
        self.cur_a = (self.cur_u - pred) / (beta * dt * dt)
        self.cur_v = self.prev_v + dt * (1 - gamma) * self.prev_a + dt * gamma * self.cur_a
        self.prev_a, self.prev_v, self.prev_u = self.cur_a.copy(), self.cur_v.copy(), self.cur_u.copy()
        self.update_strain_and_stress()


# ------------------------------------------------------------------------------------------------------------------
# Solid::MPI::LinearElasticity (source/mpi_linear_elasticity.cpp) and Solid::MPI::SharedLinearElasticity
# (source/mpi_shared_linear_elasticity.cpp), material source/linear_elastic_material.cpp:5-62
# ------------------------------------------------------------------------------------------------------------------
def linear_elastic_tensors(dim, E, nu, eta):
    """LinearElasticMaterial::get_elasticity / get_viscosity (linear_elastic_material.cpp:16-62)"""
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    mu = E / (2 * (1 + nu))
    I = np.eye(dim)
    ikjl, iljk, ijkl = np.einsum("ik,jl->ijkl", I, I), np.einsum("il,jk->ijkl", I, I), np.einsum("ij,kl->ijkl", I, I)
    return mu * (ikjl + iljk) + lam * ijkl, 0.5 * eta * (ikjl + iljk)


class LinearElasticity(SolidBase):
    """shared = False: Solid::MPI::LinearElasticity (p4est twin used stand-alone): gamma = 1/2 + damping, beta = gamma / 2,
                       system = M + beta dt^2 K (mpi_linear_elasticity.cpp:31-32, 96-121), Neumann faces only.
       shared = True : Solid::MPI::SharedLinearElasticity (the twin MPI::FSI drives): alpha = -damping, gamma = 1/2 - alpha,
                       beta = (1 + alpha)^2 / 4 in assemble_system but (1 - alpha)^2 / 4 in run_one_step - the reference's
                       own inconsistency, kept (mpi_shared_linear_elasticity.cpp:30-32, 305-307); viscous damping matrix
                       from eta; FSI traction on the deformed faces; nodal strain / stress recovery every step."""

    skip_dirichlet_faces = False

    def __init__(self, mesh: fem.BoxMesh, params, shared=False, verbose=False):
        super().__init__(mesh, params, verbose)
        self.shared = shared
        eta = params.eta[0] if getattr(params, "eta", None) else 0.0
        self.elasticity, self.viscosity = linear_elastic_tensors(self.dim, params.E[0], params.nu[0], eta)
        n_nodes = self.dofs.n_nodes
        self.stress = np.zeros((self.dim * self.dim, n_nodes))
        self.strain = np.zeros((self.dim * self.dim, n_nodes))

    def _forms(self):
        dim, nc, npc, nq = self.dim, self.mesh.n_cells, self.npc, self.nq
        n = npc * dim
        I = np.eye(dim)
        gp = np.einsum("ir,cqak->cqairk", I, self.G)  # grad phi_(a,i) [r][k] = delta(r,i) G[a][k]
        sgp = (0.5 * (gp + np.swapaxes(gp, -1, -2))).reshape(nc, nq, n, dim, dim)
        phi = np.einsum("qa,ir->qair", self.N, I).reshape(nq, n, dim)
        return sgp, phi

    def assemble_system(self, is_initial):
        p, dt, w = self.prm, self.dt, self.JxW
        sgp, phi = self._forms()
        grav = np.asarray(p.gravity[: self.dim], dtype=float)
        f = self.rho * np.einsum("qir,r,cq->ci", phi, grav, w)
        self.face_rhs(f)
        mass = lambda: self.rho * np.einsum("qir,qjr,cq->cij", phi, phi, w)
        stiff = lambda C: np.einsum("cqirs,rstu,cqjtu,cq->cij", sgp, C, sgp, w, optimize=True)
        if not self.shared:
            gamma = 0.5 + p.damping
            beta = gamma / 2
            if is_initial:
                self.system_matrix, self.system_rhs = self.scatter(mass(), f)
                self.stiffness_matrix = sp.csr_matrix((self.n, self.n))
            else:
                Ke = stiff(self.elasticity)
                self.system_matrix, self.system_rhs = self.scatter(mass() + beta * dt * dt * Ke, f)
                self.stiffness_matrix, _ = self.scatter(Ke, None)
            return
        alpha = -p.damping
        gamma = 0.5 - alpha
        beta = (1 + alpha) ** 2 / 4
        if is_initial:
            M, Ke, Ce = mass(), stiff(self.elasticity), stiff(self.viscosity)
            self.mass_matrix, _ = self.scatter(M, None)
            self.system_matrix, _ = self.scatter(M + Ce * gamma * dt * (1 + alpha) + Ke * beta * dt * dt * (1 + alpha), None)
            self.stiffness_matrix, _ = self.scatter(Ke, None)
            self.damping_matrix, _ = self.scatter(Ce, None)
        _, self.system_rhs = self.scatter(None, f)

    def run_one_step(self, first_step):
        p, dt = self.prm, self.dt
        if not self.shared:
            gamma = 0.5 + p.damping
            beta = gamma / 2
            if first_step:
                self.assemble_system(True)
                self.prev_a = self.solve(self.system_matrix, self.system_rhs)
                self.assemble_system(False)
            self.time += dt
            self.timestep += 1
            tmp2 = self.prev_u + dt * self.prev_v + (0.5 - beta) * dt * dt * self.prev_a
            tmp1 = self.system_rhs - self.stiffness_matrix @ tmp2
        else:
            alpha = -p.damping
            gamma = 0.5 - alpha
            beta = (1 - alpha) ** 2 / 4
            if first_step:
                self.assemble_system(True)
                self.prev_a = self.solve(self.mass_matrix, self.system_rhs)
            elif p.simulation_type == "FSI":
                self.assemble_system(False)
            self.time += dt
            self.timestep += 1
            tmp2 = self.prev_u + (1 + alpha) * dt * self.prev_v + (0.5 - beta) * dt * dt * (1 + alpha) * self.prev_a
            tmp4 = self.prev_v + (1 + alpha) * (1 - gamma) * dt * self.prev_a
            tmp1 = self.system_rhs - self.stiffness_matrix @ tmp2 - self.damping_matrix @ tmp4
        self.cur_a = self.solve(self.system_matrix, tmp1)
        self.cur_v = self.prev_v + dt * (1 - gamma) * self.prev_a + dt * gamma * self.cur_a
        self.cur_u = self.prev_u + dt * self.prev_v + dt * dt * (0.5 - beta) * self.prev_a + dt * dt * beta * self.cur_a
        self.prev_a, self.prev_v, self.prev_u = self.cur_a.copy(), self.cur_v.copy(), self.cur_u.copy()
        if self.shared:
            self.update_strain_and_stress()

    # -- SharedLinearElasticity::update_strain_and_stress (mpi_shared_linear_elasticity.cpp:401-531) ----------------
    def update_strain_and_stress(self):
        dim, nodes = self.dim, self.dofs.nodes
        Mref = np.einsum("qi,qj,q->ij", self.N, self.N, self.qw)
        qpt_to_dof = np.linalg.solve(Mref, (self.N * self.qw[:, None]).T)
        ue = self.cur_u[self.dofs.cell_dofs].reshape(self.mesh.n_cells, self.npc, dim)
        grad_u = np.einsum("cai,cqak->cqik", ue, self.G)
        quad_strain = 0.5 * (grad_u + np.swapaxes(grad_u, -1, -2))
        quad_stress = np.einsum("ijkl,cqkl->cqij", self.elasticity, quad_strain)
        cs = np.einsum("aq,cqij->cija", qpt_to_dof, quad_stress)
        ce = np.einsum("aq,cqij->cija", qpt_to_dof, quad_strain)
        n = self.dofs.n_nodes
        stress, strain, count = np.zeros((dim * dim, n)), np.zeros((dim * dim, n)), np.zeros(n)
        np.add.at(count, nodes.ravel(), 1.0)
        for i in range(dim):
            for j in range(dim):
                np.add.at(stress[i * dim + j], nodes.ravel(), cs[:, i, j, :].ravel())
                np.add.at(strain[i * dim + j], nodes.ravel(), ce[:, i, j, :].ravel())
        self.stress, self.strain = stress / count, strain / count
        return self.stress, self.strain
