"""TEST INFRASTRUCTURE - NOT PRODUCT CODE. One rank of a slab-partitioned run on the emulated device (see cuda_runtime.h /
comm_emul.cpp in this directory): what tests/test_ins_multigpu.py's worker does on a GPU, for any of the fluid solvers.
    python multirank_case.py <rank> <size> <rendezvous dir> <out.npz> <solver> <dim> <reps...>
Saves the owned parts of: a block mat-vec of a global vector, the right-hand side of an assembly at a global state, the
solution after two time steps from rest, and the Newton history."""
import ctypes
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]

import build_emulated  # noqa: E402
import openifem_b200._lib as product_loader  # noqa: E402

handle = ctypes.CDLL(build_emulated.build(), mode=ctypes.RTLD_GLOBAL)
handle.ifem_last_error.restype = ctypes.c_char_p
product_loader._lib = handle

import openifem_b200 as ifem  # noqa: E402
from util import cavity_prm  # noqa: E402


def main():
    rank, size, rdv, out, solver, dim = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5], int(sys.argv[6])
    reps = tuple(int(a) for a in sys.argv[7:7 + dim])
    ifem.init(0)
    if size > 1:
        idfile = os.path.join(rdv, "unique_id")
        if rank == 0:
            with open(idfile + ".tmp", "wb") as f:
                f.write(ifem.comm_unique_id())
            os.replace(idfile + ".tmp", idfile)
        else:
            while not os.path.exists(idfile):
                time.sleep(0.02)
        ifem.comm_init(rank, size, open(idfile, "rb").read())
    if solver == "FSI":
        return fsi_case(rank, size, out, dim, reps)
    if solver == "OUTPUT":
        return output_case(rank, size, out, dim, reps)
    if solver == "CKPT":
        return checkpoint_case(rank, size, out, dim, reps)
    if solver == "ACOUSTIC":
        return acoustic_case(rank, size, out)
    tria = ifem.Triangulation(dim)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0,) * dim, (1,) * dim, True)
    q1 = solver in ("SCnsIM", "SUPGInsIM")
    refined = "refined" in sys.argv[7 + dim:]
    if refined:  # a band across the last axis refined once: hanging nodes on two planes (the meshes of BASELINE configs 4 / 5)
        v, c, _ = tria.get_mesh()
        z = v[c].mean(axis=1)[:, dim - 1]
        tria.execute_refinement(((z > 0.34) & (z < 0.67)).astype(np.uint8))
    q2 = "q2" in sys.argv[7 + dim:]
    with_sa = "sa" in sys.argv[7 + dim:]  # Spalart-Allmaras model attached (walls on boundary ids 2 / 3, inflow on 0)
    if solver == "SCnsIM":
        from test_scns_gpu import scns_prm

        text = scns_prm(dim, dt=1e-3)
        if with_sa:
            text = scns_prm(dim, dt=1e-2, mu=1e-3, rho=1.0) + (
                "subsection Spalart Allmaras model\n  set Number of S-A model BCs = 3\n  set S-A model boundary id = 0, 2, 3\n"
                "  set S-A model boundary types = 1, 0, 0\n  set Initial condition coefficient = 3.0\nend\n")
        if q2:  # Taylor-Hood pair through the degree-generic kernel (csrc/scnsim_generic.cu)
            text = text.replace("set Velocity degree = 1", "set Velocity degree = 2")
            q1 = False
    elif solver == "SUPGInsIM":
        full = 3 if dim == 2 else 7
        text = cavity_prm(dim, newton_tol=1e-8, dirichlet={2: (full, [0.0] * dim), 3: (full, [0.5] + [0.0] * (dim - 1))}, neumann={0: 1.0})
        text = text.replace("set Velocity degree = 2", "set Velocity degree = 1")
    else:
        text = cavity_prm(dim, newton_tol=1e-9)
    flow = getattr(ifem.Fluid.MPI, solver)(tria, ifem.Parameters.AllParameters(text=text))
    if solver == "SCnsIM":
        flow.set_body_force(lambda p, c: 5.0 if c == 0 else 0.0)
    flow.setup()
    model = flow.attach_turbulence_model("Spalart-Allmaras") if with_sa else None
    if solver == "InsIM" and "inner32" in sys.argv[7 + dim:]:
        # device-resident fp32 inner solvers (BiCGStab on the fp16 SELL copy of A_uu, CG on the SELL copy of S_m): on emulated ranks
        # their reductions and halos go through the communicator between the kernels (no peer memory there)
        flow.set_control(a_inv_rel=1e-3, a_inv_max_it=500, fgmres_rel=1e-9, a_inv_fp32=3, cg_sm_fp32=1)
    elif solver == "InsIM":
        flow.set_control(a_inv_rel=1e-10, a_inv_max_it=5000, fgmres_rel=1e-9)
    elif solver in ("SCnsIM", "SUPGInsIM"):
        flow.set_control(fgmres_rel=1e-10, supg_ilu=0)
    n_un_glob = int(np.prod([(1 if q1 else 2) * k + 1 for k in reps]))
    n_pn_glob = int(np.prod([k + 1 for k in reps]))
    if refined:
        n_un_glob = n_pn_glob = tria.n_vertices()
    n_glob = dim * n_un_glob + n_pn_glob
    loc, glo = flow.owned_global_dofs(n_un_glob)
    gu, gp = flow.local_to_global(0).astype(np.int64), flow.local_to_global(1).astype(np.int64)
    xg = np.sin(0.11 * np.arange(n_glob)) + 0.3
    ev = 0.1 * np.cos(0.05 * np.arange(n_glob))

    def localise(vg):
        vu = vg[(gu[:, None] * dim + np.arange(dim)[None, :]).ravel()]
        return np.concatenate([vu, vg[dim * n_un_glob + gp]])

    flow.set_vector(flow.EVALUATION_POINT, localise(ev))
    flow.set_vector(flow.PRESENT, localise(0.5 * ev))
    flow.assemble(True)
    y = flow.vmult(localise(xg))
    rhs = flow.get_vector(flow.SYSTEM_RHS)
    extra = {}
    if model is not None:  # assembly of the transport system at a global state (the fluid's present_solution is 0.5 ev)
        nu0 = 1e-3 * (1.5 + np.sin(0.37 * np.arange(n_pn_glob)))
        model.set_vector(model.PRESENT, nu0[gp])
        model.set_vector(model.EVALUATION_POINT, 1.1 * nu0[gp])
        model.assemble(True)
        n_own_p = flow.partition(1)[0]
        extra = dict(glo_p=gp[:n_own_p], sa_rhs=model.get_vector(model.SYSTEM_RHS)[:n_own_p])
        model.set_vector(model.PRESENT, np.full(gp.size, 3e-3))
        model.update_boundary_condition(True)  # restores the boundary lines (no cell is inside a solid here)
    zero = np.zeros(flow.n_dofs)
    flow.set_vector(flow.EVALUATION_POINT, zero)
    flow.set_vector(flow.PRESENT, zero)
    for k in range(2):
        if model is not None:
            model.run_one_step(k == 0)
        if solver == "InsIMEX":
            flow.run_one_step(k == 0, k < 2)
        else:
            flow.run_one_step(k == 0)
    sol = flow.get_current_solution()
    hist = np.array([(h["timestep"], h["iteration"], h["abs_res"], h["gmres_its"]) for h in flow.history()], dtype=np.float64)
    if model is not None:
        extra["nu"] = model.get_vector(model.PRESENT)[:extra["glo_p"].size]
    np.savez(out, glo=glo, y=y[loc], rhs=rhs[loc], sol=sol[loc], hist=hist, n_u=dim * n_un_glob, **extra)
    if size > 1:
        ifem.comm_finalize()


def output_case(rank, size, out, dim, reps):
    """every rank writes its piece of fluid_000003 for a solution that is a known function of the position"""
    tria = ifem.Triangulation(dim)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0,) * dim, (1,) * dim, True)
    flow = ifem.Fluid.MPI.InsIM(tria, ifem.Parameters.AllParameters(text=cavity_prm(dim)))
    flow.setup()
    pts = flow.support_points()
    n_u = flow.n_u
    present = np.empty(flow.n_dofs)
    for c in range(dim):
        present[c:n_u:dim] = (c + 1) * pts[c:n_u:dim, 0] - 0.5 * pts[c:n_u:dim, 1]
    present[n_u:] = 3.0 + pts[n_u:, 0] * pts[n_u:, 1]
    flow.set_vector(flow.PRESENT, present)
    flow.set_output_directory(os.path.dirname(out))
    flow.output_results(3)
    np.savez(out, glo=np.zeros(0), y=np.zeros(0), rhs=np.zeros(0), sol=np.zeros(0), hist=np.zeros((0, 4)), n_u=0)
    if size > 1:
        ifem.comm_finalize()


def checkpoint_case(rank, size, out, dim, reps):
    """InsIM::run with checkpoints: argv[-2] = number of time steps, argv[-1] = output directory or "-" (none)"""
    n_steps, directory = int(sys.argv[-2]), sys.argv[-1]
    text = cavity_prm(dim, dt=1e-2, end_time=n_steps * 1e-2).replace("set Save interval = 1e6", "set Save interval = 0.02")
    tria = ifem.Triangulation(dim)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0,) * dim, (1,) * dim, True)
    flow = ifem.Fluid.MPI.InsIM(tria, ifem.Parameters.AllParameters(text=text))
    if directory != "-":
        flow.set_output_directory(directory)
    flow.set_control(a_inv_rel=1e-10, a_inv_max_it=5000, fgmres_rel=1e-10)
    flow.run()
    n_un_glob = int(np.prod([2 * k + 1 for k in reps]))
    loc, glo = flow.owned_global_dofs(n_un_glob)
    sol = flow.get_current_solution()
    hist = np.array([(h["timestep"], h["iteration"], h["abs_res"], h["gmres_its"]) for h in flow.history()], dtype=np.float64)
    np.savez(out, glo=glo, y=np.zeros(glo.size), rhs=np.zeros(glo.size), sol=sol[loc], hist=hist, n_u=dim * n_un_glob)
    if size > 1:
        ifem.comm_finalize()


def acoustic_case(rank, size, out):
    """the reference's acoustic_duct_wave_mpi case (tests/acoustic_cases.py), first 20 steps through SCnsIM::run: the
    time-dependent hard-coded boundary value re-makes the constraints on every rank in every step"""
    import acoustic_cases

    c = acoustic_cases.CASES["duct"]
    text = acoustic_cases.prm_text("duct", 20).replace("set Global refinements = 3, 0", "set Global refinements = 2, 0")
    tria = ifem.Triangulation(2)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, c["reps"], (0, 0), c["hi"], True)
    flow = ifem.Fluid.MPI.SCnsIM(tria, ifem.Parameters.AllParameters(text=text))
    flow.add_hard_coded_boundary_condition(0, acoustic_cases.gaussian_pulse("duct", 1e-7))
    flow.set_control(fgmres_rel=1e-10, supg_ilu=0)
    flow.run()
    n_un_glob = (c["reps"][0] * 4 + 1) * (c["reps"][1] * 4 + 1)
    loc, glo = flow.owned_global_dofs(n_un_glob)
    sol = flow.get_current_solution()
    hist = np.array([(h["timestep"], h["iteration"], h["abs_res"], h["gmres_its"]) for h in flow.history()], dtype=np.float64)
    np.savez(out, glo=glo, y=np.zeros(glo.size), rhs=np.zeros(glo.size), sol=sol[loc], hist=hist, n_u=2 * n_un_glob)
    if size > 1:
        ifem.comm_finalize()


def fsi_case(rank, size, out, dim, reps):
    """two passes of the FSI::run loop (tests/test_fsi_gpu.py test_coupled_fsi_steps_match_oracle's problem): partitioned SCnsIM
    fluid, replicated NeoHookean solid, coupling with the indicator / acceleration variant"""
    from test_fsi_gpu import _fsi_text

    params = ifem.Parameters.AllParameters(text=_fsi_text(dim))
    ftria, stria = ifem.Triangulation(dim), ifem.Triangulation(dim)
    ifem.GridGenerator.subdivided_hyper_rectangle(ftria, reps, (0.0,) * dim, (1.0,) * dim, True)
    if dim == 2:
        ifem.GridGenerator.subdivided_hyper_rectangle(stria, (4, 6), (0.3125, 0.0), (0.5625, 0.6875), True)
    else:
        ifem.GridGenerator.subdivided_hyper_rectangle(stria, (3, 3, 4), (0.25, 0.0, 0.25), (0.7, 0.6, 0.75), True)
    fluid, solid = ifem.Fluid.MPI.SCnsIM(ftria, params), ifem.Solid.MPI.HyperElasticity(stria, params)
    fluid.setup()
    solid.setup()
    fluid.set_control(fgmres_rel=1e-10, supg_ilu=0)
    coupling = ifem.MPI.FSI(fluid, solid, params, sys.argv[-1] == "dirichlet")
    if "refine" in sys.argv[7 + dim:]:
        # FSI::refine_mesh around the solid before and between the steps: new partition, transferred solution (every rank holds the
        # whole triangulation and refines it identically; the slabs avoid the planes that carry hanging nodes)
        coupling.refine_mesh(0, 2)
        fluid.set_control(fgmres_rel=1e-10, supg_ilu=0)
        coupling.run_one_step(True)
        coupling.refine_mesh(0, 2)
        coupling.run_one_step(False)
    else:
        for k in range(2):
            coupling.run_one_step(k == 0)
    n_un_glob = ftria.n_vertices()
    loc, glo = fluid.owned_global_dofs(n_un_glob)
    sol = fluid.get_current_solution()
    acc = fluid.get_vector(fluid.FSI_ACCELERATION)
    hist = np.array([(h["timestep"], h["iteration"], h["abs_res"], h["gmres_its"]) for h in fluid.history()], dtype=np.float64)
    np.savez(out, glo=glo, y=acc[loc], rhs=fluid.get_vector(fluid.SYSTEM_RHS)[loc], sol=sol[loc], hist=hist, n_u=dim * n_un_glob,
             solid=solid.get_current_solution())
    if size > 1:
        ifem.comm_finalize()


if __name__ == "__main__":
    main()
