// Fluid::MPI::InsIMEX<dim> on the device - see insimex.h. Reference: source/mpi_insimex.cpp.
#include "insimex.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <string>

namespace ifem
{
  namespace
  {
    struct SectionTimer
    {
      Context &ctx;
      double &acc;
      std::chrono::steady_clock::time_point t0;
      SectionTimer(Context &c, double &a) : ctx(c), acc(a)
      {
        cudaStreamSynchronize(ctx.stream);
        t0 = std::chrono::steady_clock::now();
      }
      ~SectionTimer()
      {
        cudaStreamSynchronize(ctx.stream);
        acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      }
    };
  } // namespace

  InsIMEX::InsIMEX(Context &ctx_, Triangulation &tria, const Parameters::AllParameters &params) : InsIM(ctx_, tria, params, true)
  {
    // BlockSchurPreconditioner::vmult (:56-133): CG for Mp 1e-6, CG for Sm 1e-3, CG for A max(1e-12, 1e-4 |src|), all
    // unpreconditioned; FGMRES to min(1e-9, 1e-8 |rhs|) (:370-371)
    control.a_inv_solver = 1;
    control.a_inv_rel = 1e-4;
    control.a_inv_floor = 1e-12;
    control.a_inv_max_it = 1 << 30; // capped by the number of velocity dofs in precondition()
    control.fgmres_rel = 1e-8;
    control.fgmres_floor = 1e-9;
    control.fgmres_floor_is_max = false;
  }

  void InsIMEX::assemble(bool use_nonzero_constraints, bool assemble_system)
  {
    SectionTimer t(ctx, timer_ms["Assemble system"]);
    if (fs.n_ranks > 1)
      {
        fs.halo_update(ctx, present_solution.p);
        fs.halo_update(ctx, fsi_acceleration.p);
      }
    InsAssembleParams p;
    p.viscosity = parameters.viscosity;
    p.gamma = parameters.grad_div;
    p.rho = parameters.fluid_rho;
    p.dt = time.get_delta_t();
    for (int d = 0; d < 3; ++d) p.gravity[d] = d < (int)parameters.gravity.size() ? parameters.gravity[d] : 0.0;
    p.explicit_convection = 1;
    p.rhs_only = assemble_system ? 0 : 1;
    // every field is evaluated at present_solution (:224-240): with evaluation point == present solution the
    // time-derivative term of the shared cell kernel vanishes and its right-hand side is the one of :272-291
    ins_assemble(ctx, fs, p, present_solution.p, present_solution.p, fsi_acceleration.p, use_nonzero_constraints, assemble_system);
  }

  std::pair<unsigned int, double> InsIMEX::solve(bool use_nonzero_constraints, bool assemble_system)
  {
    SectionTimer t(ctx, timer_ms["Solve linear system"]);
    if (assemble_system || !fs.schur_valid)
      {
        // BlockSchurPreconditioner ctor (:7-45): mass_schur = B diag(M_u)^-1 B^T. It depends on the mesh and on which dofs
        // are constrained only, so it is recomputed when the constraint set changed (as in InsIM::solve)
        if (!fs.schur_valid)
          {
            if (fs.n_ranks > 1)
              {
                InsAssembleParams p{};
                p.viscosity = parameters.viscosity;
                p.gamma = parameters.grad_div;
                p.rho = parameters.fluid_rho;
                p.dt = time.get_delta_t();
                p.explicit_convection = 1;
                ins_assemble(ctx, fs, p, present_solution.p, present_solution.p, fsi_acceleration.p, use_nonzero_constraints, true, true);
              }
            compute_mass_schur(ctx, fs);
            fs.schur_valid = true;
            sm_copy_valid = false;
          }
      }
    const VecSpace &va = fs.vs_all;
    const double nrm = nrm2(ctx, va, fs.rhs.p);
    const double tol = std::min(control.fgmres_floor, control.fgmres_rel * nrm);
    LinOp A = [&](const double *x, double *y) { block_vmult(ctx, fs, x, y); };
    LinOp P = [&](const double *x, double *y) { precondition(x, y); };
    const SolveResult r = fgmres(ctx, va, A, P, fs.rhs.p, newton_update.p, tol, n_dofs_global, control.basis_size, pool_fgmres);
    // constraints_used.distribute(solution_time_increment)
    set_flagged(ctx, fs.n_dofs, fs.d_con.p, use_nonzero_constraints ? fs.d_nonzero_val.p : nullptr, newton_update.p);
    return {(unsigned)r.iterations, r.residual};
  }

  void InsIMEX::run_one_step(bool apply_nonzero_constraints, bool assemble_system)
  {
    io_before_step();
    time.increment();
    if (verbose && fs.rank == 0)
      std::printf("%s\nTime step = %u, at t = %e\n", std::string(96, '*').c_str(), time.get_timestep(), time.current());
    const VecSpace &n = fs.vs_all;
    fill(ctx, n, 0.0, newton_update.p); // solution_time_increment = 0
    cur = NewtonRecord{};
    assemble(apply_nonzero_constraints, assemble_system);
    const auto state = solve(apply_nonzero_constraints, assemble_system);
    cur.abs_res = nrm2(ctx, n, fs.rhs.p);
    axpy(ctx, n, 1.0, newton_update.p, present_solution.p); // present_solution += solution_time_increment (:415-419)
    fs.halo_update(ctx, present_solution.p);
    cur.timestep = time.get_timestep();
    cur.iteration = 0;
    cur.rel_res = 1.0;
    cur.gmres_its = (int)state.first;
    cur.gmres_res = state.second;
    history.push_back(cur);
    if (verbose && fs.rank == 0)
      std::printf(" GMRES_ITR = %-3u GMRES_RES = %e  [cg_mp %d cg_sm %d cg_a %d / %d]\n", state.first, state.second, cur.cg_mp_its,
                  cur.cg_sm_its, cur.a_inv_its, cur.precond_applies);
    update_stress(); // :425
    io_after_step();
  }

  void InsIMEX::run()
  {
    const bool success_load = load_checkpoint(); // :455; false unless an output directory is set
    if (!dofs_ready)
      {
        triangulation.refine_global(parameters.global_refinements.empty() ? 0 : parameters.global_refinements[0]);
        setup_dofs();
        make_constraints();
        initialize_system();
      }
    // nonzero constraints at the very first time step only; the left-hand side is assembled twice: once with the
    // nonzero, once with the zero constraints (:471-479)
    while (time.end() - time.current() > 1e-12) run_one_step(time.get_timestep() == 0, time.get_timestep() < 2 || success_load);
  }
} // namespace ifem
