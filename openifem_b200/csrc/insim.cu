#include "insim.h"

#include <unordered_map>

#include "comm.h"

#include <chrono>
#include <cstdio>

namespace ifem
{
  namespace
  {
    struct ScopedTimer
    {
      Context &ctx;
      double &acc;
      std::chrono::steady_clock::time_point t0;
      ScopedTimer(Context &c, double &a) : ctx(c), acc(a)
      {
        cudaStreamSynchronize(ctx.stream);
        t0 = std::chrono::steady_clock::now();
      }
      ~ScopedTimer()
      {
        cudaStreamSynchronize(ctx.stream);
        acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      }
    };
  } // namespace

  InsIM::InsIM(Context &ctx_, Triangulation &tria, const Parameters::AllParameters &params, bool taylor_hood_only)
    : ctx(ctx_), triangulation(tria), parameters(params),
      time(params.end_time, params.time_step, params.output_interval, params.refinement_interval, params.save_interval)
  {
    if (!taylor_hood_only) return;
    // mpi_insim.cpp:135-139
    if (parameters.fluid_velocity_degree - parameters.fluid_pressure_degree != 1)
      throw std::runtime_error("Velocity finite element should be one order higher than pressure!");
    if (parameters.fluid_velocity_degree != 2) throw std::runtime_error("InsIM: only Q2/Q1 is implemented on the device");
  }

  void InsIM::add_hard_coded_boundary_condition(int id, std::function<double(const double *, unsigned int, double)> f)
  {
    hard_coded[id] = std::move(f);
    fs.base_valid = false;
  }

  void InsIM::setup_dofs()
  {
    fs.setup(ctx, triangulation, (int)parameters.fluid_velocity_degree, (int)parameters.fluid_pressure_degree, false);
    dofs_ready = true;
  }

  void InsIM::make_constraints()
  {
    // the lines depend on the mesh, the .prm maps and - through hard-coded boundary functions - on the functions' clock only:
    // unchanged inputs restore the device copies (two device-to-device copies; the host mirrors fs.con / fs.nonzero_val keep
    // these base lines as well)
    if (fs.base_valid && (hard_coded.empty() || base_bc_time == bc_time))
      {
        fs.restore_base_constraints(ctx);
        if (after_make_constraints) after_make_constraints();
        return;
      }
    base_bc_time = bc_time;
    std::function<bool(int, const double *, int, double &)> hc;
    if (!hard_coded.empty())
      hc = [this](int id, const double *pt, int c, double &v) {
        auto it = hard_coded.find(id);
        if (it == hard_coded.end()) return false;
        v = it->second(pt, (unsigned)c, bc_time);
        return true;
      };
    fs.make_constraints(ctx, triangulation, parameters.fluid_dirichlet_bcs, hc);
    fs.set_neumann_faces(ctx, triangulation, parameters.fluid_neumann_bcs);
    if (after_make_constraints) after_make_constraints();
  }

  void InsIM::upload_constraints()
  {
    fs.d_con.upload(fs.con, ctx.stream);
    fs.d_nonzero_val.upload(fs.nonzero_val, ctx.stream);
    fs.schur_valid = false;
    fs.flags_merged = true;
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  namespace
  {
    // node of a Q1 table at a position (the tables are renumbered spatially: vertices are matched by their coordinates)
    struct CoordLookup
    {
      int dim;
      double lo[3], ext[3];
      std::unordered_map<uint64_t, int> at;
      uint64_t key(const double *x) const
      {
        uint64_t k = 0;
        for (int d = dim - 1; d >= 0; --d) k = (k << 21) | (uint64_t)std::llround((x[d] - lo[d]) / ext[d] * double((1u << 21) - 2));
        return k;
      }
      CoordLookup(int dim_, const std::vector<double> &box_pts, const std::vector<double> &coords) : dim(dim_)
      {
        for (int d = 0; d < dim; ++d)
          {
            double a = box_pts[d], b = box_pts[d];
            for (size_t i = 0; i < box_pts.size() / dim; ++i)
              {
                a = std::min(a, box_pts[i * dim + d]);
                b = std::max(b, box_pts[i * dim + d]);
              }
            lo[d] = a;
            ext[d] = b > a ? b - a : 1.0;
          }
        for (size_t i = 0; i < coords.size() / dim; ++i) at.emplace(key(&coords[i * dim]), (int)i);
      }
      int find(const double *x) const
      {
        auto it = at.find(key(x));
        if (it == at.end()) throw std::runtime_error("refine_mesh: a vertex has no node at its position");
        return it->second;
      }
    };
  } // namespace

  void InsIM::after_mesh_change(const Triangulation::TransferPlan &plan, const std::vector<double> &old_vertices)
  {
    if (fs.pu != 1 || fs.pp != 1) throw std::runtime_error("refine_mesh: solution transfer is implemented for FE_Q(1) velocity and pressure");
    const int dim = fs.dim;
    // old solution per old vertex: every rank contributes the nodes it owns (the triangulation is replicated, the solution is not)
    const int64_t n_old_u = fs.n_u;
    const std::vector<double> old = present_solution.to_host(ctx.stream);
    const int nv_old = (int)(old_vertices.size() / dim);
    std::vector<std::vector<double>> old_extra;
    for (DevBuf<double> *f : transferred_scalar_fields()) old_extra.push_back(f->to_host(ctx.stream));
    const int stride = dim + 1 + (int)old_extra.size(); // per vertex: velocity, pressure, the extra scalar fields
    std::vector<double> vert_val((size_t)nv_old * stride, 0.0);
    {
      const CoordLookup vertex_at(dim, old_vertices, old_vertices);
      for (int l = 0; l < fs.n_owned_unodes; ++l)
        {
          const int v = vertex_at.find(&fs.un.coords[(size_t)l * dim]);
          for (int c = 0; c < dim; ++c) vert_val[(size_t)v * stride + c] = old[(size_t)dim * l + c];
        }
      for (int l = 0; l < fs.n_owned_pnodes; ++l)
        {
          const size_t v = (size_t)vertex_at.find(&fs.pn.coords[(size_t)l * dim]);
          vert_val[v * stride + dim] = old[(size_t)n_old_u + l];
          for (size_t e = 0; e < old_extra.size(); ++e) vert_val[v * stride + dim + 1 + e] = old_extra[e][l];
        }
      if (fs.n_ranks > 1)
        {
          DevBuf<double> d(vert_val.size());
          d.upload(vert_val, ctx.stream);
          for (size_t off = 0; off < vert_val.size(); off += (size_t)1 << 28)
            comm_allreduce_sum(*ctx.comm, d.p + off, (int)std::min<size_t>(vert_val.size() - off, (size_t)1 << 28), ctx.stream);
          vert_val = d.to_host(ctx.stream);
        }
    }
    // new spaces (on several ranks: a new partition of the new mesh)
    fs.base_valid = false;
    if (on_mesh_change) on_mesh_change();
    setup_dofs();
    make_constraints();
    initialize_system();
    const int nv_new = triangulation.n_vertices();
    if ((int64_t)plan.ptr.size() != (int64_t)nv_new + 1) throw std::runtime_error("refine_mesh: the transfer plan does not belong to this mesh");
    std::vector<double> fresh((size_t)fs.n_dofs, 0.0);
    const CoordLookup vertex_at(dim, triangulation.vertices, triangulation.vertices);
    auto value = [&](const double *x, int c) {
      const int v = vertex_at.find(x);
      double val = 0.0;
      for (int64_t k = plan.ptr[v]; k < plan.ptr[v + 1]; ++k) val += plan.weight[k] * vert_val[(size_t)plan.old_vertex[k] * stride + c];
      return val;
    };
    for (int l = 0; l < fs.un.n_nodes; ++l) // owned and ghost nodes alike
      for (int c = 0; c < dim; ++c) fresh[(size_t)dim * l + c] = value(&fs.un.coords[(size_t)l * dim], c);
    for (int l = 0; l < fs.pn.n_nodes; ++l) fresh[(size_t)fs.n_u + l] = value(&fs.pn.coords[(size_t)l * dim], dim);
    present_solution.upload(fresh, ctx.stream);
    const std::vector<DevBuf<double> *> extra = transferred_scalar_fields();
    for (size_t e = 0; e < extra.size() && e < old_extra.size(); ++e)
      {
        std::vector<double> field((size_t)fs.pn.n_nodes);
        for (int l = 0; l < fs.pn.n_nodes; ++l) field[l] = value(&fs.pn.coords[(size_t)l * dim], dim + 1 + (int)e);
        extra[e]->upload(field, ctx.stream);
      }
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void InsIM::initialize_system()
  {
    const int64_t n = fs.n_dofs;
    for (DevBuf<double> *v : {&present_solution, &evaluation_point, &solution_increment, &newton_update, &fsi_acceleration})
      {
        v->alloc(n);
        v->zero(ctx.stream);
      }
    d_binv.alloc((size_t)fs.n_owned_unodes * fs.dim * fs.dim);
    stress.alloc((size_t)fs.dim * fs.dim * fs.un.n_nodes);
    stress.zero(ctx.stream);
    d_tmp_p.alloc(fs.n_p);
    d_utmp.alloc(fs.n_u);
    if (fs.n_ranks > 1) d_utmp2.alloc(fs.n_u);
    // global sizes (iteration caps of the reference are the global matrix dimensions)
    n_dofs_global = fs.vs_all.n_owned();
    n_p_global = fs.vs_p.n_owned();
    if (fs.n_ranks > 1)
      {
        DevBuf<double> cnt(2);
        const double h[2] = {(double)n_dofs_global, (double)n_p_global};
        cnt.upload(h, 2, ctx.stream);
        comm_allreduce_sum(*ctx.comm, cnt.p, 2, ctx.stream);
        const std::vector<double> g = cnt.to_host(ctx.stream);
        n_dofs_global = (int64_t)g[0];
        n_p_global = (int64_t)g[1];
      }
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void InsIM::assemble(bool use_nonzero_constraints)
  {
    ScopedTimer t(ctx, timer_ms["Assemble system"]);
    if (fs.n_ranks > 1)
      {
        // the gather in the cell loop reads ghosted vectors (get_function_values on ghosted PETSc vectors)
        fs.halo_update(ctx, evaluation_point.p);
        fs.halo_update(ctx, present_solution.p);
        fs.halo_update(ctx, fsi_acceleration.p);
      }
    InsAssembleParams p;
    p.viscosity = parameters.viscosity;
    p.gamma = parameters.grad_div;
    p.rho = parameters.fluid_rho;
    p.dt = time.get_delta_t();
    for (int d = 0; d < 3; ++d) p.gravity[d] = d < (int)parameters.gravity.size() ? parameters.gravity[d] : 0.0;
    ins_assemble(ctx, fs, p, evaluation_point.p, present_solution.p, fsi_acceleration.p, use_nonzero_constraints, true);
  }

  // BlockSchurPreconditioner::vmult (mpi_insim.cpp:56-128)
  void InsIM::precondition(const double *src, double *dst)
  {
    const int64_t n_u = fs.n_u;
    const VecSpace &vu = fs.vs_u, &vp = fs.vs_p;
    const double *src_u = src, *src_p = src + n_u;
    double *dst_u = dst, *dst_p = dst + n_u;
    double *tmp = d_tmp_p.p, *utmp = d_utmp.p;
    const double nrm = nrm2(ctx, vp, src_p);
    const int max_p_its = (int)std::min<int64_t>(n_p_global, 1 << 30);
    {
      ScopedTimer t(ctx, timer_ms["CG for Mp"]);
      LinOp Mp = [&](const double *x, double *y) {
        fs.halo_p.update(ctx, const_cast<double *>(x));
        spmv(ctx, fs.M_p, x, y);
      };
      const SolveResult r = cg_mp_dev.solve(ctx, vp, Mp, src_p, tmp, std::max(control.cg_floor, control.cg_mp_rel * nrm), max_p_its);
      cur.cg_mp_its += r.iterations;
      scale(ctx, vp, -(parameters.viscosity + parameters.grad_div * parameters.fluid_rho), tmp);
    }
    {
      ScopedTimer t(ctx, timer_ms["CG for Sm"]);
      const SolveResult r = solve_mass_schur(control.cg_sm_fp32, src_p, nrm, dst_p, std::max(control.cg_floor, control.cg_sm_rel * nrm), max_p_its);
      cur.cg_sm_its += r.iterations;
      // dst_p = -rho/dt * dst_p + tmp
      axpby(ctx, vp, 1.0, tmp, -parameters.fluid_rho / time.get_delta_t(), dst_p);
    }
    // utmp = src_u - B^T dst_p
    fs.halo_p.update(ctx, dst_p);
    spmv(ctx, fs.A_up, dst_p, utmp);
    axpby(ctx, vu, 1.0, src_u, -1.0, utmp);
    {
      ScopedTimer t(ctx, timer_ms["A_inv"]);
      LinOp Auu = [&](const double *x, double *y) {
        fs.halo_u.update(ctx, const_cast<double *>(x));
        if (control.a_inv_fp32 == 1) spmv_fp32(ctx, fs.A_uu, x, y); else spmv(ctx, fs.A_uu, x, y);
      };
      LinOp jac = [&](const double *x, double *y) { block_diag_apply(ctx, fs.n_owned_unodes, fs.dim, d_binv.p, x, y); };
      const double unrm = nrm2(ctx, vu, utmp);
      if (control.a_inv_solver == 1)
        {
          fill(ctx, vu, 0.0, dst_u);
          const int max_u_its = (int)std::min<int64_t>(n_dofs_global - n_p_global, (int64_t)control.a_inv_max_it);
          const SolveResult r = cg(ctx, vu, Auu, utmp, dst_u, true, std::max(control.a_inv_floor, control.a_inv_rel * unrm), max_u_its, pool_ainv);
          cur.a_inv_its += r.iterations;
          cur.precond_applies++;
          return;
        }
      const SolveResult r = control.a_inv_fp32 >= 2
                              ? inner32.solve(ctx, utmp, unrm, dst_u, control.a_inv_rel, control.a_inv_max_it)
                              : bicgstab(ctx, vu, Auu, jac, utmp, dst_u, control.a_inv_rel * unrm, control.a_inv_max_it, pool_ainv);
      cur.a_inv_its += r.iterations;
      timer_ms["A_inv block-Jacobi fallbacks (count)"] = (double)inner32.n_fallbacks; // readable through ifem_insim_timer_ms
    }
    cur.precond_applies++;
  }

  // "CG for Sm" (mpi_insim.cpp:88-109): x = S_m^-1 b to |r| <= tol_abs, x0 = 0
  SolveResult InsIM::solve_mass_schur(int mode, const double *b, double b_norm, double *x, double tol_abs, int max_it)
  {
    const VecSpace &vp = fs.vs_p;
    if (mode == 0)
      {
        fill(ctx, vp, 0.0, x);
        LinOp Sm = [&](const double *in, double *out) {
          fs.halo_p.update(ctx, const_cast<double *>(in)); // both ghost layers: S_m reaches two cells deep
          spmv(ctx, fs.S_m, in, out);
        };
        return cg(ctx, vp, Sm, b, x, true, tol_abs, max_it, pool_cg);
      }
    const int precision = mode == 2 ? 16 : 32;
    if (!inner_sm.S.built() || inner_sm.S.precision != precision)
      {
        inner_sm.setup(ctx, fs.S_m, fs.pn, fs.n_ranks > 1 ? &fs.halo_p : nullptr, precision);
        sm_copy_valid = false;
      }
    if (!sm_copy_valid)
      {
        inner_sm.refresh(ctx, fs.S_m);
        // aggregates of the coarse space from the bounding box of the whole mesh (every rank holds the triangulation)
        for (int d = 0; d < fs.dim; ++d)
          {
            double lo = triangulation.vertices[d], hi = lo;
            for (int i = 0; i < triangulation.n_vertices(); ++i)
              {
                lo = std::min(lo, triangulation.vertices[(size_t)i * fs.dim + d]);
                hi = std::max(hi, triangulation.vertices[(size_t)i * fs.dim + d]);
              }
            inner_sm.box[2 * d] = lo;
            inner_sm.box[2 * d + 1] = hi;
          }
        inner_sm.build_coarse(ctx, fs.S_m, fs.pn);
        sm_copy_valid = true;
      }
    // fp32 CG occasionally stagnates on the (singular, in closed cavities) S_m where fp64 CG converges - observed once in
    // ~200 applications at config 3. The fast path is therefore capped and verified: if it has not reached the tolerance
    // after 1500 iterations (typical: 150-600) the application is redone by the fp64 CG on the CSR matrix.
    SolveResult r = inner_sm.solve(ctx, b, b_norm, x, tol_abs, std::min(max_it, 1500));
    if (!r.converged || !std::isfinite(r.residual))
      {
        const int spent = r.iterations;
        r = solve_mass_schur(0, b, b_norm, x, tol_abs, max_it);
        r.iterations += spent;
        sm_fallbacks++;
        timer_ms["CG for Sm fp64 fallbacks (count)"] += 1.0; // readable through ifem_insim_timer_ms
      }
    return r;
  }

  std::pair<unsigned int, double> InsIM::solve(bool use_nonzero_constraints)
  {
    ScopedTimer t(ctx, timer_ms["Solve linear system"]);
    // BlockSchurPreconditioner ctor (mpi_insim.cpp:13-50)
    if (!fs.schur_valid)
      {
        // B = A_pu, B^T = A_up and diag(M_u) depend on the mesh and on WHICH dofs are constrained only, so the
        // reference's per-solve mmult (mpi_insim.cpp:48-49) is hoisted: recomputed when the constraints change
        if (fs.n_ranks > 1)
          {
            InsAssembleParams p{};
            p.viscosity = parameters.viscosity;
            p.gamma = parameters.grad_div;
            p.rho = parameters.fluid_rho;
            p.dt = time.get_delta_t();
            ins_assemble(ctx, fs, p, evaluation_point.p, present_solution.p, fsi_acceleration.p, use_nonzero_constraints, true, true);
          }
        compute_mass_schur(ctx, fs);
        fs.schur_valid = true;
        sm_copy_valid = false;
      }
    block_diag_inverse(ctx, fs.A_uu, d_binv.p);
    if (control.a_inv_fp32 >= 2)
      {
        const int precision = control.a_inv_fp32 == 3 ? 16 : 32;
        if (!inner32.S.built() || inner32.S.precision != precision)
          inner32.setup(ctx, fs.A_uu, fs.un, fs.n_ranks > 1 ? &fs.halo_u : nullptr, precision);
        inner32.refresh(ctx, fs.A_uu, d_binv.p);
      }
    else if (control.a_inv_fp32 == 1)
      make_fp32_copy(ctx, fs.A_uu);
    const VecSpace &va = fs.vs_all;
    const double nrm = nrm2(ctx, va, fs.rhs.p);
    const double tol = std::max(control.fgmres_floor, control.fgmres_rel * nrm);
    LinOp A = [&](const double *x, double *y) { block_vmult(ctx, fs, x, y); };
    LinOp P = [&](const double *x, double *y) { precondition(x, y); };
    const SolveResult r = fgmres(ctx, va, A, P, fs.rhs.p, newton_update.p, tol, n_dofs_global, control.basis_size, pool_fgmres);
    {
      // the pin of the inexact inner solves: residual of the returned update under the fp64 operator
      double *res = pool_fgmres.get(0, va.n_alloc);
      block_vmult(ctx, fs, newton_update.p, res); // refreshes the ghost entries of its argument
      axpby(ctx, va, 1.0, fs.rhs.p, -1.0, res);
      cur.true_res = nrm > 0.0 ? nrm2(ctx, va, res) / nrm : 0.0;
    }
    // constraints_used.distribute(newton_update)
    set_flagged(ctx, fs.n_dofs, fs.d_con.p, use_nonzero_constraints ? fs.d_nonzero_val.p : nullptr, newton_update.p);
    return {(unsigned)r.iterations, r.residual};
  }

  void InsIM::run_one_step(bool apply_nonzero_constraints, bool /*assemble_system*/)
  {
    io_before_step();
    time.increment();
    if (verbose && fs.rank == 0)
      std::printf("%s\nTime step = %u, at t = %e\n", std::string(96, '*').c_str(), time.get_timestep(), time.current());
    double current_residual = 1.0, initial_residual = 1.0, relative_residual = 1.0;
    unsigned int outer_iteration = 0;
    const VecSpace &n = fs.vs_all;
    copy(ctx, n, present_solution.p, evaluation_point.p);
    while (relative_residual > parameters.fluid_tolerance && current_residual > 1e-11)
      {
        if (outer_iteration >= parameters.fluid_max_iterations) throw std::runtime_error("Too many Newton iterations!");
        fill(ctx, n, 0.0, newton_update.p);
        cur = NewtonRecord{};
        const bool nz = apply_nonzero_constraints && outer_iteration == 0;
        assemble(nz);
        const auto state = solve(nz);
        current_residual = nrm2(ctx, n, fs.rhs.p);
        axpy(ctx, n, 1.0, newton_update.p, evaluation_point.p);
        fs.halo_update(ctx, evaluation_point.p); // evaluation_point = tmp (ghosted), mpi_insim.cpp:444-448
        if (outer_iteration == 0) initial_residual = current_residual;
        relative_residual = current_residual / initial_residual;
        cur.timestep = time.get_timestep();
        cur.iteration = outer_iteration;
        cur.abs_res = current_residual;
        cur.rel_res = relative_residual;
        cur.gmres_its = (int)state.first;
        cur.gmres_res = state.second;
        history.push_back(cur);
        if (verbose && fs.rank == 0)
          std::printf(" ITR = %-2u ABS_RES = %e REL_RES = %e GMRES_ITR = %-3u GMRES_RES = %e  [cg_mp %d cg_sm %d a_inv %d / %d]\n",
                      outer_iteration, current_residual, relative_residual, state.first, state.second, cur.cg_mp_its,
                      cur.cg_sm_its, cur.a_inv_its, cur.precond_applies);
        outer_iteration++;
      }
    // solution_increment = present - evaluation_point ; present = evaluation_point
    lin3(ctx, n, solution_increment.p, present_solution.p, -1.0, evaluation_point.p, 0.0, evaluation_point.p);
    copy(ctx, n, evaluation_point.p, present_solution.p);
    update_stress(); // mpi_insim.cpp:475
    io_after_step();
    // mpi_insim.cpp:485-489: a "Fluid" run refines with the Kelly estimator at the refinement interval (FluidSolver::refine_mesh,
    // mpi_fluid_solver.cpp:417-488). Not built: it leaves hanging nodes on the FE_Q(2) velocity space, which the condensation of
    // hanging.h does not cover (every reference .prm keeps the interval beyond the end time; the reference's own preconditioner
    // carries a FIXME about this path, :90-98). Said once instead of silently computing on the unrefined mesh.
    if (parameters.simulation_type == "Fluid" && time.time_to_refine() && time.end() - time.current() > 1e-12 && !refine_warned)
      {
        refine_warned = true;
        if (fs.rank == 0)
          std::fprintf(stderr, "openifem_b200: InsIM reached its refinement interval; FluidSolver::refine_mesh (Kelly estimator, FE_Q(2) "
                               "hanging nodes) is not built - the run continues on the current mesh\n");
      }
  }

  void InsIM::update_stress()
  {
    ScopedTimer t(ctx, timer_ms["Update stress"]);
    if (fs.n_ranks > 1) fs.halo_update(ctx, present_solution.p);
    update_nodal_stress(ctx, fs, present_solution.p, parameters.viscosity, stress.p);
  }

  void InsIM::run()
  {
    const bool success_load = load_checkpoint(); // mpi_insim.cpp:497-498; false unless an output directory is set
    if (!dofs_ready)
      {
        triangulation.refine_global(parameters.global_refinements.empty() ? 0 : parameters.global_refinements[0]);
        setup_dofs();
        make_constraints();
        initialize_system();
      }
    // Deviation: the reference calls run_one_step(true) here even after a successful load (mpi_insim.cpp:511), which adds the
    // inhomogeneous boundary values to a solution that already carries them; SUPGFluidSolver::run (mpi_supg_solver.cpp:456-463)
    // skips it after a restart, and so do we for every fluid solver
    if (!success_load) run_one_step(true);
    while (time.end() - time.current() > 1e-12) run_one_step(false);
  }

  std::vector<double> InsIM::get_current_solution() { return present_solution.to_host(ctx.stream); }
} // namespace ifem
