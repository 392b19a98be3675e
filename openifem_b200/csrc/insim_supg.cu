// Fluid::MPI::SUPGInsIM<dim>::assemble on the device (reference source/mpi_insim_supg.cpp:15-328); kernel bodies in
// insim_supg.cuh, class in scnsim.h.
#include "insim_supg.cuh"
#include "scnsim.h"

#include <chrono>

namespace ifem
{
  namespace
  {
    // CTA = CPB cells x NU*NU threads (64 threads: 4 cells in 2-D, 1 cell in 3-D); cells of one colour per launch, so the
    // read-modify-write scatter needs no atomics and is bitwise reproducible
    template <int DIM>
    __global__ void __launch_bounds__(64) supg_ins_assemble_kernel(const SupgArgs A)
    {
      constexpr int NU = 1 << DIM, NQ = NU, PAIRS = NU * NU, CPB = 64 / PAIRS, DPC = NU * (DIM + 1);
      __shared__ SupgQPoint<DIM> sq[CPB][NQ];
      __shared__ double lrhs[CPB][DPC], ldiag[CPB][DPC];
      const int cl = threadIdx.x / PAIRS, pr = threadIdx.x % PAIRS;
      const int li = blockIdx.x * CPB + cl;
      const bool active = li < A.n_list;
      const int cell = active ? A.cell_list[li] : 0;
      if (active && pr < NQ) supg_fill_qpoint<DIM>(A, cell, pr, sq[cl][pr]);
      if (active && pr < DPC)
        {
          lrhs[cl][pr] = 0.0;
          ldiag[cl][pr] = 0.0;
        }
      __syncthreads();
      if (active) supg_pair_body<DIM>(A, cell, pr, sq[cl], lrhs[cl], ldiag[cl]);
      __syncthreads();
      if (active && pr < DPC) supg_rhs_body<DIM>(A, cell, pr, lrhs[cl], ldiag[cl]);
    }

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

  void SUPGInsIM::assemble(bool use_nonzero_constraints)
  {
    if (fs.pu != 1 || fs.pp != 1) throw std::runtime_error("SUPGInsIM: equal-order Q1/Q1 elements only (what the reference's cases use)");
    SectionTimer t(ctx, timer_ms["Assemble system"]);
    if (fs.n_ranks > 1)
      {
        fs.halo_update(ctx, evaluation_point.p);
        fs.halo_update(ctx, present_solution.p);
      }
    cudaStream_t s = ctx.stream;
    fs.A_uu.zero(s);
    fs.A_up.zero(s);
    fs.A_pu.zero(s);
    fs.A_pp.zero(s);
    fs.rhs.zero(s);
    SupgArgs a{};
    a.cell_un = fs.d_cell_un.p;
    a.cell_pn = fs.d_cell_pn.p;
    a.cell_x = fs.d_cell_x.p;
    a.tables = fs.d_tables.p;
    a.slots = fs.d_slots.p;
    a.con = fs.d_con.p;
    a.eval_pt = evaluation_point.p;
    a.present = present_solution.p;
    a.body_force = d_body_force.n ? d_body_force.p : nullptr;
    a.inhom = use_nonzero_constraints ? fs.d_nonzero_val.p : nullptr;
    a.n_u = fs.n_u;
    a.n_owned_u = fs.n_owned_unodes;
    a.n_owned_p = fs.n_owned_pnodes;
    // the first dofs_per_cell / dofs_per_vertex system shape functions (:131-137): vertex v carries dim velocity
    // components then the pressure, all with the Q1 shape of that vertex
    const int dim = fs.dim;
    a.n_h = 1 << dim;
    for (int k = 0; k < a.n_h; ++k) a.h_node[k] = k / (dim + 1);
    a.mu = parameters.viscosity;
    a.rho = parameters.fluid_rho;
    a.dt = time.get_delta_t();
    for (int d = 0; d < 3; ++d) a.grav[d] = d < (int)parameters.gravity.size() ? parameters.gravity[d] : 0.0;
    a.uu_rp = fs.A_uu.rowptr.p;
    a.up_rp = fs.A_up.rowptr.p;
    a.pu_rp = fs.A_pu.rowptr.p;
    a.pp_rp = fs.A_pp.rowptr.p;
    a.uu = fs.A_uu.val.p;
    a.up = fs.A_up.val.p;
    a.pu = fs.A_pu.val.p;
    a.pp = fs.A_pp.val.p;
    a.rhs = fs.rhs.p;
    const int n_colours = (int)fs.colour_offsets.size() - 1;
    for (int k = 0; k < n_colours; ++k)
      {
        a.n_list = fs.colour_offsets[k + 1] - fs.colour_offsets[k];
        a.cell_list = fs.d_colour_order.p + fs.colour_offsets[k];
        if (!a.n_list) continue;
        if (dim == 2)
          supg_ins_assemble_kernel<2><<<(a.n_list + 3) / 4, 64, 0, s>>>(a);
        else
          supg_ins_assemble_kernel<3><<<a.n_list, 64, 0, s>>>(a);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
    neumann_faces(ctx, fs); // the pressure face term of :292-321 is InsIM's
    fs.hanging.condense(ctx, fs, a.inhom); // hanging-node lines of a locally refined mesh
  }
} // namespace ifem
