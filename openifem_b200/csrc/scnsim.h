// Fluid::MPI::SCnsIM<dim> on the device (reference include/mpi_scnsim.h, source/mpi_scnsim.cpp:15-568) on top
// of Fluid::MPI::SUPGFluidSolver (include/mpi_supg_solver.h, source/mpi_supg_solver.cpp): slightly
// compressible Navier-Stokes, SUPG / PSPG / LSIC stabilisation, PML attenuation, artificial-fluid terms for
// the immersed solid. Equal-order Q1/Q1 elements (what every reference SCnsIM test and BASELINE configs 4, 5
// use) have their own kernel; other pairs (Q2/Q1, Q2/Q2) go through the degree-generic kernel of scnsim_generic.cu. Shares the FluidSpace / Newton / FGMRES machinery of InsIM; differs in the cell integrand, in
// update_stress() after every step and in the block preconditioner (BlockIncompSchurPreconditioner).
#pragma once
#include "ilu0.h"
#include "insim.h"
#include "spalart_allmaras.h"

namespace ifem
{
  // SCnsIM::assemble for element pairs other than Q1/Q1 (scnsim_generic.cu): fills A_uu / A_up / A_pu / A_pp and rhs of the space
  struct ScnsGenericInput
  {
    const double *eval_pt, *present, *fsi_acc, *stress, *fsi_stress, *sigma_pml, *body_force;
    double mu, rho_f, rho_s, dt, grav[3];
  };
  void scns_assemble_generic(Context &ctx, FluidSpace &fs, const ScnsGenericInput &in, bool use_nonzero_constraints);

  class SCnsIM : public InsIM
  {
  public:
    SCnsIM(Context &ctx, Triangulation &tria, const Parameters::AllParameters &params);

    // set_body_force / set_sigma_pml_field / set_initial_condition (include/mpi_fluid_solver.h:120-143):
    // f(point, component) evaluated on the host at setup, stored per quadrature point / per dof
    void set_body_force(std::function<double(const double *, unsigned int)> f) { body_force = std::move(f); }
    void set_sigma_pml_field(std::function<double(const double *, unsigned int)> f) { sigma_pml_field = std::move(f); }
    void set_initial_condition(std::function<double(const double *, unsigned int)> f) { initial_condition = std::move(f); }

    void setup_dofs() override;
    void initialize_system() override;
    void assemble(bool use_nonzero_constraints) override;
    std::pair<unsigned int, double> solve(bool use_nonzero_constraints) override;
    void run_one_step(bool apply_nonzero_constraints, bool assemble_system = true) override;
    // SUPGFluidSolver::run (mpi_supg_solver.cpp:427-486): time-dependent hard-coded boundary values
    void run() override;
    DevBuf<double> fsi_stress; // [dim(dim+1)/2][n_unodes]
    int tpp_its = 0;
    // FluidSolver::attach_turbulence_model (source/mpi_fluid_solver.cpp:53-63; TurbulenceModelFactory::create accepts
    // "Spalart-Allmaras" only, source/mpi_turbulence_model.cpp:11-26). Before or after setup.
    void attach_turbulence_model(const std::string &model_name);
    std::unique_ptr<SpalartAllmaras> turbulence_model;
    std::vector<DevBuf<double> *> transferred_scalar_fields() override
    {
      if (turbulence_model && turbulence_model->ready) return {&turbulence_model->present_solution};
      return {};
    }

  protected:
    void precondition_supg(const double *src, double *dst);
    std::function<double(const double *, unsigned int)> body_force, sigma_pml_field, initial_condition;
    DevBuf<double> d_sigma_pml, d_body_force; // [n_cells][nq], [n_cells][nq][dim] (empty when unset)
    DevBuf<double> d_rowsum_inv, d_b2pp_diag_inv, d_pt1, d_pt2, d_ut1, d_ut2;
    VecPool pool_tpp;
    // ILU(0) factors of A_vv and B2pp (ilu0.h) - the reference's Euclid factors; used on one rank up to kIluMaxRows scalar rows
    // (control.supg_ilu: -1 by size, 0 never = Jacobi factors, 1 always)
    Ilu0 ilu_vv, ilu_b2;
    bool use_ilu() const;
  };

  // Fluid::MPI::SUPGInsIM<dim> (reference include/mpi_insim_supg.h, source/mpi_insim_supg.cpp): incompressible
  // Navier-Stokes with SUPG / PSPG / LSIC stabilisation on equal-order Q1/Q1 elements. Everything around the cell loop is
  // SUPGFluidSolver (Newton loop, FGMRES + block preconditioner, update_stress, time loop) and shared with SCnsIM.
  class SUPGInsIM : public SCnsIM
  {
  public:
    SUPGInsIM(Context &ctx, Triangulation &tria, const Parameters::AllParameters &params) : SCnsIM(ctx, tria, params) {}
    void assemble(bool use_nonzero_constraints) override;
  };
} // namespace ifem
