// Fluid::MPI::InsIMEX<dim> on the device (reference include/mpi_insimex.h, source/mpi_insimex.cpp): the
// implicit-explicit twin of InsIM. The unknown of the single linear solve per time step is the increment of the
// solution; viscous, grad-div, pressure-coupling and mass/dt terms are implicit, convection is explicit. The matrix does
// not depend on the solution: it is assembled in time steps 1 (nonzero constraints) and 2 (zero constraints) only
// (:503-509), later steps run the cell kernel in its right-hand-side-only mode. Shares FluidSpace, the cell kernel
// (explicit_convection / rhs_only switches of ins_assemble), FGMRES and the Schur-complement preconditioner with InsIM;
// differs in "CG for A" (:118-131) standing in for the direct solve, in the FGMRES tolerance (:370-371) and in the
// time loop (:449-480).
#pragma once
#include "insim.h"

namespace ifem
{
  class InsIMEX : public InsIM
  {
  public:
    InsIMEX(Context &ctx, Triangulation &tria, const Parameters::AllParameters &params);

    // assemble(use_nonzero_constraints, assemble_system) (mpi_insimex.cpp:150-355)
    void assemble(bool use_nonzero_constraints, bool assemble_system);
    void assemble(bool use_nonzero_constraints) override { assemble(use_nonzero_constraints, true); }
    // solve(use_nonzero_constraints, assemble_system) (:357-386): the solution lands in newton_update
    // (= solution_time_increment of the reference)
    std::pair<unsigned int, double> solve(bool use_nonzero_constraints, bool assemble_system);
    std::pair<unsigned int, double> solve(bool use_nonzero_constraints) override { return solve(use_nonzero_constraints, true); }
    void run_one_step(bool apply_nonzero_constraints, bool assemble_system = true) override;
    void run() override;
  };
} // namespace ifem
