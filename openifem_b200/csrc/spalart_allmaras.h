// Fluid::MPI::SpalartAllmaras<dim> on the device (reference include/mpi_spalart_allmaras.h, source/mpi_spalart_allmaras.cpp, base
// class include/mpi_turbulence_model.h, source/mpi_turbulence_model.cpp): one-equation transport of the working viscosity nu~
// on the scalar space FE_Q(velocity degree) of the fluid solver. Attached with FluidSolver::attach_turbulence_model
// ("Spalart-Allmaras", source/mpi_fluid_solver.cpp:53-63); advanced before every fluid step (source/mpi_supg_solver.cpp:456-468,
// source/mpi_fsi.cpp:1207-1210); SCnsIM::assemble reads the eddy viscosity mu_t = f_v1 nu~ rho (source/mpi_scnsim.cpp:198-216).
//
// The model shares everything mesh-related with its fluid solver (the FluidSolverExtractor of the reference): cell lists and
// colouring, the Q1 tables, the pattern and the scatter slots of the pressure block (the scalar space of an equal-order Q1/Q1
// solver has the nodes of the pressure space), the halo of pressure vectors, the hanging-node fold plan of A_pp. Equal-order
// Q1/Q1 fluid solvers only (what every reference SCnsIM case uses).
//
// Not built: the wall function of immersed (FSI) walls - update_moving_wall_distance (:17-130) and the y+ lines of
// update_boundary_condition (:194-216), which need the shear velocities FSI::find_solid_bc samples at image points
// (source/mpi_fsi.cpp:784-847); get_shear_velocity itself (:227-293) is here. The lines of the cells inside the solid are built.
//
// One statement of the reference cannot be kept literally: `r` of the destruction term is read from a lambda that evaluates
// std::min({nu~ / (S~ kappa^2 d^2), 10.0}) and drops the result (:757-770) - indeterminate whenever |S~| > 1e-8. The value that
// expression computes, r = min(nu~ / (S~ kappa^2 d^2), 10) - the published model - is used here and in the oracle.
#pragma once
#include <utility>
#include <vector>

#include "ilu0.h"
#include "krylov.h"
#include "linalg.h"

namespace ifem
{
  class SCnsIM;

  class SpalartAllmaras
  {
  public:
    struct Record
    {
      unsigned int iteration;
      double abs_res, rel_res;
      int gmres_its;
      double gmres_res;
    };

    SpalartAllmaras(Context &ctx, SCnsIM &fluid);

    void make_constraints();                         // :352-412
    void initialize_system();                        // :555-581 (+ TurbulenceModel::initialize_system, setup_cell_property)
    void setup_cell_property();                      // :415-552 fixed wall distance
    void update_boundary_condition(bool first_step); // :133-224 (cells inside the immersed solid)
    void assemble(bool use_nonzero_constraints);     // :620-832
    std::pair<unsigned int, double> solve(bool use_nonzero_constraints); // :835-861
    void run_one_step(bool apply_nonzero_constraints);                   // :296-349
    void update_eddy_viscosity();                                        // :864-889
    double get_shear_velocity(double vel, double init_guess) const;      // :227-293
    // refine_mesh: the cached lines and the ILU analysis belong to the old mesh. nu~ itself travels with the fluid's solution
    // (InsIM::after_mesh_change = pre_refine_mesh / post_refine_mesh, :594-617); the eddy viscosity restarts from zero until the
    // next model step, as TurbulenceModel::initialize_system leaves it (source/mpi_turbulence_model.cpp:123-125)
    void mesh_changed()
    {
      constraints_made = false;
      ilu = Ilu0();
    }

    Context &ctx;
    SCnsIM &fluid;
    DevBuf<double> present_solution, evaluation_point, newton_update, eddy_viscosity, fixed_wall_distance, system_rhs;
    Bcsr system_matrix;
    DevBuf<unsigned char> d_con, d_base_con;
    DevBuf<double> d_nonzero_val, d_base_val;
    std::vector<Record> history;
    bool verbose = false;
    bool ready = false;

  private:
    bool use_ilu() const;
    DevBuf<double> d_diag_inv;
    Ilu0 ilu;
    VecPool pool;
    bool constraints_made = false;
    size_t base_nodes = 0;
  };
} // namespace ifem
