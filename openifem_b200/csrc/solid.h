// Solid::MPI::HyperElasticity<dim> and Solid::MPI::[Shared]LinearElasticity<dim> on the device (reference include/mpi_hyper_elasticity.h,
// source/mpi_hyper_elasticity.cpp; base class source/mpi_solid_solver.cpp): total-Lagrangian
// NeoHookean solid, Newmark-beta time integration, Newton iteration, CG linear solves.
//
// Per-quadrature-point history (Internal::PointHistory: F_inv, tau, Jc, det F) lives in flat device
// arrays indexed [cell][q] instead of the reference's shared_ptr<PointHistory> per point
// (mpi_hyper_elasticity.h:59-66); `update_qph` is one kernel, one thread per quadrature point.
// The solid mesh is small and (as in MPI::FSI's SharedSolidSolver) replicated on every rank.
#pragma once
#include <map>

#include "insim.h" // Time
#include "krylov.h"
#include "mesh.h"
#include "output.h"
#include <memory>
#include "parameters.h"

namespace ifem
{
  struct SolidSpace
  {
    int dim = 0, degree = 1, npc = 0, nq = 0, nv = 0, nsym = 0;
    int n_cells = 0;
    NodeTable nt;
    int64_t n_dofs = 0;
    Pattern P;
    std::vector<int> colour_order, colour_offsets;
    std::vector<unsigned char> con;
    DevBuf<int> d_cell_nodes, d_colour_order, d_con_idx;
    DevBuf<unsigned char> d_slots, d_con;
    int n_con = 0;
    DevBuf<double> d_N;    // [nq][npc]
    DevBuf<double> d_G;    // [n_cells][nq][npc][dim] physical gradients on the reference configuration
    DevBuf<double> d_JxW;  // [n_cells][nq]
    DevBuf<double> d_node_x; // [n_nodes][dim]
    // PointHistory
    DevBuf<double> d_Finv; // [n_cells][nq][dim*dim]
    DevBuf<double> d_tau;  // [n_cells][nq][dim*dim]
    DevBuf<double> d_Jc;   // [n_cells][nq][nsym*nsym]  (Voigt pairs: diagonal first, then (0,1),(0,2),(1,2))
    DevBuf<double> d_detF; // [n_cells][nq]
    // Neumann faces
    DevBuf<int> d_nface;       // (cell, face)
    DevBuf<double> d_nface_val; // dim values per face (traction) or 1 (pressure)
    DevBuf<double> d_face_tables;
    int n_nfaces = 0, nqf = 0, neumann_is_pressure = 0;
    bool neumann_skips_dirichlet_faces = true; // mpi_hyper_elasticity.cpp:452-456; the linear solvers integrate them
    Bcsr K, M;
    DevBuf<double> rhs;

    void setup(Context &ctx, const Triangulation &tria, const Parameters::AllParameters &prm);
  };

  // What Solid::MPI::SolidSolver / SharedSolidSolver hold for every material (include/mpi_solid_solver.h:75-160,
  // include/mpi_shared_solid_solver.h:91-185): the space, the Newmark vectors, the linear solver and what MPI::FSI
  // reads and writes in the solid.
  class SolidSolver
  {
  public:
    SolidSolver(Context &ctx, Triangulation &tria, const Parameters::AllParameters &params);
    virtual ~SolidSolver() = default;
    void run();
    virtual void run_one_step(bool first_step) = 0;
    virtual void assemble_system(bool initial_step) = 0;
    virtual void update_strain_and_stress() = 0;
    std::vector<double> get_current_solution();
    void setup_dofs();
    virtual void initialize_system();
    std::pair<unsigned int, double> solve(Bcsr &A, double *x, const double *b);

    Context &ctx;
    Triangulation &triangulation;
    Parameters::AllParameters parameters;
    SolidSpace ss;
    Time time;
    bool verbose = false, dofs_ready = false;
    DevBuf<double> stress, strain; // [dim*dim][n_nodes] nodal stress / strain (hyperelastic: Cauchy stress, deformation gradient)
    // what MPI::FSI writes into the solid (include/mpi_shared_solid_solver.h: fsi_stress_rows, fluid_velocity,
    // fluid_pressure; used by mpi_fsi.cpp:793-806 and the FSI traction term mpi_shared_hyper_elasticity.cpp:495-554)
    DevBuf<double> fsi_stress_rows; // [dim][n_dofs]: row d1 of the fluid stress at every vertex
    DevBuf<double> fluid_velocity;  // [n_dofs]
    DevBuf<double> fluid_pressure;  // [n_nodes]
    DevBuf<double> current_displacement, current_velocity, current_acceleration, previous_displacement, previous_velocity,
      previous_acceleration;
    struct Record
    {
      unsigned int timestep, iteration;
      double res_F, res_U;
      int cg_its;
    };
    std::vector<Record> history;
    // Result files and checkpoints (solver_io.cu; formats in output.h): off until a directory is set
    // (mpi_shared_solid_solver.cpp:237-337, 452-571)
    void set_output_directory(const std::string &dir);
    void output_results(unsigned int output_index);
    void save_checkpoint(int output_index);
    bool load_checkpoint();
    // what run() does instead of the first step after a successful load: rebuild what later steps reuse
    virtual void after_restart() { assemble_system(true); }
    std::string output_directory;
    std::unique_ptr<io::PVDWriter> pvd_writer;
    std::map<std::string, double> timer_ms;

  protected:
    void io_before_step();
    void io_after_step();
    double get_error(const double *v);
    // traction / pressure faces, or (FSI) the fluid traction on the deformed faces, added to ss.rhs
    void neumann_rhs();
    DevBuf<double> d_binv, d_tmp, d_pred, d_update, d_qpt_to_dof, d_count;
    VecPool pool;
  };

  class HyperElasticity : public SolidSolver
  {
  public:
    // shared: 1 = Solid::MPI::SharedHyperElasticity (the replicated twin MPI::FSI takes: extra Newton stop on a vanishing update,
    // nodal strain / stress every step), 0 = Solid::MPI::HyperElasticity, -1 = pick by `Simulation type` (FSI -> the twin)
    HyperElasticity(Context &ctx, Triangulation &tria, const Parameters::AllParameters &params, int shared = -1);
    bool shared_twin = false;
    void run_one_step(bool first_step) override;
    void initialize_system() override;
    void update_qph(const double *u_dev);
    void after_restart() override
    {
      update_qph(current_displacement.p); // the point history follows the loaded displacement
      assemble_system(true);
    }
    void assemble_system(bool initial_step) override;
    // SharedHyperElasticity::update_strain_and_stress (source/mpi_shared_hyper_elasticity.cpp:599-714): Cauchy stress
    // tau / J and deformation gradient F at the quadrature points, projected to the nodes and averaged
    void update_strain_and_stress() override;

  private:
    DevBuf<double> d_cell_mat; // [n_cells][2]: material parameters of every cell from its material id
  };

  // Solid::MPI::LinearElasticity<dim> (source/mpi_linear_elasticity.cpp; shared = false) and
  // Solid::MPI::SharedLinearElasticity<dim> (source/mpi_shared_linear_elasticity.cpp; shared = true, the twin MPI::FSI
  // drives): small-strain elasticity with the material of source/linear_elastic_material.cpp, Newmark-beta in
  // acceleration form - the matrices are assembled once, every step is two products and one CG solve.
  class LinearElasticity : public SolidSolver
  {
  public:
    LinearElasticity(Context &ctx, Triangulation &tria, const Parameters::AllParameters &params, bool shared);
    void run_one_step(bool first_step) override;
    void initialize_system() override;
    void assemble_system(bool is_initial) override;
    // mpi_shared_linear_elasticity.cpp:401-531: sym grad u and C : sym grad u at the quadrature points -> nodes, averaged
    void update_strain_and_stress() override;

    void after_restart() override
    {
      assemble_system(true);
      if (!shared) assemble_system(false); // system = M + beta dt^2 K and K (mpi_linear_elasticity.cpp:207-214)
    }

    const bool shared;
    // system_matrix = ss.K, mass_matrix = ss.M (shared twin only)
    Bcsr stiffness_matrix, damping_matrix;

  private:
    double lambda = 0, mu = 0, eta = 0;
    DevBuf<double> d_tmp2, d_tmp3;
  };
} // namespace ifem
