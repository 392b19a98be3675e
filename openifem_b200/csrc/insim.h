// Fluid::MPI::InsIM<dim> on the device (reference include/mpi_insim.h:42-194,
// source/mpi_insim.cpp): implicit incompressible Navier-Stokes, Newton iteration
// with grad-div stabilisation, FGMRES + block Schur preconditioner. The class keeps
// the reference's method names; dimension is a run-time property of the
// triangulation (the templated facade in include/openifem/ instantiates <2>/<3>).
#pragma once
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "fluid.h"
#include "inner32.h"
#include "krylov.h"
#include "output.h"

namespace ifem
{
  // Utils::Time (reference include/utilities.h:27-63, source/utilities.cpp:6-36)
  class Time
  {
  public:
    Time(double time_end, double delta_t, double output_interval, double refinement_interval, double save_interval)
      : timestep(0), time_current(0.0), delta_t(delta_t), time_end(time_end), output_interval(output_interval),
        refinement_interval(refinement_interval), save_interval(save_interval)
    {
    }
    double current() const { return time_current; }
    double end() const { return time_end; }
    double get_delta_t() const { return delta_t; }
    unsigned int get_timestep() const { return timestep; }
    bool time_to_output() const { return due(output_interval); }
    bool time_to_refine() const { return due(refinement_interval); }
    bool time_to_save() const { return due(save_interval); }
    void increment() { time_current += delta_t; ++timestep; }
    void decrement() { time_current -= delta_t; --timestep; }
    void set_delta_t(double d) { delta_t = d; }

  private:
    bool due(double interval) const
    {
      const auto delta = static_cast<unsigned int>(interval / delta_t);
      return delta != 0 && timestep >= delta && timestep % delta == 0;
    }
    unsigned int timestep;
    double time_current, delta_t;
    const double time_end, output_interval, refinement_interval, save_interval;
  };

  struct NewtonRecord
  {
    unsigned int timestep, iteration;
    double abs_res, rel_res;
    int gmres_its;
    double gmres_res;
    int cg_mp_its, cg_sm_its, a_inv_its, precond_applies;
    double true_res; // |b - A x| / |b| recomputed with the fp64 operator after the solve
  };

  // Tolerances of the linear solvers; defaults = Fluid::MPI::InsIM
  // (mpi_insim.cpp:73-109, 379-380). serial() = Fluid::InsIM (insim.cpp:353-358).
  struct InsSolverControl
  {
    double fgmres_rel = 1e-4, fgmres_floor = 1e-12;
    bool fgmres_floor_is_max = true; // max(floor, rel*|rhs|)
    double cg_mp_rel = 1e-6, cg_sm_rel = 1e-3, cg_floor = 1e-10;
    // A~^-1: the reference factorises with MUMPS; here an inner BiCGStab on A_uu with
    // the node-block Jacobi preconditioner, run to a_inv_rel * |src| (hard part 2)
    double a_inv_rel = 1e-3;
    int a_inv_max_it = 2000;
    // 0: BiCGStab + node-block Jacobi (InsIM); 1: plain CG to max(a_inv_floor, a_inv_rel |src|), the "CG for A" of
    // InsIMEX (mpi_insimex.cpp:118-131; A_uu is symmetric positive definite there)
    int a_inv_solver = 0;
    double a_inv_floor = 0.0;
    // 0: fp64 BiCGStab on the BCSR matrix; 1: same, A_uu streamed as fp32; 2: fp32 BiCGStab on the sliced copy of
    // A_uu (inner32.h); 3: as 2 with the matrix values of the copy stored as row-scaled fp16. Legal because FGMRES
    // is flexible; operator, residuals and Krylov basis stay fp64
    int a_inv_fp32 = 0;
    // "CG for Sm": 0 fp64 CG on the CSR matrix (reference arithmetic); 1 fp32 CG on a SELL-32 copy of S_m, driven from
    // device-resident scalars; 2 as 1 with row-scaled fp16 matrix values
    int cg_sm_fp32 = 0;
    int basis_size = 30;
    // SUPG solvers: ILU(0) factors for P_vv and B2pp as in the reference (1), Jacobi factors (0), or by problem size (-1)
    int supg_ilu = -1;
    static InsSolverControl serial()
    {
      InsSolverControl c;
      c.fgmres_rel = 1e-8;
      c.fgmres_floor = 1e-10;
      c.cg_sm_rel = 1e-6;
      return c;
    }
  };

  class InsIM
  {
  public:
    InsIM(Context &ctx, Triangulation &tria, const Parameters::AllParameters &params, bool taylor_hood_only = true);
    virtual ~InsIM() = default;

    virtual void run();
    virtual void run_one_step(bool apply_nonzero_constraints, bool assemble_system = true);
    // BlockVector of n_u + n_p doubles, copied to the host
    std::vector<double> get_current_solution();
    void add_hard_coded_boundary_condition(int id, std::function<double(const double *, unsigned int, double)> f);

    virtual void setup_dofs();
    void make_constraints();
    // push the host constraint flags / values (fs.con, fs.nonzero_val) to the device after a merge
    void upload_constraints();
    virtual void initialize_system();
    // second half of refine_mesh (source/mpi_fluid_solver.cpp:466-487, source/mpi_fsi.cpp:1090-1110) after the triangulation was
    // coarsened / refined: new dofs, constraints and system, present_solution interpolated onto the new mesh through the
    // vertex transfer plan (FE_Q(1) velocity and pressure, one rank); everything else starts as initialize_system() leaves it
    virtual void after_mesh_change(const Triangulation::TransferPlan &plan, const std::vector<double> &old_vertices);
    virtual void assemble(bool use_nonzero_constraints);
    virtual std::pair<unsigned int, double> solve(bool use_nonzero_constraints);
    // FluidSolver::update_stress (source/mpi_fluid_solver.cpp:716-811): nodal viscous stress 2 mu sym grad v from
    // present_solution, quadrature values projected to the dofs of FE_Q(pu) and averaged over the cells
    void update_stress();
    DevBuf<double> stress; // [dim*dim][n_velocity_nodes]

    // Result files and checkpoints (solver_io.cu; formats in output.h). Off until a directory is set; then run() /
    // run_one_step() write fluid_NNNNNN.{pvtu,procRRRR.vtu} + fluid.pvd at the output interval, NNNNNN.fluid_checkpoint at
    // the save interval, and run() restarts from the latest checkpoint found (mpi_fluid_solver.cpp:491-713)
    void set_output_directory(const std::string &dir);
    void output_results(unsigned int output_index);
    void save_checkpoint(int output_index);
    bool load_checkpoint();
    std::string output_directory;
    std::unique_ptr<io::PVDWriter> pvd_writer;

    Context &ctx;
    Triangulation &triangulation;
    Parameters::AllParameters parameters;
    FluidSpace fs;
    Time time;
    InsSolverControl control;
    bool verbose = false;
    DevBuf<double> present_solution, evaluation_point, solution_increment, newton_update, fsi_acceleration;
    std::vector<NewtonRecord> history;
    std::map<int, std::function<double(const double *, unsigned int, double)>> hard_coded;
    // clock of the hard-coded boundary functions (Function::advance_time): InsIM::run never advances it, SUPGFluidSolver::run
    // advances it by dt before every make_constraints() (mpi_supg_solver.cpp:438-444, 470-478)
    double bc_time = 0.0;
    bool bc_clock_started = false; // SUPGFluidSolver::run advanced the boundary functions' clock before the first step
    double base_bc_time = 0.0; // clock of the boundary functions the cached constraint lines were made at
    // per-section device time, keyed by the reference's TimerOutput section names
    std::map<std::string, double> timer_ms;
    bool dofs_ready = false;
    bool refine_warned = false;
    // called at the end of make_constraints(): an attached turbulence model re-makes its own lines (mpi_fluid_solver.cpp:276-279)
    std::function<void()> after_make_constraints;
    // refine_mesh: further scalar nodal fields (pressure-node numbering) carried to the new mesh with present_solution - nu~ of an
    // attached turbulence model (pre_refine_mesh / post_refine_mesh, source/mpi_spalart_allmaras.cpp:594-617); on_mesh_change
    // runs before the new spaces are set up
    virtual std::vector<DevBuf<double> *> transferred_scalar_fields() { return {}; }
    std::function<void()> on_mesh_change;

  protected:
    void io_before_step(); // output of step 0
    void io_after_step();  // output / checkpoint when due
    void precondition(const double *src, double *dst);
    DevBuf<double> d_binv, d_con_vals, d_tmp_p, d_utmp, d_utmp2;
    int64_t n_dofs_global = 0, n_p_global = 0;
    VecPool pool_fgmres, pool_cg, pool_ainv;

  public:
    InnerSolver32 inner32;
    InnerCG32 inner_sm;
    DeviceCG64 cg_mp_dev; // "CG for Mp"
    // "CG for Sm": x = S_m^-1 b (pressure vectors on the device) in the given cg_sm_fp32 mode; b_norm = |b| over all ranks
    SolveResult solve_mass_schur(int mode, const double *b, double b_norm, double *x, double tol_abs, int max_it);

  protected:
    NewtonRecord cur{};
    bool sm_copy_valid = false; // inner_sm holds the current S_m

  public:
    long long sm_fallbacks = 0; // fp32 "CG for Sm" applications redone in fp64 (see solve_mass_schur)

  protected:
  };
} // namespace ifem
