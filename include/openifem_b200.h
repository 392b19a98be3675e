/* openifem_b200 C ABI - the drop-in boundary of the B200-native hot path.
 *
 * OpenIFEM has no FFI of its own: its boundary is the C++ class surface consumed by
 * the test drivers (reference include/mpi_fluid_solver.h:99-183, include/mpi_insim.h:42-87,
 * include/parameters.h:191). This header is the extern "C" layer underneath our
 * same-named C++ classes (include/openifem/*.h); every entry point cites the
 * reference member it stands for. Plain pointers and sizes only; handles are
 * opaque; all functions return 0 on success and a non-zero code on failure with the
 * message available from ifem_last_error() (the reference throws dealii exceptions;
 * the C++ facade converts the code back into a throw).
 *
 * There is no CPU fallback: every compute entry point fails with IFEM_ERR_NO_DEVICE
 * when no CUDA device is present.
 */
#ifndef OPENIFEM_B200_H
#define OPENIFEM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IFEM_OK 0
#define IFEM_ERR 1
#define IFEM_ERR_NO_DEVICE 2

typedef struct ifem_tria ifem_tria;     /* dealii::Triangulation<dim> stand-in */
typedef struct ifem_params ifem_params; /* Parameters::AllParameters */
typedef struct ifem_insim ifem_insim;   /* Fluid::MPI::InsIM<dim> */

const char *ifem_last_error(void);
int ifem_version(void);
/* bind this process (= one rank) to a CUDA device; called once before any solver is created */
int ifem_init(int device);
/* host threads (OpenMP) used by the setup code: mesh tables, sparsity patterns, partitioning */
int ifem_set_host_threads(int n);
/* number of kernels launched by the library so far in this process */
int ifem_kernel_launches(int64_t *count);
/* measured FP64 FMA throughput of the device (register-only FMA chains, TFLOP/s): the compute-side denominator of the assembly
 * kernels' roofline (bench.py) */
int ifem_bench_fp64_peak(double *tflops);
/* collective self test of the peer-memory link between the ranks of one node (csrc/peer.h): shared buffers written by kernels of
 * the peers and `rounds` in-kernel all-reduces of known values; *mismatches = 0 when everything arrived, -1 when the link is
 * inactive (one rank, IPC unavailable, IFEM_PEER=0) */
int ifem_peer_selftest(int rounds, int64_t *mismatches);

/* ---- ranks: one process per GPU. Rank 0 creates the NCCL unique id, the launcher broadcasts the 128 bytes
 *      (torch.distributed / MPI / files) and every rank calls ifem_comm_init before creating solvers.
 *      Replaces MPI_COMM_WORLD of the reference (source/mpi_fluid_solver.cpp:37). ---- */
int ifem_comm_unique_id(unsigned char id[128]);
int ifem_comm_init(int rank, int size, const unsigned char id[128]);
int ifem_comm_finalize(void);

/* ---- host-side domain decomposition (what p4est + DoFHandler::distribute_dofs decide in the reference,
 *      include/mpi_fluid_solver.h:187, source/mpi_fluid_solver.cpp:140-152); no device needed ---- */
typedef struct ifem_partition ifem_partition;
int ifem_partition_create(const ifem_tria *t, int velocity_degree, int pressure_degree, int rank, int size, ifem_partition **out);
int ifem_partition_destroy(ifem_partition *p);
/* which: 0 velocity nodes, 1 pressure nodes */
/* n_layer1 = owned + layer-1 ghosts; n_messages = halo messages (one per ghost layer and neighbour) */
int ifem_partition_counts(const ifem_partition *p, int which, int *n_owned, int *n_layer1, int *n_local, int *n_messages, int *n_local_cells);
int ifem_partition_local_to_global(const ifem_partition *p, int which, int *global_ids);
int ifem_partition_neighbour(const ifem_partition *p, int which, int k, int *rank, int *n_send, int *recv_offset, int *recv_count);
int ifem_partition_send_list(const ifem_partition *p, int which, int k, int *local_ids);

/* ---- Triangulation: GridGenerator calls of the reference test drivers
 *      (tests/fluid_cavity/fluid_cavity.cpp:28-34, tests/fluid_pipe_mpi/fluid_pipe_mpi.cpp:37-45) ---- */
int ifem_tria_create(int dim, ifem_tria **out);
int ifem_tria_destroy(ifem_tria *t);
int ifem_tria_subdivided_hyper_rectangle(ifem_tria *t, const unsigned int *repetitions, const double *p1, const double *p2,
                                         int colorize);
int ifem_tria_hyper_cube(ifem_tria *t, double left, double right, int colorize);
int ifem_tria_refine_global(ifem_tria *t, int times);
/* GridTools::shift(offset, tria): offset [dim] added to every vertex */
int ifem_tria_shift(ifem_tria *t, const double *offset);
/* cell->set_refine_flag() on the cells with flags[cell] != 0 followed by execute_coarsening_and_refinement()
 * (tests/fsi_leaflet_mpi/fsi_leaflet_mpi.cpp:66-76, tests/fsi-wall-3D/fsi-wall-3D.cpp:47-53): the flagged cells are replaced by their
 * children, hanging vertices appear on the interface (one level of difference only). n = number of active cells. */
int ifem_tria_execute_refinement(ifem_tria *t, const unsigned char *flags, int64_t n);
/* set_refine_flag() / set_coarsen_flag() on the flagged cells, prepare_coarsening_and_refinement() and
 * execute_coarsening_and_refinement() as FSI::refine_mesh / FluidSolver::refine_mesh use them (source/mpi_fsi.cpp:1062-1088,
 * source/mpi_fluid_solver.cpp:417-488): one level up or down per call, a family is coarsened when all its children are flagged,
 * flags are adjusted to p4est's 2:1 balance over shared vertices. coarsen_flags may be NULL. The values of a Q1 field on the new
 * vertices in terms of the old ones (parallel::distributed::SolutionTransfer) are kept for ifem_tria_get_transfer_plan. */
int ifem_tria_execute_coarsening_and_refinement(ifem_tria *t, const unsigned char *refine_flags, const unsigned char *coarsen_flags, int64_t n);
/* CSR lists new vertex -> (old vertex, weight) of the last ifem_tria_execute_coarsening_and_refinement; any pointer may be NULL */
int ifem_tria_get_transfer_plan(const ifem_tria *t, int64_t *n_new_vertices, int64_t *n_entries, int64_t *ptr, int *old_vertex, double *weight);
/* cell->level() of every active cell */
int ifem_tria_get_levels(const ifem_tria *t, int *levels);
/* hanging vertices of the active mesh: vertex[k] carries the mean of its n_masters[k] (2 or 4) master vertices masters[4 k ..];
 * call with NULL arrays to get the count */
int ifem_tria_get_hanging(const ifem_tria *t, int64_t *n_hanging, int *vertex, int *n_masters, int *masters);
/* cell->set_material_id() for every active cell (1-based part numbers; the solid solvers pick the material parameters of a cell's
 * part, source/mpi_hyper_elasticity.cpp:226-228); n must equal the number of active cells */
int ifem_tria_set_material_ids(ifem_tria *t, const int *ids, int64_t n);
int ifem_tria_counts(const ifem_tria *t, int64_t *n_vertices, int64_t *n_cells, int64_t *n_boundary_faces);
/* Utils::GridCreator<dim>::flow_around_cylinder(tria) (source/utilities.cpp:343-574; dim of the handle): the mesh of
 * tests/fluid_cylinder_mpi*. In 2-D the cells around the hole carry their polar / transfinite charts, which
 * ifem_tria_refine_global honours like deal.II's manifolds; the 3-D mesh is the extruded one and refines flat. */
int ifem_tria_flow_around_cylinder(ifem_tria *t);
/* copy the mesh out: vertices [n_vertices][dim], cells [n_cells][2^dim] (lexicographic corners), boundary_faces
 * [n_boundary_faces][3] = (cell, face_no = 2*axis+side, boundary id); any pointer may be NULL */
int ifem_tria_get_mesh(const ifem_tria *t, double *vertices, int *cells, int *boundary_faces);
/* hand an existing mesh to the library - what a deal.II-side host does with its own Triangulation (the reference's solvers take
 * any parallel::distributed::Triangulation, include/mpi_fluid_solver.h:99, read with GridIn in e.g. tests/fsi_*): active cells
 * only, the arrays of ifem_tria_get_mesh; corners in deal.II's lexicographic order (GeometryInfo<dim>::vertex numbering), face_no =
 * deal.II's face number 2*axis+side, boundary ids as set on the faces; material_ids [n_cells] or NULL. Cells must be positively
 * oriented; hanging vertices (one level of difference) are found from the geometry. */
int ifem_tria_set_mesh(ifem_tria *t, int64_t n_vertices, const double *vertices, int64_t n_cells, const int *cells, int64_t n_boundary_faces,
                       const int *boundary_faces, const int *material_ids);

/* ---- Parameters::AllParameters(prm_file) (include/parameters.h:191, source/parameters.cpp:618-658) ---- */
int ifem_params_from_file(const char *prm_file, ifem_params **out);
int ifem_params_from_text(const char *prm_text, ifem_params **out);
int ifem_params_destroy(ifem_params *p);

/* ---- Fluid::MPI::InsIM<dim> (include/mpi_insim.h:42-87, source/mpi_insim.cpp) ---- */
typedef struct
{
  double fgmres_rel, fgmres_floor; /* SolverControl tol = max(floor, rel*|rhs|), mpi_insim.cpp:379-380 */
  double cg_mp_rel, cg_sm_rel, cg_floor; /* mpi_insim.cpp:73-109 */
  double a_inv_rel;   /* inner Krylov stand-in for the MUMPS LU of A~ (mpi_insim.cpp:124-127) */
  int a_inv_max_it;
  int basis_size;     /* SolverFGMRES max_basis_size (deal.II default 30) */
  int a_inv_fp32;     /* inner solve: 0 fp64; 1 fp64 BiCGStab, A_uu streamed as fp32; 2 fp32 BiCGStab on the sliced (SELL-32)
                         copy of A_uu; 3 as 2, matrix values of the copy stored as row-scaled fp16. Preconditioner only -
                         operator, residuals and FGMRES basis stay fp64 */
  int cg_sm_fp32;     /* "CG for Sm" (mpi_insim.cpp:88-109): 0 fp64 CG on the CSR matrix; 1 fp32 CG on a SELL-32 copy of S_m;
                         2 as 1 with row-scaled fp16 matrix values. Preconditioner only, as above */
  int supg_ilu;       /* SCnsIM / SUPGInsIM block preconditioner (mpi_supg_solver.cpp:35-192): 1 ILU(0) factors of A_vv and B2pp like the
                         reference's Euclid ones (one rank), 0 Jacobi factors, -1 ILU(0) up to 60 000 velocity rows on one rank */
} ifem_ins_control;

typedef struct
{
  unsigned int timestep, iteration;
  double abs_res, rel_res; /* the ITR / ABS_RES / REL_RES line, mpi_insim.cpp:456-460 */
  int gmres_its;
  double gmres_res;
  int cg_mp_its, cg_sm_its, a_inv_its, precond_applies;
  double true_res; /* |b - A x|_2 / |b|_2 of the linear solve, recomputed in fp64 with the operator after FGMRES returned
                      (the parity pin of the inexact / reduced-precision inner solves: they act inside the preconditioner only) */
} ifem_newton_record;

/* InsIM(tria, parameters): the solver keeps a reference to the caller-owned triangulation
 * (mpi_fluid_solver.h:187) and a copy of the parameters (:222) */
int ifem_insim_create(ifem_tria *tria, const ifem_params *params, ifem_insim **out);
int ifem_insim_destroy(ifem_insim *s);
/* FluidSolver::add_hard_coded_boundary_condition(id, f(point, component, time)) (include/mpi_fluid_solver.h:104-108): used
 * by make_constraints for the Dirichlet ids of the .prm when "Use hard-coded boundary values = 1" semantics are wanted;
 * call before ifem_insim_setup / ifem_insim_run */
typedef double (*ifem_bc_fn)(const double *point, unsigned int component, double time, void *user);
int ifem_insim_add_hard_coded_boundary_condition(ifem_insim *s, int boundary_id, ifem_bc_fn f, void *user);
int ifem_insim_default_control(int serial_twin, ifem_ins_control *out);
/* the tolerances the solver currently runs with (each solver class starts from the reference's values for that class) */
int ifem_insim_get_control(const ifem_insim *s, ifem_ins_control *out);
int ifem_insim_set_control(ifem_insim *s, const ifem_ins_control *c);
int ifem_insim_set_verbose(ifem_insim *s, int verbose);
/* setup_dofs(); make_constraints(); initialize_system();  (mpi_insim.cpp:504-506) - no refinement */
int ifem_insim_setup(ifem_insim *s);
/* what FSI::run does to the fluid before its loop (source/mpi_fsi.cpp:1136-1142): refine_global(Global refinements[0]),
 * setup_dofs(), make_constraints(), initialize_system(); a no-op when the solver is already set up */
int ifem_insim_setup_with_refinement(ifem_insim *s);
/* run(): refine_global(Global refinements[0]) + setup + time loop (mpi_insim.cpp:492-519) */
int ifem_insim_run(ifem_insim *s);
/* run_one_step(apply_nonzero_constraints) (mpi_insim.cpp:397-490) */
int ifem_insim_run_one_step(ifem_insim *s, int apply_nonzero_constraints);
/* assemble(use_nonzero_constraints) (mpi_insim.cpp:152-362) at the current evaluation_point */
int ifem_insim_assemble(ifem_insim *s, int use_nonzero_constraints);
/* solve(use_nonzero_constraints) (mpi_insim.cpp:364-395): newton_update from the assembled system */
int ifem_insim_solve(ifem_insim *s, int use_nonzero_constraints, unsigned int *its, double *res);
int ifem_insim_sizes(const ifem_insim *s, int64_t *n_u, int64_t *n_p, int64_t *nnz_system, int64_t *nnz_mp, int64_t *nnz_schur);
/* support point of every dof, [n_dofs][dim] */
int ifem_insim_support_points(const ifem_insim *s, double *pts);
/* get_current_solution() (mpi_fluid_solver.h:113): block vector [u | p] to a host buffer */
int ifem_insim_get_current_solution(ifem_insim *s, double *host);
/* which: 0 present_solution, 1 evaluation_point, 2 fsi_acceleration, 3 newton_update, 4 system_rhs, 5 diag(M_u) (n_u) */
int ifem_insim_set_vector(ifem_insim *s, int which, const double *host);
int ifem_insim_get_vector(ifem_insim *s, int which, double *host);
int ifem_insim_set_indicator(ifem_insim *s, const int *host_indicator);
/* assembled operators as scalar CSR on the host (parity tests): which = 0 system_matrix, 1 M_p, 2 mass_schur */
int ifem_insim_get_matrix(ifem_insim *s, int which, int64_t *rowptr, int *col, double *val);
/* y = system_matrix * x with host buffers (BlockSparseMatrix::vmult) */
int ifem_insim_vmult(ifem_insim *s, const double *x_host, double *y_host);
int ifem_insim_history(const ifem_insim *s, int max_records, ifem_newton_record *out, int *n_records);
/* accumulated device milliseconds of a TimerOutput section ("Assemble system", "Solve linear system",
 * "CG for Mp", "CG for Sm", "A_inv") */
int ifem_insim_timer_ms(const ifem_insim *s, const char *section, double *ms);
int ifem_insim_time(const ifem_insim *s, unsigned int *timestep, double *current);
/* this rank's share of the dofs: vectors exchanged with the host are LOCAL block vectors
 * [u of local nodes (owned first, then ghosts) | p of local nodes]; which: 0 velocity nodes, 1 pressure nodes */
int ifem_insim_partition(const ifem_insim *s, int which, int *n_owned_nodes, int *n_local_nodes);
int ifem_insim_local_to_global(const ifem_insim *s, int which, int *global_node_ids);

/* ---- result files and checkpoints (formats: openifem_b200/csrc/output.h). FluidSolver::output_results / save_checkpoint /
 *      load_checkpoint (source/mpi_fluid_solver.cpp:491-713). Nothing is written until a directory is set; then run() and
 *      run_one_step() write fluid_NNNNNN.pvtu + .procRRRR.vtu pieces + fluid.pvd at the output interval and
 *      NNNNNN.fluid_checkpoint at the save interval, and run() restarts from the latest checkpoint in the directory.
 *      dir = NULL or "" switches it off again. ---- */
int ifem_insim_set_output_directory(ifem_insim *s, const char *dir);
int ifem_insim_output_results(ifem_insim *s, unsigned int output_index);
int ifem_insim_save_checkpoint(ifem_insim *s, int output_index);
int ifem_insim_load_checkpoint(ifem_insim *s, int *loaded);
/* time.current() and time.get_timestep() of the solver (Utils::Time) */
int ifem_insim_get_time(const ifem_insim *s, double *time, unsigned int *timestep);
/* host-side pieces of the formats, usable without a device: an ASCII .vtu from lexicographically ordered quad / hex cells,
 * point fields [n_points][ncomp] and cell fields [n_cells]; a .pvd collection as Utils::PVDWriter writes it (source/
 * utilities.cpp:38-81); deal.II Vector<double>::block_write / block_read streams (the solid checkpoint files) */
int ifem_write_vtu(const char *path, int dim, int64_t n_points, const double *points, int64_t n_cells, const int *cells,
                   int n_point_fields, const char *const *point_names, const int *point_ncomp, const double *const *point_data,
                   int n_cell_fields, const char *const *cell_names, const double *const *cell_data);
int ifem_write_pvd(const char *path, const char *pvtu_prefix, int n, const double *times, const unsigned int *timesteps);
int ifem_block_write(const char *path, int64_t n, const double *values);
int ifem_block_read(const char *path, int64_t capacity, double *values, int64_t *n);
/* FluidSolver::output_results on host arrays (what ifem_insim_output_results does after downloading the solver state):
 * the triangulation with FE_Q(pu) / FE_Q(pp) numbering of this library, present_solution / fsi_acceleration [dim n_u + n_p],
 * indicator [n_cells] or NULL, stress [dim dim][n_velocity_nodes] or NULL -> <dir>/fluid_NNNNNN.pvtu + piece */
int ifem_fluid_write_results_host(const ifem_tria *tria, int pu, int pp, const double *present, const double *fsi_acceleration,
                                  const int *indicator, const double *stress, const char *dir, unsigned int output_index);

/* ---- Fluid::MPI::InsIMEX<dim> (include/mpi_insimex.h, source/mpi_insimex.cpp): implicit-explicit twin of InsIM, Q2/Q1.
 *      The handle is an ifem_insim: setup, run (the time loop of mpi_insimex.cpp:449-480), vectors, matrices, history and
 *      timers apply; newton_update holds solution_time_increment. The entry points below carry the reference's second
 *      argument `assemble_system` (matrix + preconditioner rebuilt, or right-hand side only). ---- */
int ifem_insimex_create(ifem_tria *tria, const ifem_params *params, ifem_insim **out);
/* InsIMEX::assemble(use_nonzero_constraints, assemble_system) (source/mpi_insimex.cpp:150-355) */
int ifem_insimex_assemble(ifem_insim *s, int use_nonzero_constraints, int assemble_system);
/* InsIMEX::solve(use_nonzero_constraints, assemble_system) (:357-386): FGMRES iterations and residual */
int ifem_insimex_solve(ifem_insim *s, int use_nonzero_constraints, int assemble_system, unsigned int *iterations, double *residual);
/* InsIMEX::run_one_step(apply_nonzero_constraints, assemble_system) (:388-447) */
int ifem_insimex_run_one_step(ifem_insim *s, int apply_nonzero_constraints, int assemble_system);

/* ---- Fluid::MPI::SCnsIM<dim> (include/mpi_scnsim.h, source/mpi_scnsim.cpp:15-568) on SUPGFluidSolver
 *      (source/mpi_supg_solver.cpp). The handle is an ifem_insim: every ifem_insim_* entry point (setup, run,
 *      run_one_step, assemble, solve, vectors, matrices, history) applies. Q1/Q1 elements. ---- */
typedef double (*ifem_field_fn)(const double *point, unsigned int component, void *user);
int ifem_scnsim_create(ifem_tria *tria, const ifem_params *params, ifem_insim **out);
/* Fluid::MPI::SUPGInsIM<dim>(tria, params) (include/mpi_insim_supg.h, source/mpi_insim_supg.cpp): incompressible
 * Navier-Stokes with SUPG / PSPG / LSIC stabilisation, Q1/Q1, on the same SUPGFluidSolver machinery; every ifem_insim_* and
 * ifem_scnsim_set_* entry point applies (the PML field and the FSI terms have no effect in this solver) */
int ifem_supg_insim_create(ifem_tria *tria, const ifem_params *params, ifem_insim **out);
/* set_body_force / set_sigma_pml_field / set_initial_condition (include/mpi_fluid_solver.h:120-143); call before setup */
int ifem_scnsim_set_body_force(ifem_insim *s, ifem_field_fn f, void *user);
int ifem_scnsim_set_sigma_pml_field(ifem_insim *s, ifem_field_fn f, void *user);
int ifem_scnsim_set_initial_condition(ifem_insim *s, ifem_field_fn f, void *user);
/* update_stress() (source/mpi_fluid_solver.cpp:716-811): nodal viscous stress from present_solution */
int ifem_scnsim_update_stress(ifem_insim *s);
/* which: 0 stress [dim*dim][n_velocity_nodes], 1 fsi_stress [dim(dim+1)/2][n_velocity_nodes] */
int ifem_scnsim_get_field(ifem_insim *s, int which, double *host);
int ifem_scnsim_set_field(ifem_insim *s, int which, const double *host);

/* ---- Fluid::MPI::SpalartAllmaras<dim> (include/mpi_spalart_allmaras.h, source/mpi_spalart_allmaras.cpp) on
 *      Fluid::MPI::TurbulenceModel (include/mpi_turbulence_model.h): the turbulence model of an SCnsIM solver. The model lives
 *      inside the fluid solver's handle, like the reference's FluidSolver::turbulence_model. ---- */
/* FluidSolver::attach_turbulence_model(model_name) (source/mpi_fluid_solver.cpp:53-63); "Spalart-Allmaras" is the one model the
 * reference's factory knows (source/mpi_turbulence_model.cpp:11-26). Before or after ifem_insim_setup; from then on ifem_insim_run
 * advances the model before every fluid step (source/mpi_supg_solver.cpp:456-468) and SCnsIM::assemble adds the eddy viscosity
 * (source/mpi_scnsim.cpp:198-216). Equal-order Q1/Q1 solvers. */
int ifem_insim_attach_turbulence_model(ifem_insim *s, const char *model_name);
/* vectors over the scalar support points (pressure-node numbering of this rank): 0 present_solution (nu~), 1 evaluation_point,
 * 2 eddy_viscosity (get_eddy_viscosity(), include/mpi_turbulence_model.h:40), 3 system_rhs, 4 newton_update,
 * 5 fixed wall distance (setup_cell_property, :415-552). set: 0, 1, 2 */
int ifem_turbulence_get_vector(ifem_insim *s, int which, double *host);
int ifem_turbulence_set_vector(ifem_insim *s, int which, const double *host);
/* SpalartAllmaras::assemble(use_nonzero_constraints) (:620-832); the system matrix is ifem_insim_get_matrix(s, 3, ...) with the
 * pattern of M_p */
int ifem_turbulence_assemble(ifem_insim *s, int use_nonzero_constraints);
/* run_one_step(apply_nonzero_constraints) (:296-349): Newton loop + update_eddy_viscosity (:864-889) */
int ifem_turbulence_run_one_step(ifem_insim *s, int apply_nonzero_constraints);
/* update_boundary_condition(first_step) (:133-224): lines of the cells inside the immersed solid (CellProperty::indicator == 1) */
int ifem_turbulence_update_boundary_condition(ifem_insim *s, int first_step);
/* get_shear_velocity(vel, init_guess) (:227-293) */
int ifem_turbulence_get_shear_velocity(ifem_insim *s, double vel, double init_guess, double *out);
/* Newton records of the model (the " ITR = .. ABS_RES = .." lines of :333-338): abs_res / gmres_its of the last max_records */
int ifem_turbulence_history(ifem_insim *s, int max_records, double *abs_res, int *gmres_its, int *n_records);

/* ---- Solid::MPI::HyperElasticity<dim> (include/mpi_hyper_elasticity.h:96-176, source/mpi_hyper_elasticity.cpp;
 *      base class include/mpi_solid_solver.h:75-79, source/mpi_solid_solver.cpp) ---- */
typedef struct ifem_hyper ifem_hyper;
typedef struct
{
  unsigned int timestep, iteration;
  double res_F, res_U; /* the "res_F = / res_U =" line, mpi_hyper_elasticity.cpp:174-179 */
  int cg_its;
} ifem_solid_record;
int ifem_hyper_create(ifem_tria *tria, const ifem_params *params, ifem_hyper **out);
/* the class named explicitly: shared = 0 Solid::MPI::HyperElasticity<dim> (source/mpi_hyper_elasticity.cpp), shared = 1
 * Solid::MPI::SharedHyperElasticity<dim> (source/mpi_shared_hyper_elasticity.cpp: Newton loop also stops on |update| <= 1e-12, :125-127;
 * update_strain_and_stress after every step, :204-205). ifem_hyper_create picks the twin when `Simulation type = FSI`. */
int ifem_hyper_create_twin(ifem_tria *tria, const ifem_params *params, int shared, ifem_hyper **out);
/* Solid::MPI::LinearElasticity<dim>(tria, params) (shared = 0; include/mpi_linear_elasticity.h, source/mpi_linear_elasticity.cpp)
 * or Solid::MPI::SharedLinearElasticity<dim>(tria, params) (shared = 1, the replicated twin MPI::FSI takes;
 * source/mpi_shared_linear_elasticity.cpp). The handle is the common solid-solver handle: every ifem_hyper_* entry point
 * applies except update_qph / get_qph (hyperelastic only). */
int ifem_linear_elasticity_create(ifem_tria *tria, const ifem_params *params, int shared, ifem_hyper **out);
int ifem_hyper_destroy(ifem_hyper *s);
/* SharedSolidSolver::output_results / save_checkpoint / load_checkpoint (source/mpi_shared_solid_solver.cpp:237-337, 452-571):
 * solid_NNNNNN.pvtu + piece + solid.pvd; NNNNNN.solid_checkpoint_{displacement,velocity,acceleration} in deal.II's
 * block_write format. Off until a directory is set. */
int ifem_hyper_set_output_directory(ifem_hyper *s, const char *dir);
int ifem_hyper_output_results(ifem_hyper *s, unsigned int output_index);
int ifem_hyper_save_checkpoint(ifem_hyper *s, int output_index);
int ifem_hyper_load_checkpoint(ifem_hyper *s, int *loaded);
int ifem_hyper_get_time(const ifem_hyper *s, double *time, unsigned int *timestep);
int ifem_hyper_set_verbose(ifem_hyper *s, int verbose);
/* setup_dofs(); initialize_system() (incl. setup_qph) - no refinement */
int ifem_hyper_setup(ifem_hyper *s);
/* what FSI::run does to the solid before its loop (source/mpi_fsi.cpp:1127, 1134-1135): refine_global(Global
 * refinements[1]), setup_dofs(), initialize_system(); a no-op when the solver is already set up */
int ifem_hyper_setup_with_refinement(ifem_hyper *s);
/* run(): refine_global(Global refinements[1]) + setup + time loop (mpi_solid_solver.cpp:316-328) */
int ifem_hyper_run(ifem_hyper *s);
/* run_one_step(first_step) (mpi_hyper_elasticity.cpp:83-207) */
int ifem_hyper_run_one_step(ifem_hyper *s, int first_step);
/* update_qph(current_displacement) (:241-275) */
int ifem_hyper_update_qph(ifem_hyper *s);
/* assemble_system(initial_step) (:317-535): initial -> mass_matrix, else system_matrix; system_rhs in both */
int ifem_hyper_assemble_system(ifem_hyper *s, int initial_step);
int ifem_hyper_sizes(const ifem_hyper *s, int64_t *n_dofs, int64_t *nnz, int64_t *n_quadrature_points, int *nsym);
/* get_current_solution(): current_displacement (mpi_solid_solver.cpp:331-334) */
int ifem_hyper_get_current_solution(ifem_hyper *s, double *host);
/* which: 0 current_displacement 1 current_velocity 2 current_acceleration 3 previous_displacement
 *        4 previous_velocity 5 previous_acceleration 6 system_rhs */
int ifem_hyper_set_vector(ifem_hyper *s, int which, const double *host);
int ifem_hyper_get_vector(ifem_hyper *s, int which, double *host);
/* which: 0 system_matrix, 1 mass_matrix, 2 stiffness_matrix, 3 damping_matrix (2, 3: linear elasticity only); scalar CSR on the host */
int ifem_hyper_get_matrix(ifem_hyper *s, int which, int64_t *rowptr, int *col, double *val);
/* PointHistory arrays [cell][q]: F_inv [dim*dim], tau [dim*dim], Jc [nsym*nsym] (Voigt pairs (0,0),(1,1)[,(2,2)],(0,1)[,(0,2),(1,2)]), det F */
int ifem_hyper_get_qph(ifem_hyper *s, double *F_inv, double *tau, double *Jc, double *det_F);
/* update_strain_and_stress() of the shared solid solvers (source/mpi_shared_hyper_elasticity.cpp:599-714);
 * which: 0 stress, 1 strain, each [dim*dim][n_nodes] */
int ifem_hyper_update_strain_and_stress(ifem_hyper *s);
int ifem_hyper_get_nodal_tensor(ifem_hyper *s, int which, double *host);
int ifem_hyper_set_nodal_tensor(ifem_hyper *s, int which, const double *host);
/* what MPI::FSI::find_solid_bc wrote: fsi_stress_rows [dim][n_dofs], fluid_velocity [n_dofs], fluid_pressure [n_nodes] */
int ifem_hyper_get_fsi_inputs(ifem_hyper *s, double *fsi_stress_rows, double *fluid_velocity, double *fluid_pressure);
int ifem_hyper_history(const ifem_hyper *s, int max_records, ifem_solid_record *out, int *n_records);

/* ---- MPI::FSI<dim> immersed coupling kernels (include/mpi_fsi.h:39-47, source/mpi_fsi.cpp:95-119, 143-224,
 *      292-319, 324-663). The solid is the replicated solver the reference's MPI::FSI takes
 *      (SharedSolidSolver, mpi_fsi.h:39-40). ---- */
typedef struct ifem_fsi ifem_fsi;
/* FSI(fluid, solid, parameters, use_dirichlet_bc): holds references to both solvers (mpi_fsi.h:135-136) */
int ifem_fsi_create(ifem_insim *fluid, ifem_hyper *solid, const ifem_params *params, int use_dirichlet_bc, ifem_fsi **out);
int ifem_fsi_destroy(ifem_fsi *f);
/* update_solid_box(): box[2*dim] = min0, max0, min1, max1, ... of the solid moved by its current displacement */
int ifem_fsi_update_solid_box(ifem_fsi *f, double *box);
/* update_indicator(): CellProperty::indicator of every local fluid cell */
int ifem_fsi_update_indicator(ifem_fsi *f);
int ifem_fsi_get_indicator(ifem_fsi *f, int *indicator_host);
/* find_fluid_bc(): fills the fluid's fsi_acceleration (read it with ifem_insim_get_vector(s, 2, ..)); with
 * use_dirichlet_bc the inner constraints are merged into the fluid's (left object wins) */
int ifem_fsi_find_fluid_bc(ifem_fsi *f);
/* inner constraints of the last find_fluid_bc before the merge: flag and inhomogeneity per local fluid dof */
int ifem_fsi_get_inner_constraints(ifem_fsi *f, unsigned char *flags, double *inhomogeneity);
/* point_in_solid for a batch of points [n][dim] on the current deformed solid */
int ifem_fsi_point_in_solid(ifem_fsi *f, int n, const double *points, int *inside);
/* GridInterpolator::point_value of a solid field (0 velocity, 1 acceleration, 2 displacement): values [n][dim],
 * found[n] = index of the solid cell used or -1 (value 0, as utilities.cpp:228-233) */
int ifem_fsi_interpolate(ifem_fsi *f, int which, int n, const double *points, double *values, int *found);
/* find_solid_bc() (source/mpi_fsi.cpp:666-867): fills the solid's fsi_stress_rows / fluid_velocity / fluid_pressure */
int ifem_fsi_find_solid_bc(ifem_fsi *f);
/* one pass of the time loop of FSI::run (source/mpi_fsi.cpp:1172-1214) / the whole loop (no refinement, no checkpoints) */
/* FSI::set_penetration_criterion(criterion, direction) (include/mpi_fsi.h:44-47, source/mpi_fsi.cpp:1229-1237): with a
 * criterion set, the solid step of the time loop goes through apply_contact_model (:869-970); direction [dim] */
typedef double (*ifem_point_fn)(const double *point, void *user);
int ifem_fsi_set_penetration_criterion(ifem_fsi *f, ifem_point_fn criterion, void *user, const double *direction);
/* solid steps taken inside apply_contact_model so far */
int ifem_fsi_contact_iterations(const ifem_fsi *f, int *n);
int ifem_fsi_run_one_step(ifem_fsi *f, int first_step);
/* the same pass up to and including find_fluid_bc, without the fluid time step: the constraints, indicator, fsi_stress and
 * fsi_acceleration the fluid solver is about to see (tests compare them and the assembled system with the oracle) */
int ifem_fsi_prepare_fluid_step(ifem_fsi *f, int first_step);
int ifem_fsi_run(ifem_fsi *f);
/* FSI::refine_mesh(min_grid_level, max_grid_level) (source/mpi_fsi.cpp:1024-1117): flags from the distance of every fluid cell to the
 * boundary of the deformed solid, coarsening and refinement of the fluid triangulation, solution transfer, new fluid system.
 * ifem_fsi_run calls it like the reference when `Refinement interval` < `End time` (:1164-1168, :1215-1218) */
int ifem_fsi_refine_mesh(ifem_fsi *f, unsigned int min_grid_level, unsigned int max_grid_level);
/* n_steps passes of the coupled loop (ifem_fsi_run_one_step) between two CUDA events on the library's stream: the span covers
 * the device work and the host orchestration between the kernels (bench.py) */
int ifem_fsi_bench_steps(ifem_fsi *f, int n_steps, int first_step, double *ms_total);
int ifem_fsi_timer_ms(const ifem_fsi *f, const char *section, double *ms);

/* kernel shape of the row-plane BCSR mat-vec for short rows and the off-diagonal blocks (tuning sweeps, scripts/spmv_short_sweep.py):
 * key = 10 * lanes per block row + unroll, 0 = default */
int ifem_set_spmv_short_variant(int key);
/* ILU(0) of a scalar CSR matrix (sorted columns, diagonal present) and one application x = U^-1 L^-1 b on the device - the factors
 * behind the block preconditioner of the SUPG solvers (Hypre Euclid in the reference, source/preconditioner_pilut.cpp:124-138).
 * factors [nnz] (L strictly lower with unit diagonal implied, U upper, in place) or NULL; level counts of the two sweeps (tests) */
int ifem_ilu0_apply(int n, const int64_t *rowptr, const int *col, const double *val, const double *b, double *factors, double *x,
                    int *n_levels_lower, int *n_levels_upper);

/* ---- measurement hooks (bench.py): device-resident, CUDA-event timed on the library's stream ---- */
/* reps applications of the block SpMV on resident vectors; returns mean ms per application and the
 * algorithmic bytes of one application */
int ifem_insim_bench_vmult(ifem_insim *s, int reps, double *ms_per_apply, double *bytes_per_apply);
/* same for the velocity-velocity block alone (the dominant kernel) */
int ifem_insim_bench_spmv_uu(ifem_insim *s, int reps, double *ms_per_apply, double *bytes_per_apply);
/* fp32-streamed copy of A_uu (inner solve only) */
int ifem_insim_bench_spmv_uu_fp32(ifem_insim *s, int reps, double *ms_per_apply, double *bytes_per_apply);
/* "CG for Sm" on its own (test hook): x = S_m^-1 b to rel_tol * |b| in cg_sm_fp32 mode 0 / 1 / 2; b, x host vectors of
 * the local pressure size. S_m must exist (ifem_insim_solve has run) */
int ifem_insim_solve_mass_schur(ifem_insim *s, int mode, const double *b, double rel_tol, int max_it, double *x, int *its,
                                double *residual);
/* product kernel the fp32 inner solver uses from now on (tuning hook; see InnerSolver32::spmv for the encoding) */
int ifem_insim_set_inner_variant(ifem_insim *s, int variant);
/* the product kernel of the fp32 inner solver on the sliced (SELL-32) copy of A_uu (a_inv_fp32 = 2 / 3); precision = 32
 * or 16 (storage of the matrix values; 0: as built); variant = kernel shape (0: default); padding = stored / actual
 * blocks; max_rel_err (may be NULL) = max |y32 - y64| / max |y64| against the fp64 product with the present rhs as x */
int ifem_insim_bench_spmv_uu_sell(ifem_insim *s, int precision, int variant, int reps, double *ms_per_apply,
                                  double *bytes_per_apply, double *padding, double *max_rel_err);
int ifem_insim_bench_assemble(ifem_insim *s, int reps, double *ms_per_assembly);
/* n_steps calls of run_one_step bracketed by CUDA events on the library's stream; total device ms */
int ifem_insim_bench_steps(ifem_insim *s, int n_steps, int first_applies_nonzero_constraints, double *ms_total);

#ifdef __cplusplus
}
#endif
#endif
