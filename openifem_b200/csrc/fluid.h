// Device-resident state of Fluid::MPI::FluidSolver<dim> (reference
// include/mpi_fluid_solver.h:185-287, source/mpi_fluid_solver.cpp:116-365): FE
// space, DoF numbering, Dirichlet constraints, sparsity patterns and the block
// matrices / vectors the assembly kernels write and the Krylov loop reads.
//
// DoF layout (one rank): velocity block node-major, u dof = dim*node + c; pressure
// block behind it, p dof = n_u + pnode (block structure [u | p] as produced by
// DoFRenumbering::component_wise with block_component = {0,..,0,1},
// mpi_fluid_solver.cpp:125-138). The 2x2 block system is stored as four BCSR
// matrices on NODE patterns: A_uu (dim x dim blocks), A_up (dim x 1), A_pu (1 x dim),
// A_pp (1 x 1, only allocated for the slightly compressible solver). Of the
// reference's full-pattern mass_matrix only what is ever read is stored:
// diag(M_u) and M_p (mpi_insim.cpp:44, 80-82).
#pragma once
#include <memory>

#include "fe_tables.h"
#include "halo.h"
#include "hanging.h"
#include "linalg.h"
#include "mesh.h"
#include "parameters.h"

namespace ifem
{
  struct FluidSpace
  {
    int dim = 0, pu = 0, pp = 0;
    int nu = 0, np = 0, nq = 0, nv = 0; // per cell: velocity nodes, pressure nodes, quadrature points, vertices
    int n_cells = 0;
    NodeTable un, pn;                     // LOCAL node tables of this rank (owned nodes first, then ghosts)
    int64_t n_u = 0, n_p = 0, n_dofs = 0; // local vector sizes: dim * un.n_nodes, pn.n_nodes, sum
    // domain decomposition (single rank: everything owned, no halos)
    int rank = 0, n_ranks = 1;
    Partition part;
    NodeTable un_global, pn_global;       // kept on multi-rank runs for the global constraint pass
    std::vector<int> local_cells;         // global cell id of each local cell
    int n_owned_unodes = 0, n_owned_pnodes = 0;
    int n_layer1_unodes = 0, n_layer1_pnodes = 0; // owned + layer-1 ghosts
    std::vector<int> colour_n1;                   // per colour: number of layer-1 cells (they come first)
    bool schur_valid = false;                     // S_m matches the current constraints (B and diag(M_u) do not depend on the solution)
    Halo halo_u, halo_p, halo_s;
    VecSpace vs_all, vs_u, vs_p;          // owned entries of a block / velocity / pressure vector
    // refresh ghost entries of a block vector [u | p]
    void halo_update(Context &ctx, double *x) { halo_u.update(ctx, x); halo_p.update(ctx, x + n_u); }

    // host FE tables (kept for face terms and point evaluation)
    FEQ fe_u, fe_p, fe_geo;
    Quadrature quad;
    ShapeTable tab_u, tab_p, tab_geo;

    // host patterns
    Pattern P_uu, P_up, P_pu, P_pp, P_schur;
    std::vector<int> colour_order, colour_offsets;

    // constraints (host mirror): flag per dof, nonzero value per dof
    std::vector<unsigned char> con;
    std::vector<double> nonzero_val;
    // hanging-node lines of a locally refined mesh (FE_Q(1) spaces): condensed after the cell loop, see hanging.h
    HangingConstraints hanging;

    // ---- device ----
    DevBuf<int> d_cell_un, d_cell_pn, d_colour_order;
    DevBuf<double> d_cell_x;  // [n_cells][nv][dim]
    DevBuf<double> d_tables;  // N[nq][nu] | dN[nq][nu][dim] | Np[nq][np] | dNgeo[nq][nv][dim] | qw[nq]
    DevBuf<double> d_tables_s; // what the INS assembly stages in shared memory (one TMA bulk copy): N | Np | dNgeo | qw, padded to 16 B
    DevBuf<unsigned char> d_slots; // per cell: uu[nu][nu] | up[nu][np] | pu[np][nu] | pp[np][np]
    DevBuf<unsigned char> d_con;
    DevBuf<double> d_nonzero_val;
    // the lines make_constraints() produced, kept on the device: a coupling step that re-makes the constraints every pass
    // (MPI::FSI::run, source/mpi_fsi.cpp:1190-1198) restores them with two device copies instead of a host pass + upload
    DevBuf<unsigned char> d_base_con;
    DevBuf<double> d_base_val;
    bool base_valid = false, flags_merged = false;
    void restore_base_constraints(Context &ctx);
    DevBuf<int> d_con_idx; // list of constrained dofs
    int n_con = 0;
    DevBuf<int> d_indicator; // CellProperty::indicator
    // boundary faces carrying a pressure Neumann condition: (cell, face_no) + value
    DevBuf<int> d_nface_cell;
    DevBuf<double> d_nface_val;
    DevBuf<double> d_face_tables; // per face: Nu_face[2*dim][nqf][nu] | dNgeo_face[2*dim][nqf][nv][dim] | qwf[nqf]
    int n_nfaces = 0, nqf = 0;

    Bcsr A_uu, A_up, A_pu, A_pp, M_p, S_m;
    DevBuf<double> d_qpt_to_dof;  // [nu][nq] projection from quadrature points to the dofs of scalar FE_Q(pu)
    DevBuf<double> d_stress_count; // [n_velocity_nodes] cells around every node (update_stress)
    DevBuf<double> diag_Mu; // [n_u]
    DevBuf<double> rhs;     // [n_dofs]

    int slots_per_cell() const { return nu * nu + 2 * nu * np + np * np; }

    void setup(Context &ctx, const Triangulation &tria, int pu, int pp, bool with_App);
    // Dirichlet constraints from the .prm maps, "first boundary id wins"
    // (mpi_fluid_solver.cpp:165-280). hard_coded(id, point, component) may be null.
    void make_constraints(Context &ctx, const Triangulation &tria,
                          const std::map<unsigned int, std::pair<unsigned int, std::vector<double>>> &dirichlet,
                          const std::function<bool(int, const double *, int, double &)> &hard_coded);
    void set_neumann_faces(Context &ctx, const Triangulation &tria, const std::map<unsigned int, double> &neumann);
  };

  struct InsAssembleParams
  {
    double viscosity, gamma, rho, dt;
    double gravity[3];
    // InsIMEX: matrix without the convection terms / right-hand side only (mpi_insimex.cpp:150-355)
    int explicit_convection = 0, rhs_only = 0;
  };

  // Fluid::MPI::InsIM<dim>::assemble (reference source/mpi_insim.cpp:152-362).
  // eval_pt / present / fsi_acc are block vectors of n_dofs doubles on the device;
  // fills A_uu, A_up, A_pu, M_p, diag_Mu, rhs of the space.
  // schur_pass: recompute only A_up / A_pu / diag(M_u), including the rows of layer-1 ghost velocity nodes
  void ins_assemble(Context &ctx, FluidSpace &fs, const InsAssembleParams &prm, const double *eval_pt, const double *present,
                    const double *fsi_acc, bool use_nonzero_constraints, bool assemble_mass, bool schur_pass = false);

  // pressure Neumann faces: rhs_i -= phi_i . n  p  JxW_face on unconstrained owned rows (mpi_insim.cpp:313-341,
  // mpi_scnsim.cpp:516-546)
  void neumann_faces(Context &ctx, FluidSpace &fs);

  // FluidSolver::update_stress (source/mpi_fluid_solver.cpp:716-811) into stress [dim*dim][n_velocity_nodes]; ghost
  // entries are refreshed by halo exchange
  void update_nodal_stress(Context &ctx, FluidSpace &fs, const double *present, double viscosity, double *stress);

  // y = A x on the 2x2 block system (BlockSparseMatrix::vmult)
  void block_vmult(Context &ctx, const FluidSpace &fs, const double *x, double *y);

  // S_m = B diag(M_u)^-1 B^T on the fixed Schur pattern (mpi_insim.cpp:44-49), owned pressure rows
  void compute_mass_schur(Context &ctx, FluidSpace &fs);
  // y_p = B diag(M_u)^-1 B^T x_p applied matrix-free (two SpMVs + halos): the multi-rank form of S_m
  void apply_mass_schur_matrix_free(Context &ctx, FluidSpace &fs, const double *x_p, double *y_p, double *tmp_u);
} // namespace ifem
