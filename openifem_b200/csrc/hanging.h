// Hanging-node constraints of a locally refined fluid mesh on the device.
//
// Reference: DoFTools::make_hanging_node_constraints into nonzero_constraints / zero_constraints
// (source/mpi_fluid_solver.cpp:182-184), resolved against the Dirichlet lines by AffineConstraints::close() (:273-274)
// and honoured by every distribute_local_to_global of the assembly loops (source/mpi_scnsim.cpp:548-560,
// source/mpi_insim_supg.cpp:310-322) and by constraints.distribute after each solve (source/mpi_supg_solver.cpp:323-325).
// The meshes that need it: the band-refined channels of tests/fsi_leaflet_mpi/fsi_leaflet_mpi.cpp:66-76 and
// tests/fsi-wall-3D/fsi-wall-3D.cpp:47-53 (BASELINE configs 4 and 5), FE_Q(1) velocity and pressure.
//
// Design: the cell kernels stay as they are and assemble a hanging dof like a free one (their in-kernel elimination only
// knows Dirichlet lines). A post-pass then condenses the four block matrices and the right-hand side through the lines,
// A <- C^T A C, b <- C^T (b - A g), which is what the reference's cell-wise scatter sums up to: (1) columns of hanging nodes
// are folded into the columns of their masters (a master that carries a Dirichlet value moves to the right-hand side
// instead), (2) rows of hanging nodes are added to the rows of their masters, (3) the hanging rows keep a diagonal entry and
// rhs = diagonal x inhomogeneity like every constrained row. The work is proportional to the refinement interface (a few
// hundred nodes at config 5), every target row is owned by one thread and visited in a fixed order: bitwise reproducible.
// Dirichlet flags are read at condensation time, so lines merged into the constraints during a run (FSI::find_fluid_bc,
// source/mpi_fsi.cpp:626-650) need no host-side re-resolution.
#pragma once
#include "device.cuh"
#include "linalg.h"
#include "mesh.h"

namespace ifem
{
  struct FluidSpace;

  // hanging nodes of one node space (velocity or pressure nodes of this rank, local numbering)
  struct HangingNodes
  {
    int n = 0;
    std::vector<int> node, n_masters, master; // master: [n][4]
    DevBuf<int> d_node, d_n_masters, d_master;
    DevBuf<double> d_diag; // [n][components]: |diagonal| of the hanging rows before the condensation
  };

  // what the condensation of one block matrix needs, built once from its pattern
  struct FoldPlan
  {
    int n_rows = 0;             // owned rows that hold at least one hanging column
    DevBuf<int> d_row, d_ptr;   // row ids and offsets into the items
    DevBuf<int> d_item;         // per item 6 ints: position of the hanging column in the row, index of the hanging node in the
                                // column space, positions of its (up to four) master columns in the row
    int n_mrows = 0;            // owned master rows
    DevBuf<int> d_mrow, d_mptr; // master row ids and offsets into the slaves
    DevBuf<int> d_mslave;       // index of the hanging node (row space) whose row is added to the master row
  };

  // 1 for every node of the table that sits on a hanging vertex of the triangulation (such a dof keeps its hanging-node line:
  // VectorTools::interpolate_boundary_values never overwrites an existing line)
  std::vector<char> hanging_node_flags(const Triangulation &tria, const NodeTable &nt);

  struct HangingConstraints
  {
    bool active = false;
    HangingNodes u, p;
    FoldPlan uu, up, pu, pp;
    bool with_pp = false;
    std::vector<unsigned char> is_hanging_dof; // [n_dofs] of the block vector (empty when inactive)
    DevBuf<unsigned char> d_is_hanging_dof;

    // hanging vertices of the triangulation -> lines on the local velocity / pressure nodes; per-cell node lists extended by
    // the masters (for the sparsity patterns). Throws unless both spaces are FE_Q(1).
    void find(const Triangulation &tria, const FluidSpace &fs, std::vector<int> &cell_un_ext, std::vector<int> &cell_pn_ext, int &width);
    // fold plans from the final patterns
    void plan(Context &ctx, const FluidSpace &fs);
    // condensation of A_uu / A_up / A_pu / A_pp and rhs; inhom = nonzero constraint values per dof or null (zero constraints)
    void condense(Context &ctx, FluidSpace &fs, const double *inhom) const;
    // x_h = sum_k w_k x_master(k) for every hanging dof of the block vector x (after the Dirichlet entries were set)
    void distribute(Context &ctx, const FluidSpace &fs, double *x) const;
    // the same condensation / distribution for ONE scalar system on the pressure node space (pattern and fold plan of A_pp):
    // the transport equation of a turbulence model (source/mpi_spalart_allmaras.cpp:817-826, :855-858). con / inhom are the
    // model's own lines, indexed by pressure node.
    void condense_scalar(Context &ctx, const FluidSpace &fs, Bcsr &A, double *rhs, const unsigned char *con, const double *inhom) const;
    void distribute_scalar(Context &ctx, double *x) const;
  };
} // namespace ifem
