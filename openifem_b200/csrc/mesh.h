// Minimal stand-ins for the deal.II mesh / DoF objects the OpenIFEM hot path is
// written against (deal.II is not available in this image): a hexahedral /
// quadrilateral Triangulation with boundary ids, GridGenerator box meshes with
// colorize = true (boundary id = 2*axis + side), uniform refinement, FE_Q(p)
// node numbering (p = 1, 2) and node-level sparsity patterns.
//
// Reference: triangulations are created in the test drivers
// (tests/fluid_cavity/fluid_cavity.cpp:28-34, tests/fluid_pipe_mpi/fluid_pipe_mpi.cpp:37-45)
// and consumed by Fluid::MPI::FluidSolver::setup_dofs / initialize_system
// (source/mpi_fluid_solver.cpp:116-162, 305-365).
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace ifem
{
  // Chart of one coarse cell next to a circular hole (2-D): corners v[0..3] in lexicographic order, the edge t = 0
  // (v0 -> v1) is an arc about c (PolarManifold: linear in radius and angle), the other three edges are straight;
  // points inside follow the transfinite interpolation of the four edges (TransfiniteInterpolationManifold).
  struct PolarArcChart
  {
    double v[4][2], c[2];
    double r0 = 0, r1 = 0, a0 = 0, da = 0;
    void init();
    void eval(double s, double t, double *x) const;
  };

  struct Triangulation
  {
    int dim = 0;
    std::vector<double> vertices;    // [n_vertices][dim]
    std::vector<int> cells;          // [n_cells][2^dim], lexicographic (x fastest) vertex order
    std::vector<int> boundary_faces; // [n_bfaces][3] = (cell, face_no = 2*axis+side, boundary id)
    std::vector<int> material_id;    // [n_cells]
    // optional curved description honoured by refine_global (2-D): a cell with chart_of_cell >= 0 occupies the
    // rectangle chart_box = (s0, t0, s1, t1) of that chart and places its new vertices on it
    std::vector<int> chart_of_cell;  // [n_cells] or empty
    std::vector<double> chart_box;   // [n_cells][4]
    std::vector<PolarArcChart> charts;

    int n_vertices() const { return dim ? (int)(vertices.size() / dim) : 0; }
    int verts_per_cell() const { return 1 << dim; }
    int n_cells() const { return dim ? (int)(cells.size() / verts_per_cell()) : 0; }
    int n_boundary_faces() const { return (int)(boundary_faces.size() / 3); }
    int n_active_cells() const { return n_cells(); }

    // Hanging vertices of a locally refined mesh (one level of difference across an edge / face, as deal.II keeps it): vertex h
    // sits at the midpoint of a coarse edge (2 masters) or at the centre of a coarse face (3-D, 4 masters) and carries the mean of
    // its masters in a continuous FE_Q(1) field (DoFTools::make_hanging_node_constraints, source/mpi_fluid_solver.cpp:182-184)
    struct Hanging
    {
      int vertex = -1, n_masters = 0;
      int master[4] = {-1, -1, -1, -1};
    };
    std::vector<Hanging> hanging;
    std::vector<int> cell_level; // [n_cells] refinement level of every active cell (empty: all 0)

    // Refinement forest, kept so that cells can be coarsened again (FSI::refine_mesh, source/mpi_fsi.cpp:1024-1117): the 2^dim
    // children of a refined cell form a family; the parent is recovered from them (its corner c is corner c of child c, its
    // boundary faces are those of the children on that side) together with the family it belonged to itself.
    struct Family
    {
      int parent_family = -1, parent_child_no = 0; // what the parent was a child of (-1: a cell of the coarse mesh)
      int parent_level = 0, material = 1;
    };
    std::vector<Family> families;
    std::vector<int> cell_family;   // [n_cells] family of an active cell, -1 for coarse-mesh cells (empty: all -1)
    std::vector<int> cell_child_no; // [n_cells] which child of its parent (lexicographic)

    // values of a Q1 field on the new vertices in terms of the old ones (parallel::distributed::SolutionTransfer for FE_Q(1):
    // a vertex that existed keeps its value, a vertex created inside a refined cell interpolates the cell's corners)
    struct TransferPlan
    {
      std::vector<int64_t> ptr;    // [n_new_vertices + 1]
      std::vector<int> old_vertex; // old vertex ids
      std::vector<double> weight;
    };
    // cell->set_refine_flag() / set_coarsen_flag() + prepare_coarsening_and_refinement() + execute_coarsening_and_refinement():
    // one level up or down per call. A family is coarsened when all its children are active and flagged; flags are then adjusted
    // until neighbouring cells (sharing a vertex - p4est's full 2:1 balance) differ by at most one level: refinement wins over
    // a neighbour's coarsening and forces coarser neighbours to refine. Straight-sided meshes only.
    void execute_coarsening_and_refinement(const std::vector<unsigned char> &refine_flags, const std::vector<unsigned char> &coarsen_flags,
                                           TransferPlan *plan = nullptr);
    int n_levels() const;

    void refine_global(int times);
    // cell->set_refine_flag() on the flagged cells + execute_coarsening_and_refinement() (tests/fsi_leaflet_mpi/fsi_leaflet_mpi.cpp:
    // 66-76): flagged cells are replaced by their 2^dim children, the others stay; straight-sided meshes only (no charts).
    // Throws if the result would put a hanging vertex on a master that is itself hanging (deal.II would refine further cells
    // to keep the 2:1 balance; callers here flag whole bands).
    void execute_refinement(const std::vector<unsigned char> &refine_flags);
    // recompute `hanging` from the geometry of the active cells
    void find_hanging_vertices();
  };

  namespace GridGenerator
  {
    void subdivided_hyper_rectangle(Triangulation &tria, const std::vector<unsigned int> &repetitions, const double *p1,
                                    const double *p2, bool colorize);
    void hyper_cube(Triangulation &tria, int dim, double left, double right, bool colorize);
  } // namespace GridGenerator

  // Utils::GridCreator<dim>::flow_around_cylinder (reference source/utilities.cpp:343-574): channel
  // [0, 2.2] x [0, 0.41] (3-D: [-0.3, 2.2] x [0, 0.41]^2, the 2-D mesh extruded in 8 layers, flat refinement) around
  // a cylinder of diameter 0.1 about (0.2, 0.2). Boundary ids 2-D: 0 inflow, 1 outflow, 2 y = 0, 3 y = 0.41,
  // 4 cylinder; 3-D: 0 / 1 x, 2 / 3 y, 4 / 5 z, 6 cylinder.
  namespace GridCreator
  {
    void flow_around_cylinder(Triangulation &tria, int dim);
  }

  // FE_Q(p) node numbering on a triangulation: nodes are unique geometric entities
  // (vertices, edge / face / cell midpoints for p = 2), numbered in lexicographic
  // (z, y, x) order of their position so that rows of the operators stay local in
  // memory and a slab partition owns contiguous ranges.
  struct NodeTable
  {
    int p = 0, nodes_per_cell = 0, n_nodes = 0, dim = 0;
    std::vector<int> cell_nodes; // [n_cells][nodes_per_cell], local order lexicographic
    std::vector<double> coords;  // [n_nodes][dim]
  };
  NodeTable build_node_table(const Triangulation &tria, int p);

  // local node indices of FE_Q(p) on face 2*axis+side
  std::vector<int> face_local_nodes(int dim, int p, int face_no);

  // Node-level CSR pattern: row node r couples with every column node that shares a
  // cell with it (DoFTools::make_sparsity_pattern without coupling table,
  // source/mpi_fluid_solver.cpp:311-312). Columns sorted.
  struct Pattern
  {
    int n_rows = 0, n_cols = 0;
    std::vector<int64_t> rowptr;
    std::vector<int> col;
  };
  Pattern build_pattern(int n_cells, const int *row_table, int nr, int n_rows, const int *col_table, int ncl, int n_cols);

  // Pattern of B * B^T on pressure nodes (compute_mmult_pattern,
  // source/mpi_fluid_solver.cpp:326-329): p-nodes of all cells that share a vertex
  // with a cell containing the row node.
  Pattern build_schur_pattern(const Triangulation &tria, const NodeTable &pn);

  // Pattern of A * B for CSR patterns A (rows x mid) and B (mid x cols): union of B's rows over A's columns.
  Pattern product_pattern(const Pattern &A, const Pattern &B, int n_cols);

  // Greedy colouring: cells of one colour share no node of `table`.
  // Returns cell ids grouped by colour and the group offsets.
  void colour_cells(int n_cells, const int *table, int per_cell, int n_nodes, std::vector<int> &order,
                    std::vector<int> &offsets);
} // namespace ifem
