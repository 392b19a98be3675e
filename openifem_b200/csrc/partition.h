// Domain decomposition of the fluid mesh over the ranks of one NVSwitch box
// (SURVEY 8e). The reference partitions cells along p4est's space-filling curve
// (parallel::distributed::Triangulation, include/mpi_fluid_solver.h:187) and lets
// every rank own a contiguous DoF range per block (source/mpi_fluid_solver.cpp:
// 140-152); PETSc then ships off-rank matrix rows in compress(add)
// (mpi_insim.cpp:359-361) and ghost values in VecGhostUpdate.
//
// Here cells are ordered along z-major slabs (on a uniform all-to-all NVSwitch
// fabric a 1-D slab partition with <= 2 neighbours is as good as 3-D blocks) and a
// node belongs to the lowest rank of the cells around it (deal.II's rule). Each
// rank assembles "owner computes": its local cell set is every cell that touches
// an owned node, so the off-rank row shipment disappears; what remains is the
// ghost-value halo exchange. Local numbering: owned nodes first (ascending global
// id), then ghosts grouped by owner (ascending global id inside a group), so a
// neighbour's message lands contiguously in the ghost tail of a vector.
#pragma once
#include <vector>

#include "mesh.h"

namespace ifem
{
  // Nodes of a rank come in three classes: owned | ghost layer 1 (nodes of cells that touch an owned node) |
  // ghost layer 2 (nodes of cells that touch a layer-1 node). Operator rows are owned; layer 1 is what a
  // mat-vec reads; layer 2 exists so that the rows of B^T for layer-1 velocity nodes - and with them the
  // explicit Schur complement B diag(M_u)^-1 B^T of the owned pressure rows - can be formed without
  // communication. Messages: one per (layer, neighbour), in that order on both sides.
  struct NodePartition
  {
    int n_owned = 0, n_layer1 = 0, n_local = 0; // n_layer1 = owned + layer-1 ghosts
    std::vector<int> local_to_global;          // [n_local]
    std::vector<int> neighbours;               // rank of every message (a rank appears once per layer)
    std::vector<std::vector<int>> send_local;  // per message: owned local ids the neighbour holds as ghosts
    std::vector<int> recv_offset, recv_count;  // per message: ghost range (local id = recv_offset, count)
  };

  struct Partition
  {
    int rank = 0, size = 1;
    std::vector<int> local_cells; // global cell ids, ascending: every cell within two layers of the owned nodes
    std::vector<char> cell_layer; // 1: touches an owned node (assembled every time), 2: only needed for the Schur pass
    NodePartition u, p;
  };

  // rank of every cell: equal contiguous chunks of the (z, y, x)-sorted cell sequence
  std::vector<int> slab_cell_ranks(const Triangulation &tria, int size);

  Partition build_partition(const Triangulation &tria, const NodeTable &un, const NodeTable &pn, int rank, int size);

  // local node table (cell_nodes in local ids, coords of local nodes) from a global one
  NodeTable localise(const NodeTable &global, const std::vector<int> &local_cells, const NodePartition &np);
} // namespace ifem
