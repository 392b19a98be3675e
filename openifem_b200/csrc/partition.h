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
  struct NodePartition
  {
    int n_owned = 0, n_local = 0;
    std::vector<int> local_to_global;          // [n_local]
    std::vector<int> neighbours;               // ranks exchanged with (ascending)
    std::vector<std::vector<int>> send_local;  // per neighbour: owned local ids it holds as ghosts
    std::vector<int> recv_offset, recv_count;  // per neighbour: ghost range (local id = recv_offset, count)
  };

  struct Partition
  {
    int rank = 0, size = 1;
    std::vector<int> local_cells; // global cell ids (ascending slab order)
    NodePartition u, p;
  };

  // rank of every cell: equal contiguous chunks of the (z, y, x)-sorted cell sequence
  std::vector<int> slab_cell_ranks(const Triangulation &tria, int size);

  Partition build_partition(const Triangulation &tria, const NodeTable &un, const NodeTable &pn, int rank, int size);

  // local node table (cell_nodes in local ids, coords of local nodes) from a global one
  NodeTable localise(const NodeTable &global, const std::vector<int> &local_cells, const NodePartition &np);
} // namespace ifem
