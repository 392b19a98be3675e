// Ghost-DoF halo exchange (replaces PETSc VecGhostUpdate / the VecScatter inside
// MatMult; reference call sites: evaluation_point = tmp, source/mpi_insim.cpp:444-448,
// present_solution = ..., :473): owned values are packed by a gather kernel and
// sent with ncclSend/ncclRecv inside one group; every neighbour's message lands
// contiguously in the ghost tail of the vector, so there is no unpack step.
#pragma once
#include "device.cuh"
#include "partition.h"

namespace ifem
{
  struct Halo
  {
    int bs = 1; // doubles per node
    int n_owned = 0, n_local = 0;
    std::vector<int> neighbours, send_off, send_cnt, recv_off, recv_cnt; // in nodes
    DevBuf<int> d_send_idx;
    DevBuf<double> d_send_buf;
    int n_send_total = 0;

    void init(Context &ctx, const NodePartition &np, int block_size);
    // refresh the ghost entries of v (node-major, bs doubles per node) from their owners
    void update(Context &ctx, double *v);
    bool active() const { return !neighbours.empty(); }
  };
} // namespace ifem
