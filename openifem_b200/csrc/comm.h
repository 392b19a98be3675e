// Rank-to-rank communication for the domain-decomposed path: one
// process per GPU, an NCCL communicator over NVLink/NVSwitch. Only two operations
// exist on the hot path: the scalar all-reduce behind every Krylov dot product
// (reference: PETSc VecDot/VecNorm -> MPI_Allreduce) and the ghost-DoF halo
// exchange before an operator application (reference: PETSc VecGhostUpdate /
// MatMult VecScatter). NCCL is loaded at run time (dlopen of the torch-bundled
// libnccl.so.2) so that the library also loads on a single-GPU box without it.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

namespace ifem
{
  struct Comm
  {
    int rank = 0, size = 1;
    void *nccl = nullptr; // ncclComm_t
  };

  // 128-byte NCCL unique id, created on rank 0 and broadcast by the host launcher
  // (torch.distributed / env rendezvous in bench.py).
  void comm_get_unique_id(unsigned char id[128]);
  Comm *comm_create(int rank, int size, const unsigned char id[128]);
  void comm_destroy(Comm *c);

  void comm_allreduce_sum(Comm &c, double *dev, int n, cudaStream_t s);
  // exchange with at most two slab neighbours: send `send_lo`/`send_hi` counts from
  // packed device buffers, receive into ghost buffers
  void comm_sendrecv(Comm &c, int peer, const double *send, int64_t n_send, double *recv, int64_t n_recv, cudaStream_t s);
  void comm_sendrecv_f32(Comm &c, int peer, const float *send, int64_t n_send, float *recv, int64_t n_recv, cudaStream_t s);
  void comm_group_start(Comm &c);
  void comm_group_end(Comm &c);
} // namespace ifem
