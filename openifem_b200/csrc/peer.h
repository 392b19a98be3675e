// Peer-memory link between the ranks of one NVLink / NVSwitch node.
//
// The reference does its Krylov dot products with PETSc VecDot -> MPI_Allreduce and its ghost updates with VecScatter
// (behind the KSPCG / MatMult calls at source/mpi_insim.cpp:81, 105, 117, 388). A straight NCCL translation costs one
// collective launch (~10-25 us) per dot product and per halo, which at 8 GPUs is as long as the products of the inner
// solvers themselves. Here the two exchanges are part of the kernels that produce the data:
//   * all-reduce of up to 4 doubles: the last CTA of a reducing kernel stores its sums straight into every peer's
//     buffer over NVLink as 8-byte {value half, epoch} words (NCCL's "LL" idea: data and flag travel in one store, no
//     fence), polls its own buffer for the peers' words and adds them in rank order - every rank gets the same bits;
//   * halo: a pack kernel writes the owned boundary values directly into the neighbour's ghost segment and raises a
//     per-sender epoch flag there; a one-warp kernel waits for the flags of its own neighbours before the product runs.
// Buffers are exchanged as CUDA IPC handles over the existing communicator once at set-up. Without peer access (one
// rank, more than 8 ranks, IPC refused, IFEM_PEER=0, the CPU emulator of the tests) the same device-resident algorithms
// run with ncclAllReduce / ncclSend / ncclRecv between the kernels instead.
#pragma once
#include <cstdint>
#include <vector>

#include "device.cuh"

namespace ifem
{
  constexpr int kPeerMaxRanks = 8;
  constexpr int kPeerMaxVals = 4;              // doubles per all-reduce
  constexpr int kPeerWords = 2 * kPeerMaxVals; // 32-bit halves

  // by-value kernel argument
  struct PeerDev
  {
    int rank = 0, size = 1;
    int active = 0;                 // 1: peer stores; 0: the sums stay local (one rank, or NCCL follows the kernel)
    unsigned long long *ll_local = nullptr;            // [2][kPeerMaxRanks][kPeerWords]
    unsigned long long *ll_remote[kPeerMaxRanks] = {}; // the same buffer of every rank (own entry = ll_local)
    unsigned int *epoch = nullptr;  // reductions carried out so far (device resident: skipped launches do not count)
  };

  struct PeerLink
  {
    int rank = 0, size = 1;
    bool active = false;
    int mask = 3; // IFEM_PEER: bit 0 = all-reduces through the link, bit 1 = halos through the link (0 switches the link off)
    unsigned long long *ll = nullptr; // own LL buffer (shared)
    unsigned int *epoch = nullptr;
    std::vector<void *> ll_peers;
    std::vector<void *> opened;       // every mapping opened from a peer (closed in the destructor)
    std::vector<void *> owned;        // shared allocations of this rank

    ~PeerLink();
    // collective over the communicator; leaves active = false if any rank cannot share memory
    void init(Context &ctx);
    // collective: allocate `bytes` (rounded up to 2 MiB) on every rank and return everybody's mapping of everybody's
    // buffer: result[r] is rank r's buffer as addressable from this rank (result[rank] = own). Empty when inactive.
    std::vector<void *> alloc_shared(Context &ctx, size_t bytes);
    PeerDev dev() const;
  };

  // the link of this process (created on first use after the communicator exists; inactive on a single rank)
  PeerLink &peer_link(Context &ctx);
  void peer_link_reset(); // communicator torn down
  // collective self test of the link (tests, diagnostics): every rank shares three buffers of different sizes, writes a
  // rank-specific pattern into every peer's copy with a kernel, and checks what the peers wrote into its own; then runs
  // `rounds` all-reduces of known values. Returns the number of mismatches on this rank (0 = fine, -1 = link inactive).
  int64_t peer_selftest(Context &ctx, int rounds);

  // collective helper on the existing communicator: every rank contributes n int64 values, all get the size x n table
  std::vector<int64_t> comm_allgather_i64(Context &ctx, const std::vector<int64_t> &mine);
} // namespace ifem

namespace ifem
{
  // how the reducing kernels of the device-resident solvers complete a sum (peer_dev.cuh): inside the kernel (one rank, or
  // several ranks with a peer link: adv = 1) or followed by an NCCL all-reduce and a one-thread kernel (nccl = true)
  struct ReduceMode
  {
    PeerDev pd;
    int adv = 1;
    bool nccl = false;
  };
  ReduceMode reduce_mode(Context &ctx);
} // namespace ifem
