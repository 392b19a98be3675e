// Device side of peer.h: reductions that finish inside the kernel that produced the partial sums - block order sum by the
// last CTA to arrive, then (with an active peer link) the cross-rank sum through NVLink stores - and the ghost push / wait
// pair. Everything here is deterministic: sums are taken in CTA order and in rank order.
#pragma once
#include "peer.h"

namespace ifem
{
  // v[q] summed over the CTA; valid in thread 0 afterwards. blockDim.x a multiple of 32, at most 1024.
  template <int NR>
  __device__ __forceinline__ void cta_sum(double (&v)[NR])
  {
    __shared__ double sh[NR][32];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
#pragma unroll
    for (int q = 0; q < NR; ++q)
      {
        const double s = warp_sum(v[q]);
        if (l == 0) sh[q][w] = s;
      }
    __syncthreads();
    if (w == 0)
      {
#pragma unroll
        for (int q = 0; q < NR; ++q)
          {
            double s = l < nw ? sh[q][l] : 0.0;
            s = warp_sum(s);
            v[q] = s;
          }
      }
    __syncthreads();
  }

  // Sum of tot[0..NR) (valid in thread 0) over the ranks of the link; called by every thread of ONE CTA per rank.
  // Slot parity: a rank can only post reduction e + 1 after it has read all words of reduction e, and nobody completes
  // e + 1 without that post, so the words of e + 2 never land on unread words of e. `epoch` counts the reductions that
  // were actually carried out (launches that skip because a solver is done do not post), identically on every rank.
  template <int NR>
  __device__ __forceinline__ void peer_allreduce(double (&tot)[NR], const PeerDev &pd)
  {
    static_assert(NR <= kPeerMaxVals, "at most kPeerMaxVals doubles per all-reduce");
    constexpr int W = 2 * NR;
    __shared__ unsigned int words[kPeerMaxRanks][W];
    __shared__ double in[NR];
    if (threadIdx.x == 0)
      {
#pragma unroll
        for (int q = 0; q < NR; ++q) in[q] = tot[q];
      }
    __syncthreads();
    const unsigned int e = *pd.epoch + 1u;
    for (int t = threadIdx.x; t < pd.size * W; t += blockDim.x)
      {
        const int p = t / W, w = t % W;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(in[w >> 1]);
        const unsigned int half = (w & 1) ? (unsigned int)(bits >> 32) : (unsigned int)bits;
        const size_t slot = (size_t)(e & 1u) * kPeerMaxRanks;
        volatile unsigned long long *dst = pd.ll_remote[p] + (slot + pd.rank) * kPeerWords + w;
        *dst = ((unsigned long long)e << 32) | half;
        volatile const unsigned long long *src = pd.ll_local + (slot + p) * kPeerWords + w;
        unsigned long long v;
        do
          {
            v = *src;
          }
        while ((unsigned int)(v >> 32) != e);
        words[p][w] = (unsigned int)v;
      }
    __syncthreads();
    if (threadIdx.x == 0)
      {
#pragma unroll
        for (int q = 0; q < NR; ++q)
          {
            double s = 0.0;
            for (int p = 0; p < pd.size; ++p)
              s += __longlong_as_double((long long)(((unsigned long long)words[p][2 * q + 1] << 32) | words[p][2 * q]));
            tot[q] = s;
          }
        *pd.epoch = e;
      }
    __syncthreads();
  }

  // End of a reducing kernel, called by every thread of every CTA with its partial sums in acc. partials: [gridDim.x][NR];
  // counter: zero before the first launch (re-armed here). Returns true in all threads of the last CTA to arrive, with the
  // totals - over the CTAs, and over the ranks if the peer link is active - in acc of thread 0 and stored to red[0..NR).
  template <int NR>
  __device__ __forceinline__ bool finish_reduce(double (&acc)[NR], double *__restrict__ partials, unsigned int *__restrict__ counter,
                                                double *__restrict__ red, const PeerDev &pd)
  {
    cta_sum<NR>(acc);
    __shared__ int is_last;
    if (threadIdx.x == 0)
      {
#pragma unroll
        for (int q = 0; q < NR; ++q) partials[(size_t)blockIdx.x * NR + q] = acc[q];
        __threadfence();
        is_last = atomicAdd(counter, 1u) == gridDim.x - 1 ? 1 : 0;
      }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
#pragma unroll
    for (int q = 0; q < NR; ++q) acc[q] = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x)
      {
#pragma unroll
        for (int q = 0; q < NR; ++q) acc[q] += __ldcg(partials + (size_t)i * NR + q);
      }
    cta_sum<NR>(acc);
    if (threadIdx.x == 0) *counter = 0u;
    if (pd.active) peer_allreduce<NR>(acc, pd);
    if (threadIdx.x == 0)
      {
#pragma unroll
        for (int q = 0; q < NR; ++q) red[q] = acc[q];
      }
    return true;
  }

  // ---- ghost push --------------------------------------------------------------------------------------------------
  constexpr int kPeerMaxNeighbours = 4; // distinct neighbour ranks
  constexpr int kPeerMaxMsgs = 8;       // messages: a neighbour may receive several segments (one per ghost layer)

  template <typename T>
  struct PeerHaloDev
  {
    int n_msg = 0, n_nb = 0;
    int width = 1;                          // values per node
    int send_off[kPeerMaxMsgs + 1];         // in nodes, into send_pos: message m covers [send_off[m], send_off[m + 1])
    T *dst[kPeerMaxMsgs];                   // where message m lands in the receiver's vector
    unsigned int *flag[kPeerMaxNeighbours]; // the neighbour's arrival flag for this rank
    int nb_rank[kPeerMaxNeighbours];
    const unsigned int *my_flags = nullptr; // [kPeerMaxRanks] arrival flags of this rank, indexed by sender
    unsigned int *epoch = nullptr;          // pushes carried out so far
    unsigned int *counter = nullptr;        // CTA arrival counter of the push kernel
  };

  // x[send_pos[k]] -> the neighbours' ghost segments; the last CTA raises the arrival flags (after a system fence: all
  // CTAs fenced their stores before they arrived at the counter).
  template <typename T>
  __global__ void __launch_bounds__(256) peer_halo_push_kernel(PeerHaloDev<T> h, const int *__restrict__ send_pos, const T *__restrict__ x,
                                                                const int *__restrict__ skip)
  {
    if (skip && *skip) return;
    const int total = h.send_off[h.n_msg] * h.width;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x)
      {
        const int k = t / h.width, c = t - k * h.width;
        int m = 0;
        while (m + 1 < h.n_msg && k >= h.send_off[m + 1]) ++m;
        h.dst[m][(size_t)(k - h.send_off[m]) * h.width + c] = x[(size_t)send_pos[k] * h.width + c];
      }
    __threadfence_system();
    __syncthreads();
    __shared__ int is_last;
    if (threadIdx.x == 0) is_last = atomicAdd(h.counter, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    if (threadIdx.x == 0)
      {
        __threadfence_system();
        const unsigned int e = *h.epoch + 1u;
        for (int nb = 0; nb < h.n_nb; ++nb) *(volatile unsigned int *)h.flag[nb] = e;
        *h.epoch = e;
        *h.counter = 0u;
      }
  }

  // wait until every neighbour has pushed as many halos as this rank
  template <typename T>
  __global__ void peer_halo_wait_kernel(PeerHaloDev<T> h, const int *__restrict__ skip)
  {
    if (skip && *skip) return;
    const unsigned int e = *h.epoch;
    if ((int)threadIdx.x < h.n_nb)
      {
        volatile const unsigned int *f = h.my_flags + h.nb_rank[threadIdx.x];
        while ((int)(*f - e) < 0)
          {
          }
      }
    __syncthreads();
    __threadfence_system();
  }
} // namespace ifem
