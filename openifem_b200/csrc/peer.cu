#include "peer.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "comm.h"
#include "peer_dev.cuh"

namespace ifem
{
  namespace
  {
    std::unique_ptr<PeerLink> g_link;
    constexpr size_t kShareGranule = size_t(2) << 20; // shared allocations are whole 2 MiB blocks of their own
    constexpr int kHandleBytes = (int)sizeof(cudaIpcMemHandle_t);
  } // namespace

  std::vector<int64_t> comm_allgather_i64(Context &ctx, const std::vector<int64_t> &mine)
  {
    const int size = ctx.comm ? ctx.comm->size : 1, rank = ctx.comm ? ctx.comm->rank : 0;
    const size_t n = mine.size();
    if (size == 1) return mine;
    // the communicator's only reduction is a sum of doubles: every rank fills its own row of a zero table
    std::vector<double> tab((size_t)size * n, 0.0);
    for (size_t k = 0; k < n; ++k) tab[(size_t)rank * n + k] = (double)mine[k];
    DevBuf<double> d(tab.size());
    d.upload(tab, ctx.stream);
    comm_allreduce_sum(*ctx.comm, d.p, (int)tab.size(), ctx.stream);
    tab = d.to_host(ctx.stream);
    std::vector<int64_t> out(tab.size());
    for (size_t k = 0; k < tab.size(); ++k) out[k] = (int64_t)tab[k];
    return out;
  }

  PeerLink::~PeerLink()
  {
    for (void *p : opened) cudaIpcCloseMemHandle(p);
    for (void *p : owned) cudaFree(p);
    if (epoch) cudaFree(epoch);
  }

  std::vector<void *> PeerLink::alloc_shared(Context &ctx, size_t bytes)
  {
    std::vector<void *> out;
    if (!active) return out;
    bytes = ((bytes + kShareGranule - 1) / kShareGranule) * kShareGranule;
    void *mine = nullptr;
    cudaIpcMemHandle_t h;
    std::memset(&h, 0, sizeof h);
    bool ok = cudaMalloc(&mine, bytes) == cudaSuccess;
    if (ok) ok = cudaMemsetAsync(mine, 0, bytes, ctx.stream) == cudaSuccess && cudaStreamSynchronize(ctx.stream) == cudaSuccess;
    if (ok) ok = cudaIpcGetMemHandle(&h, mine) == cudaSuccess;
    if (!ok) cudaGetLastError(); // clear the sticky-free error state
    std::vector<int64_t> msg(kHandleBytes + 1);
    for (int k = 0; k < kHandleBytes; ++k) msg[k] = reinterpret_cast<const unsigned char *>(&h)[k];
    msg[kHandleBytes] = ok ? 1 : 0;
    const std::vector<int64_t> all = comm_allgather_i64(ctx, msg);
    bool all_ok = true;
    for (int r = 0; r < size; ++r) all_ok = all_ok && all[(size_t)r * (kHandleBytes + 1) + kHandleBytes] == 1;
    std::vector<void *> mapped((size_t)size, nullptr);
    bool open_ok = all_ok;
    if (all_ok)
      for (int r = 0; r < size; ++r)
        {
          if (r == rank)
            {
              mapped[r] = mine;
              continue;
            }
          cudaIpcMemHandle_t hr;
          for (int k = 0; k < kHandleBytes; ++k) reinterpret_cast<unsigned char *>(&hr)[k] = (unsigned char)all[(size_t)r * (kHandleBytes + 1) + k];
          void *p = nullptr;
          if (cudaIpcOpenMemHandle(&p, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
            {
              cudaGetLastError();
              open_ok = false;
              break;
            }
          mapped[r] = p;
        }
    // every rank has to reach the same verdict
    const std::vector<int64_t> verdict = comm_allgather_i64(ctx, {open_ok ? 1 : 0});
    bool good = true;
    for (int64_t v : verdict) good = good && v == 1;
    if (!good)
      {
        for (int r = 0; r < size; ++r)
          if (r != rank && mapped[r]) cudaIpcCloseMemHandle(mapped[r]);
        if (mine) cudaFree(mine);
        active = false;
        return out;
      }
    owned.push_back(mine);
    for (int r = 0; r < size; ++r)
      if (r != rank) opened.push_back(mapped[r]);
    return mapped;
  }

  void PeerLink::init(Context &ctx)
  {
    rank = ctx.comm ? ctx.comm->rank : 0;
    size = ctx.comm ? ctx.comm->size : 1;
    active = false;
    if (size < 2 || size > kPeerMaxRanks) return;
    if (const char *e = std::getenv("IFEM_PEER")) mask = std::atoi(e) & 3;
    if (mask == 0) return;
    active = true; // tentatively: alloc_shared clears it on any failure
    ll_peers = alloc_shared(ctx, sizeof(unsigned long long) * 2 * kPeerMaxRanks * kPeerWords);
    if (!active) return;
    ll = static_cast<unsigned long long *>(ll_peers[rank]);
    IFEM_CUDA(cudaMalloc(&epoch, sizeof(unsigned int)));
    IFEM_CUDA(cudaMemsetAsync(epoch, 0, sizeof(unsigned int), ctx.stream));
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  PeerDev PeerLink::dev() const
  {
    PeerDev d;
    d.rank = rank;
    d.size = size;
    d.active = active ? 1 : 0;
    if (active)
      {
        d.ll_local = ll;
        for (int r = 0; r < size; ++r) d.ll_remote[r] = static_cast<unsigned long long *>(ll_peers[r]);
        d.epoch = epoch;
      }
    return d;
  }

  PeerLink &peer_link(Context &ctx)
  {
    if (!g_link)
      {
        g_link = std::make_unique<PeerLink>();
        g_link->init(ctx);
      }
    return *g_link;
  }

  void peer_link_reset() { g_link.reset(); }

  ReduceMode reduce_mode(Context &ctx)
  {
    ReduceMode m;
    if (ctx.comm && ctx.comm->size > 1)
      {
        PeerLink &link = peer_link(ctx);
        const bool on = link.active && (link.mask & 1);
        if (on) m.pd = link.dev();
        m.nccl = !on;
        m.adv = on ? 1 : 0;
      }
    return m;
  }
} // namespace ifem

namespace ifem
{
  namespace
  {
    // dst[off + i] = tag + i
    __global__ void selftest_write_kernel(float *dst, int64_t off, int64_t n, float tag)
    {
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[off + i] = tag + (float)(i % 1024);
    }
    __global__ void selftest_check_kernel(const float *buf, int64_t off, int64_t n, float tag, unsigned long long *bad)
    {
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (buf[off + i] != tag + (float)(i % 1024)) atomicAdd(bad, 1ull);
    }
    __global__ void selftest_reduce_kernel(double a, double b, double *partials, unsigned int *counter, double *red, PeerDev pd)
    {
      double acc[2] = {threadIdx.x == 0 && blockIdx.x == 0 ? a : 0.0, threadIdx.x == 1 && blockIdx.x == gridDim.x - 1 ? b : 0.0};
      finish_reduce<2>(acc, partials, counter, red, pd);
    }
  } // namespace

  int64_t peer_selftest(Context &ctx, int rounds)
  {
    PeerLink &link = peer_link(ctx);
    if (!link.active) return -1;
    const int rank = link.rank, size = link.size;
    int64_t bad_total = 0;
    DevBuf<unsigned long long> bad(1);
    bad.zero(ctx.stream);
    for (size_t bytes : {size_t(1) << 20, size_t(5) << 20, size_t(300) << 20})
      {
        const std::vector<void *> bufs = link.alloc_shared(ctx, bytes);
        if (bufs.empty()) return -1;
        const int64_t seg = (int64_t)(bytes / sizeof(float)) / size; // every rank owns one segment of every copy
        for (int r = 0; r < size; ++r)
          selftest_write_kernel<<<64, 256, 0, ctx.stream>>>(static_cast<float *>(bufs[r]), seg * rank, seg, (float)(100 * rank + r));
        IFEM_KERNEL_CHECK();
        IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
        comm_allgather_i64(ctx, {1}); // everybody has written
        for (int r = 0; r < size; ++r)
          selftest_check_kernel<<<64, 256, 0, ctx.stream>>>(static_cast<const float *>(bufs[rank]), seg * r, seg, (float)(100 * r + rank), bad.p);
        IFEM_KERNEL_CHECK();
        bad_total += (int64_t)bad.to_host(ctx.stream)[0];
        bad.zero(ctx.stream);
        comm_allgather_i64(ctx, {1}); // nobody rewrites a buffer that is still being checked
      }
    // ghost push: every rank sends 100 000 "nodes" of width 4 to every other rank with the halo kernels; rank s's message lands
    // at float offset (8 + s) * 2^20 of the receiver's buffer; three rounds with fresh values
    if (size <= kPeerMaxNeighbours + 1)
      {
        const int n_nodes = 100000, w = 4;
        const std::vector<void *> src = link.alloc_shared(ctx, size_t(64) << 20), flg = link.alloc_shared(ctx, sizeof(unsigned int) * kPeerMaxRanks);
        if (src.empty() || flg.empty()) return -1;
        DevBuf<unsigned int> hs(2);
        hs.zero(ctx.stream);
        std::vector<int> pos;
        for (int nb = 0; nb < size - 1; ++nb)
          for (int k = 0; k < n_nodes; ++k) pos.push_back((k * 7) % n_nodes);
        DevBuf<int> d_pos;
        d_pos.upload(pos, ctx.stream);
        PeerHaloDev<float> h;
        h.n_nb = h.n_msg = size - 1;
        h.width = w;
        int k = 0;
        for (int r = 0; r < size; ++r)
          if (r != rank)
            {
              h.send_off[k] = k * n_nodes;
              h.dst[k] = static_cast<float *>(src[r]) + (int64_t)(8 + rank) * (1 << 20);
              h.flag[k] = static_cast<unsigned int *>(flg[r]) + rank;
              h.nb_rank[k] = r;
              ++k;
            }
        h.send_off[h.n_msg] = h.n_msg * n_nodes;
        h.my_flags = static_cast<const unsigned int *>(flg[rank]);
        h.epoch = hs.p;
        h.counter = hs.p + 1;
        float *mine = static_cast<float *>(src[rank]);
        DevBuf<float> got((size_t)n_nodes * w);
        for (int round = 1; round <= 3; ++round)
          {
            selftest_write_kernel<<<64, 256, 0, ctx.stream>>>(mine, 0, (int64_t)n_nodes * w, (float)(1000 * round + rank));
            peer_halo_push_kernel<float><<<32, 256, 0, ctx.stream>>>(h, d_pos.p, mine, nullptr);
            peer_halo_wait_kernel<float><<<1, 32, 0, ctx.stream>>>(h, nullptr);
            IFEM_KERNEL_CHECK();
            IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
            for (int r = 0; r < size; ++r)
              if (r != rank)
                {
                  // node k of rank r's message = node (7 k mod n) of r's vector = tag + ((4 (7k mod n) + c) mod 1024)
                  const std::vector<float> hgot = [&] {
                    IFEM_CUDA(cudaMemcpyAsync(got.p, mine + (int64_t)(8 + r) * (1 << 20), got.n * sizeof(float), cudaMemcpyDeviceToDevice, ctx.stream));
                    return got.to_host(ctx.stream);
                  }();
                  for (int kk = 0; kk < n_nodes; ++kk)
                    for (int c = 0; c < w; ++c)
                      if (hgot[(size_t)kk * w + c] != (float)(1000 * round + r) + (float)(((int64_t)((kk * 7) % n_nodes) * w + c) % 1024)) ++bad_total;
                }
            comm_allgather_i64(ctx, {1});
          }
      }
    // all-reduces: rank r contributes (r + 1) * k and 0.5^k
    DevBuf<double> partials(64 * 2), red(kPeerMaxVals);
    DevBuf<unsigned int> counter(1);
    counter.zero(ctx.stream);
    const PeerDev pd = link.dev();
    for (int k = 1; k <= rounds; ++k)
      {
        selftest_reduce_kernel<<<32, 256, 0, ctx.stream>>>((double)(rank + 1) * k, std::ldexp(1.0, -(k % 40)), partials.p, counter.p, red.p, pd);
        IFEM_KERNEL_CHECK();
        if (k % 97 == 0 || k == rounds)
          {
            const std::vector<double> h = red.to_host(ctx.stream);
            const double want0 = 0.5 * size * (size + 1) * k, want1 = size * std::ldexp(1.0, -(k % 40));
            if (h[0] != want0 || h[1] != want1) ++bad_total;
          }
      }
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
    return bad_total;
  }
} // namespace ifem
