#include "peer.h"

#include <cstdlib>
#include <cstring>
#include <memory>

#include "comm.h"

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
