#include "comm.h"

#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "device.cuh"

// Minimal run-time binding of the NCCL entry points the path uses. The library is
// the torch-bundled libnccl.so.2 (already mapped into the process when the
// launcher imported torch); $IFEM_NCCL_LIB overrides the name.
namespace ifem
{
  namespace
  {
    struct Uid
    {
      char internal[128];
    };
    struct NcclApi
    {
      void *h = nullptr;
      int (*GetUniqueId)(void *) = nullptr;
      int (*CommInitRank)(void **, int, /*ncclUniqueId by value*/ Uid, int) = nullptr;
      int (*CommDestroy)(void *) = nullptr;
      int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
      int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
      int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
      int (*GroupStart)() = nullptr;
      int (*GroupEnd)() = nullptr;
      const char *(*GetErrorString)(int) = nullptr;
    };
    constexpr int kNcclFloat64 = 8, kNcclFloat32 = 7, kNcclSum = 0;

    NcclApi &api()
    {
      static NcclApi a;
      if (a.h) return a;
      const char *name = std::getenv("IFEM_NCCL_LIB");
      a.h = dlopen(name ? name : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!a.h) throw std::runtime_error(std::string("openifem_b200: cannot load NCCL: ") + dlerror());
      auto sym = [&](const char *s) {
        void *p = dlsym(a.h, s);
        if (!p) throw std::runtime_error(std::string("openifem_b200: NCCL symbol missing: ") + s);
        return p;
      };
      a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
      a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
      a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
      a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
      a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
      a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
      a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
      a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
      a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
      return a;
    }

    void check(int rc, const char *what)
    {
      if (rc != 0) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + api().GetErrorString(rc));
    }
  } // namespace

  void comm_get_unique_id(unsigned char id[128])
  {
    Uid u;
    check(api().GetUniqueId(&u), "ncclGetUniqueId");
    std::memcpy(id, u.internal, 128);
  }

  Comm *comm_create(int rank, int size, const unsigned char id[128])
  {
    Comm *c = new Comm;
    c->rank = rank;
    c->size = size;
    if (size > 1)
      {
        Uid u;
        std::memcpy(u.internal, id, 128);
        check(api().CommInitRank(&c->nccl, size, u, rank), "ncclCommInitRank");
      }
    return c;
  }

  void comm_destroy(Comm *c)
  {
    if (!c) return;
    if (c->nccl) api().CommDestroy(c->nccl);
    delete c;
  }

  void comm_allreduce_sum(Comm &c, double *dev, int n, cudaStream_t s)
  {
    if (c.size <= 1) return;
    check(api().AllReduce(dev, dev, (size_t)n, kNcclFloat64, kNcclSum, c.nccl, s), "ncclAllReduce");
  }

  void comm_sendrecv(Comm &c, int peer, const double *send, int64_t n_send, double *recv, int64_t n_recv, cudaStream_t s)
  {
    if (n_send) check(api().Send(send, (size_t)n_send, kNcclFloat64, peer, c.nccl, s), "ncclSend");
    if (n_recv) check(api().Recv(recv, (size_t)n_recv, kNcclFloat64, peer, c.nccl, s), "ncclRecv");
  }
  void comm_sendrecv_f32(Comm &c, int peer, const float *send, int64_t n_send, float *recv, int64_t n_recv, cudaStream_t s)
  {
    if (n_send) check(api().Send(send, (size_t)n_send, kNcclFloat32, peer, c.nccl, s), "ncclSend");
    if (n_recv) check(api().Recv(recv, (size_t)n_recv, kNcclFloat32, peer, c.nccl, s), "ncclRecv");
  }
  void comm_group_start(Comm &) { check(api().GroupStart(), "ncclGroupStart"); }
  void comm_group_end(Comm &) { check(api().GroupEnd(), "ncclGroupEnd"); }
} // namespace ifem
