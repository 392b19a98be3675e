// TEST INFRASTRUCTURE - NOT PRODUCT CODE (see cuda_runtime.h in this directory). Stands in for openifem_b200/csrc/comm.cpp
// (NCCL) in the emulated build: the ranks are processes on one machine that exchange through files in a rendezvous directory
// named by the "unique id". Same semantics as the product's communicator: a sum all-reduce of doubles (summed in rank order,
// so every rank gets the same bits), pairwise send/receive, and group start/end inside which all sends are posted before any
// receive is awaited (NCCL's group semantics - the halo exchange relies on it to be deadlock-free).
#include "comm.h"

#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace ifem
{
  namespace
  {
    struct Pending
    {
      int peer;
      void *recv;
      size_t bytes;
      long seq;
    };
    struct State
    {
      std::string dir;
      long ar_seq = 0;
      std::map<int, long> send_seq, recv_seq;
      bool in_group = false;
      std::vector<Pending> pending;
    };
    State &st(Comm &c) { return *static_cast<State *>(c.nccl); }

    void write_file(const std::string &path, const void *data, size_t bytes)
    {
      const std::string tmp = path + ".tmp";
      FILE *f = std::fopen(tmp.c_str(), "wb");
      if (!f) throw std::runtime_error("comm_emul: cannot write " + tmp);
      if (bytes) std::fwrite(data, 1, bytes, f);
      std::fclose(f);
      if (std::rename(tmp.c_str(), path.c_str()) != 0) throw std::runtime_error("comm_emul: cannot rename " + tmp);
    }
    void read_file(const std::string &path, void *data, size_t bytes)
    {
      for (long spins = 0;; ++spins)
        {
          struct stat sb;
          if (stat(path.c_str(), &sb) == 0 && (size_t)sb.st_size == bytes)
            {
              FILE *f = std::fopen(path.c_str(), "rb");
              if (f)
                {
                  const size_t got = bytes ? std::fread(data, 1, bytes, f) : 0;
                  std::fclose(f);
                  if (got == bytes) return;
                }
            }
          if (spins > 6000000) throw std::runtime_error("comm_emul: timed out waiting for " + path);
          usleep(spins < 200 ? 20 : 100);
        }
    }
    void finish(Comm &c, const Pending &p)
    {
      const std::string path = st(c).dir + "/sr_" + std::to_string(p.peer) + "_" + std::to_string(c.rank) + "_" + std::to_string(p.seq);
      read_file(path, p.recv, p.bytes);
      std::remove(path.c_str()); // only this rank reads it
    }
    // as ncclSend / ncclRecv in the product: an empty side of the pair is skipped, so the k-th non-empty send of a rank to a
    // peer meets the k-th non-empty receive of that peer from the rank
    void sendrecv_bytes(Comm &c, int peer, const void *send, size_t n_send, void *recv, size_t n_recv)
    {
      State &s = st(c);
      if (n_send)
        {
          const long ss = s.send_seq[peer]++;
          write_file(s.dir + "/sr_" + std::to_string(c.rank) + "_" + std::to_string(peer) + "_" + std::to_string(ss), send, n_send);
        }
      if (!n_recv) return;
      const Pending p{peer, recv, n_recv, s.recv_seq[peer]++};
      if (s.in_group)
        s.pending.push_back(p);
      else
        finish(c, p);
    }
  } // namespace

  void comm_get_unique_id(unsigned char id[128])
  {
    std::memset(id, 0, 128);
    std::snprintf(reinterpret_cast<char *>(id), 128, "/tmp/ifem_emul_comm_%ld_%ld", (long)getpid(), (long)time(nullptr));
  }

  Comm *comm_create(int rank, int size, const unsigned char id[128])
  {
    auto *c = new Comm;
    c->rank = rank;
    c->size = size;
    auto *s = new State;
    s->dir = reinterpret_cast<const char *>(id);
    mkdir(s->dir.c_str(), 0700);
    c->nccl = s;
    return c;
  }

  void comm_destroy(Comm *c)
  {
    if (!c) return;
    delete static_cast<State *>(c->nccl);
    delete c;
  }

  void comm_allreduce_sum(Comm &c, double *dev, int n, cudaStream_t)
  {
    State &s = st(c);
    const long k = s.ar_seq++;
    auto name = [&](int r, long q) { return s.dir + "/ar_" + std::to_string(r) + "_" + std::to_string(q); };
    // every other rank has read sequence k - 2 of this rank before it wrote its k - 1, which this rank has read
    if (k >= 2) std::remove(name(c.rank, k - 2).c_str());
    write_file(name(c.rank, k), dev, (size_t)n * sizeof(double));
    std::vector<double> sum((size_t)n, 0.0), tmp((size_t)n);
    for (int r = 0; r < c.size; ++r)
      {
        read_file(name(r, k), tmp.data(), (size_t)n * sizeof(double));
        for (int i = 0; i < n; ++i) sum[i] += tmp[i];
      }
    std::memcpy(dev, sum.data(), (size_t)n * sizeof(double));
  }

  void comm_sendrecv(Comm &c, int peer, const double *send, int64_t n_send, double *recv, int64_t n_recv, cudaStream_t)
  {
    sendrecv_bytes(c, peer, send, (size_t)n_send * sizeof(double), recv, (size_t)n_recv * sizeof(double));
  }
  void comm_sendrecv_f32(Comm &c, int peer, const float *send, int64_t n_send, float *recv, int64_t n_recv, cudaStream_t)
  {
    sendrecv_bytes(c, peer, send, (size_t)n_send * sizeof(float), recv, (size_t)n_recv * sizeof(float));
  }
  void comm_group_start(Comm &c) { st(c).in_group = true; }
  void comm_group_end(Comm &c)
  {
    State &s = st(c);
    s.in_group = false;
    for (const Pending &p : s.pending) finish(c, p);
    s.pending.clear();
  }
} // namespace ifem
