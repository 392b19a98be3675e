#include "halo.h"

#include "comm.h"

namespace ifem
{
  namespace
  {
    __global__ void halo_pack_kernel(int n, int bs, const int *__restrict__ idx, const double *__restrict__ v, double *__restrict__ buf)
    {
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= n * bs) return;
      const int k = t / bs, c = t % bs;
      buf[t] = v[(int64_t)idx[k] * bs + c];
    }
  } // namespace

  void Halo::init(Context &ctx, const NodePartition &np, int block_size)
  {
    bs = block_size;
    n_owned = np.n_owned;
    n_local = np.n_local;
    neighbours = np.neighbours;
    recv_off = np.recv_offset;
    recv_cnt = np.recv_count;
    std::vector<int> all;
    send_off.clear();
    send_cnt.clear();
    for (const auto &l : np.send_local)
      {
        send_off.push_back((int)all.size());
        send_cnt.push_back((int)l.size());
        all.insert(all.end(), l.begin(), l.end());
      }
    n_send_total = (int)all.size();
    if (n_send_total)
      {
        d_send_idx.upload(all, ctx.stream);
        d_send_buf.alloc((size_t)n_send_total * bs);
        IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
      }
  }

  void Halo::update(Context &ctx, double *v)
  {
    if (neighbours.empty()) return;
    if (!ctx.comm) throw std::runtime_error("Halo::update: no communicator");
    if (n_send_total)
      {
        const int total = n_send_total * bs;
        halo_pack_kernel<<<(total + 255) / 256, 256, 0, ctx.stream>>>(n_send_total, bs, d_send_idx.p, v, d_send_buf.p);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
    comm_group_start(*ctx.comm);
    for (size_t k = 0; k < neighbours.size(); ++k)
      comm_sendrecv(*ctx.comm, neighbours[k], d_send_buf.p + (size_t)send_off[k] * bs, (int64_t)send_cnt[k] * bs,
                    v + (size_t)recv_off[k] * bs, (int64_t)recv_cnt[k] * bs, ctx.stream);
    comm_group_end(*ctx.comm);
  }
} // namespace ifem
