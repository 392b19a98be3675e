#include <algorithm>
#include <cmath>

#include "comm.h"
#include "krylov.h"
#include "peer_dev.cuh"

namespace ifem
{
  namespace
  {
    constexpr int kT = 256;

    struct Seg
    {
      int64_t len0, shift, total; // owned entry k lives at k (k < len0) or k + shift
      __host__ __device__ int64_t operator()(int64_t k) const { return k < len0 ? k : k + shift; }
    };

    struct State
    {
      double rr, alpha, beta, tol2;
      int its, max_it, done, converged;
    };
    enum Stage { kInit, kDot, kXR };

    __device__ __forceinline__ void advance(int stage, State *st, const double *red)
    {
      if (stage == kInit)
        {
          st->rr = red[0];
          if (!(red[0] > st->tol2))
            {
              st->done = 1;
              st->converged = red[0] <= st->tol2 ? 1 : 0;
            }
          return;
        }
      if (st->done) return;
      if (stage == kDot)
        {
          if (!(red[0] > 0.0) || !isfinite(red[0]))
            st->done = 1;
          else
            st->alpha = st->rr / red[0];
        }
      else
        {
          const double rr_new = red[0];
          st->beta = rr_new / st->rr;
          st->rr = rr_new;
          st->its += 1;
          if (!(rr_new > st->tol2) || !isfinite(rr_new))
            {
              st->done = 1;
              st->converged = rr_new <= st->tol2 ? 1 : 0;
            }
          else if (st->its >= st->max_it)
            st->done = 1;
        }
    }

    __global__ void advance_kernel(int stage, State *st, const double *red) { advance(stage, st, red); }

    __global__ void begin_kernel(State *st, double tol2, int max_it)
    {
      st->rr = st->alpha = st->beta = 0.0;
      st->tol2 = tol2;
      st->its = 0;
      st->max_it = max_it;
      st->done = st->converged = 0;
    }

    // r = p = b, x = 0; |r|^2
    __global__ void __launch_bounds__(kT)
    init_kernel(Seg sg, const double *__restrict__ b, double *__restrict__ r, double *__restrict__ p, double *__restrict__ x, State *st,
                double *__restrict__ partials, unsigned int *__restrict__ counter, double *__restrict__ red, PeerDev pd, int adv)
    {
      double acc[1] = {0.0};
      for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < sg.total; k += (int64_t)gridDim.x * blockDim.x)
        {
          const int64_t i = sg(k);
          const double v = b[i];
          r[i] = v;
          p[i] = v;
          x[i] = 0.0;
          acc[0] = fma(v, v, acc[0]);
        }
      if (finish_reduce<1>(acc, partials, counter, red, pd) && adv && threadIdx.x == 0) advance(kInit, st, acc);
    }

    __global__ void __launch_bounds__(kT)
    dot_kernel(Seg sg, const double *__restrict__ p, const double *__restrict__ ap, State *st, double *__restrict__ partials,
               unsigned int *__restrict__ counter, double *__restrict__ red, PeerDev pd, int adv)
    {
      if (st->done) return;
      double acc[1] = {0.0};
      for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < sg.total; k += (int64_t)gridDim.x * blockDim.x)
        {
          const int64_t i = sg(k);
          acc[0] = fma(p[i], ap[i], acc[0]);
        }
      if (finish_reduce<1>(acc, partials, counter, red, pd) && adv && threadIdx.x == 0) advance(kDot, st, acc);
    }

    __global__ void __launch_bounds__(kT)
    xr_kernel(Seg sg, State *st, double *__restrict__ x, const double *__restrict__ p, double *__restrict__ r, const double *__restrict__ ap,
              double *__restrict__ partials, unsigned int *__restrict__ counter, double *__restrict__ red, PeerDev pd, int adv)
    {
      if (st->done) return;
      const double alpha = st->alpha;
      double acc[1] = {0.0};
      for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < sg.total; k += (int64_t)gridDim.x * blockDim.x)
        {
          const int64_t i = sg(k);
          x[i] = fma(alpha, p[i], x[i]);
          const double rc = fma(-alpha, ap[i], r[i]);
          r[i] = rc;
          acc[0] = fma(rc, rc, acc[0]);
        }
      if (finish_reduce<1>(acc, partials, counter, red, pd) && adv && threadIdx.x == 0) advance(kXR, st, acc);
    }

    __global__ void __launch_bounds__(kT) p_kernel(Seg sg, const State *__restrict__ st, const double *__restrict__ r, double *__restrict__ p)
    {
      if (st->done) return;
      const double beta = st->beta;
      for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < sg.total; k += (int64_t)gridDim.x * blockDim.x)
        {
          const int64_t i = sg(k);
          p[i] = fma(beta, p[i], r[i]);
        }
    }
  } // namespace

  DeviceCG64::~DeviceCG64()
  {
    if (h_state) cudaFreeHost(h_state);
  }

  SolveResult DeviceCG64::solve(Context &ctx, const VecSpace &n, const LinOp &A, const double *b, double *x, double tol_abs, int max_it)
  {
    SolveResult out;
    const int64_t owned = n.n_owned();
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((owned + kT * 4 - 1) / (kT * 4), (int64_t)ctx.sm_count * 4));
    if ((int64_t)r.n < n.n_alloc)
      {
        r.alloc(n.n_alloc);
        p.alloc(n.n_alloc);
        ap.alloc(n.n_alloc);
        p.zero(ctx.stream); // ghost entries are read by A before the first halo fills them
      }
    if ((int)partials.n < grid) partials.alloc(grid);
    if (!red.p)
      {
        red.alloc(kPeerMaxVals);
        red.zero(ctx.stream);
        counter.alloc(1);
        counter.zero(ctx.stream);
        state.alloc((sizeof(State) + sizeof(int) - 1) / sizeof(int));
        state.zero(ctx.stream);
        IFEM_CUDA(cudaMallocHost(&h_state, sizeof(State)));
      }
    State *st = reinterpret_cast<State *>(state.p);
    const State *h = static_cast<const State *>(h_state);
    const Seg sg{n.len0, n.off1 - n.len0, owned};
    const ReduceMode m = reduce_mode(ctx);
    auto launched = [&](int k = 1) {
      IFEM_KERNEL_CHECK();
      ctx.kernel_launches += k;
    };
    auto after = [&](int stage) {
      if (!m.nccl) return;
      comm_allreduce_sum(*ctx.comm, red.p, 1, ctx.stream);
      advance_kernel<<<1, 1, 0, ctx.stream>>>(stage, st, red.p);
      launched();
    };
    begin_kernel<<<1, 1, 0, ctx.stream>>>(st, tol_abs * tol_abs, max_it);
    launched();
    init_kernel<<<grid, kT, 0, ctx.stream>>>(sg, b, r.p, p.p, x, st, partials.p, counter.p, red.p, m.pd, m.adv);
    launched();
    after(kInit);
    int enqueued = 0;
    int chunk = std::max(4, last_its); // the same system is solved again and again: expect the same count
    while (true)
      {
        chunk = std::max(1, std::min(chunk, max_it - enqueued));
        for (int k = 0; k < chunk; ++k)
          {
            A(p.p, ap.p);
            dot_kernel<<<grid, kT, 0, ctx.stream>>>(sg, p.p, ap.p, st, partials.p, counter.p, red.p, m.pd, m.adv);
            launched();
            after(kDot);
            xr_kernel<<<grid, kT, 0, ctx.stream>>>(sg, st, x, p.p, r.p, ap.p, partials.p, counter.p, red.p, m.pd, m.adv);
            launched();
            after(kXR);
            p_kernel<<<grid, kT, 0, ctx.stream>>>(sg, st, r.p, p.p);
            launched();
          }
        enqueued += chunk;
        IFEM_CUDA(cudaMemcpyAsync(h_state, state.p, sizeof(State), cudaMemcpyDeviceToHost, ctx.stream));
        IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
        if (h->done || enqueued >= max_it) break;
        chunk = 4;
      }
    out.iterations = h->its;
    out.residual = std::sqrt(std::max(0.0, h->rr));
    out.converged = h->converged != 0;
    last_its = h->its;
    return out;
  }
} // namespace ifem
