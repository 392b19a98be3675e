// CUDA plumbing shared by all translation units: error checking, RAII device
// buffers, the per-process execution context (one device, one compute stream, one
// communicator - mirrors "one thread per MPI rank") and small device
// helpers. sm_100a only; there is no CPU fallback anywhere behind these calls.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace ifem
{
  inline void cuda_check(cudaError_t e, const char *what, const char *file, int line)
  {
    if (e != cudaSuccess)
      {
        char buf[512];
        std::snprintf(buf, sizeof buf, "CUDA error %s (%s) at %s:%d", cudaGetErrorName(e), what, file, line);
        throw std::runtime_error(buf);
      }
  }
#define IFEM_CUDA(x) ::ifem::cuda_check((x), #x, __FILE__, __LINE__)
#define IFEM_KERNEL_CHECK() ::ifem::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__)

  template <typename T>
  struct DevBuf
  {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
      if (this != &o)
        {
          release();
          p = o.p; n = o.n;
          o.p = nullptr; o.n = 0;
        }
      return *this;
    }
    ~DevBuf() { release(); }
    void release()
    {
      if (p) cudaFree(p);
      p = nullptr;
      n = 0;
    }
    void alloc(size_t n_)
    {
      release();
      n = n_;
      if (n) IFEM_CUDA(cudaMalloc(&p, n * sizeof(T)));
    }
    void zero(cudaStream_t s) { if (n) IFEM_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    void upload(const T *h, size_t count, cudaStream_t s)
    {
      if (count > n) throw std::runtime_error("DevBuf::upload overflow");
      if (count) IFEM_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void upload(const std::vector<T> &h, cudaStream_t s)
    {
      if (n < h.size()) alloc(h.size());
      upload(h.data(), h.size(), s);
    }
    void download(T *h, size_t count, cudaStream_t s) const
    {
      if (count > n) throw std::runtime_error("DevBuf::download overflow");
      if (count) IFEM_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
      IFEM_CUDA(cudaStreamSynchronize(s));
    }
    std::vector<T> to_host(cudaStream_t s) const
    {
      std::vector<T> h(n);
      download(h.data(), n, s);
      return h;
    }
  };

  // Communicator interface (rank-local no-op by default; NCCL implementation in
  // comm.cpp). Only the two operations the path needs: scalar
  // all-reduce for Krylov dot products and ghost-DoF halo exchange.
  struct Comm;

  struct Context
  {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    // scratch for reductions
    DevBuf<double> partials; // [n_partials * max_results]
    DevBuf<double> results;  // small device scalars
    double *h_results = nullptr; // pinned
    Comm *comm = nullptr;
    int spmv_variant = 0; // 0 = default kernel; see linalg.cu
    int spmv_l2hint = 0;  // 1: x gathers carry an L2 evict_last policy (IFEM_SPMV_L2HINT)
    int spmv_rpw = 1;     // rows per lane group and CTA chunk (x reuse through L1); IFEM_SPMV_RPW
    int spmv_short = 0;   // kernel shape for short rows / off-diagonal blocks: 10 * lanes per row + unroll (IFEM_SPMV_SHORT)
    long long kernel_launches = 0; // counted by every launcher (bench "gpu_launches")
    Context();
    ~Context();
    Context(const Context &) = delete;
  };

  Context &default_context();

#ifdef __CUDACC__
  __device__ __forceinline__ double warp_sum(double v)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  // One bulk copy global -> shared through the TMA engine (cp.async.bulk, 1-D form) for tables every CTA stages once:
  // thread 0 arms an mbarrier with the byte count and issues the copy, every thread waits on the barrier phase. dst, src and
  // bytes must be multiples of 16; `bar` is an 8-byte shared-memory word of the CTA, used once (phase 0).
  __device__ __forceinline__ void tma_stage_1d(void *smem_dst, const void *gmem_src, unsigned int bytes, unsigned long long *bar)
  {
#ifdef IFEM_EMULATED_DEVICE
    for (unsigned int i = threadIdx.x; i < bytes / 8; i += blockDim.x)
      static_cast<double *>(smem_dst)[i] = static_cast<const double *>(gmem_src)[i];
    (void)bar;
    __syncthreads();
#else
    const unsigned int bar_s = (unsigned int)__cvta_generic_to_shared(bar), dst_s = (unsigned int)__cvta_generic_to_shared(smem_dst);
    if (threadIdx.x == 0)
      {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s), "l"(gmem_src),
                     "r"(bytes), "r"(bar_s)
                     : "memory");
      }
    __syncthreads(); // the barrier is initialised before anybody polls it
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "IFEM_TMA_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
                 "@p bra IFEM_TMA_DONE;\n"
                 "bra IFEM_TMA_WAIT;\n"
                 "IFEM_TMA_DONE:\n"
                 "}" ::"r"(bar_s)
                 : "memory");
#endif
  }
  // start moving the line at p into L2 (no register, no dependency): issued ahead of a read-modify-write whose operands are
  // being computed
  __device__ __forceinline__ void prefetch_l2(const void *p)
  {
#ifdef IFEM_EMULATED_DEVICE
    (void)p;
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
  }
  // streaming (read-once) loads: keep the matrix stream from evicting x out of L2
  __device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
  __device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }
#endif
} // namespace ifem
