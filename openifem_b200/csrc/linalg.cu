#include "linalg.h"

#include <algorithm>
#include <cstdlib>

#include "comm.h"

namespace ifem
{
  // ---------------------------------------------------------------------------
  // Context
  // ---------------------------------------------------------------------------
  static constexpr int kMaxPartials = 4096;
  static constexpr int kOrthoMaxBasis = 64, kOrthoMaxGrid = 2048; // fused Gram-Schmidt (orthogonalise_cgs2)

  Context::Context()
  {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      throw std::runtime_error("openifem_b200: no CUDA device available - this library has no CPU fallback");
    IFEM_CUDA(cudaGetDevice(&device));
    cudaDeviceProp prop;
    IFEM_CUDA(cudaGetDeviceProperties(&prop, device));
    sm_count = prop.multiProcessorCount;
    IFEM_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    partials.alloc(kOrthoMaxBasis * kOrthoMaxGrid > kMaxPartials ? kOrthoMaxBasis * kOrthoMaxGrid : kMaxPartials);
    results.alloc(256);
    IFEM_CUDA(cudaMallocHost(&h_results, 256 * sizeof(double)));
    if (const char *v = std::getenv("IFEM_SPMV_VARIANT")) spmv_variant = std::atoi(v);
    if (const char *v = std::getenv("IFEM_SPMV_RPW")) spmv_rpw = std::atoi(v);
    if (const char *v = std::getenv("IFEM_SPMV_SHORT")) spmv_short = std::atoi(v);
    if (const char *v = std::getenv("IFEM_SPMV_L2HINT")) spmv_l2hint = std::atoi(v);
  }

  Context::~Context()
  {
    if (h_results) cudaFreeHost(h_results);
    if (stream) cudaStreamDestroy(stream);
  }

  Context &default_context()
  {
    static Context ctx;
    return ctx;
  }

  // ---------------------------------------------------------------------------
  // Bcsr
  // ---------------------------------------------------------------------------
  void Bcsr::init(const Pattern &P, int R_, int C_, cudaStream_t s)
  {
    R = R_;
    C = C_;
    n_brows = P.n_rows;
    n_bcols = P.n_cols;
    n_blocks = (int64_t)P.col.size();
    rowptr.alloc(P.rowptr.size());
    rowptr.upload(P.rowptr.data(), P.rowptr.size(), s);
    col.alloc(P.col.size());
    col.upload(P.col.data(), P.col.size(), s);
    val.alloc((size_t)n_blocks * R * C);
    val.zero(s);
    IFEM_CUDA(cudaStreamSynchronize(s));
    const double avg = n_brows ? double(n_blocks) / n_brows : 0.0;
    tpr = avg >= 40 ? 32 : avg >= 20 ? 16 : avg >= 10 ? 8 : 4;
  }

  void Bcsr::to_host_csr(cudaStream_t s, std::vector<int64_t> &rp, std::vector<int> &ci, std::vector<double> &v) const
  {
    const std::vector<int64_t> brp = rowptr.to_host(s);
    const std::vector<int> bci = col.to_host(s);
    const std::vector<double> bv = val.to_host(s);
    rp.assign((size_t)n_brows * R + 1, 0);
    ci.resize((size_t)n_blocks * R * C);
    v.resize((size_t)n_blocks * R * C);
    int64_t pos = 0;
    for (int i = 0; i < n_brows; ++i)
      {
        const int64_t base = brp[i], nb = brp[i + 1] - base;
        for (int r = 0; r < R; ++r)
          {
            rp[(size_t)i * R + r] = pos;
            for (int64_t j = 0; j < nb; ++j)
              for (int c = 0; c < C; ++c)
                {
                  ci[pos] = bci[base + j] * C + c;
                  v[pos] = bv[base * R * C + (int64_t)(r * C + c) * nb + j];
                  ++pos;
                }
          }
      }
    rp[(size_t)n_brows * R] = pos;
  }

  // ---------------------------------------------------------------------------
  // SpMV: TPR lanes cooperate on one block row; every plane of the row is read
  // with unit stride across the lanes (coalesced, read-once -> streaming loads),
  // x is gathered through L1/L2, the R partial sums are shuffled down.
  // ---------------------------------------------------------------------------
  __device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }

  // L2 residency hints: the matrix is read exactly once per product (evict_first), the x window of a sweep is
  // re-read by the next rows / planes (evict_last), so that the 40-80 GB matrix stream does not push x out of L2
  __device__ __forceinline__ uint64_t l2_policy_evict_last()
  {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
  }
  __device__ __forceinline__ double ld_x_hint(const double *p, uint64_t pol)
  {
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
  }

  // one block row per lane group
  template <int R, int C, int TPR, typename VT, int UNROLL = 2, int MINB = 1, bool HINT = false>
  __global__ void __launch_bounds__(256, MINB)
  bcsr_spmv_row_kernel(int n_brows, const int64_t *__restrict__ rowptr, const int *__restrict__ col,
                       const VT *__restrict__ val, const double *__restrict__ x, double *__restrict__ y, int accumulate)
  {
    uint64_t pol = 0;
    if (HINT) pol = l2_policy_evict_last();
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = gt / TPR;
    const int lane = (int)(gt % TPR);
    double acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0;
    if (row < n_brows)
      {
        const int64_t base = rowptr[row];
        const int nb = (int)(rowptr[row + 1] - base);
        const VT *v = val + base * (R * C);
        const int *ci = col + base;
#pragma unroll UNROLL
        for (int j = lane; j < nb; j += TPR)
          {
            const int c0 = ld_stream(ci + j);
            double xv[C];
#pragma unroll
            for (int c = 0; c < C; ++c) xv[c] = HINT ? ld_x_hint(x + (int64_t)c0 * C + c, pol) : __ldg(x + (int64_t)c0 * C + c);
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
              for (int c = 0; c < C; ++c) acc[r] = fma((double)ld_stream(v + (int64_t)(r * C + c) * nb + j), xv[c], acc[r]);
          }
      }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int o = TPR / 2; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
    if (row < n_brows && lane == 0)
      {
#pragma unroll
        for (int r = 0; r < R; ++r)
          {
            double *yp = y + row * R + r;
            *yp = accumulate ? (*yp + acc[r]) : acc[r];
          }
      }
  }

  // A CTA of 256 threads = G = 256 / TPR lane groups; group g handles rows row0 + k * G + g, k < rpw, of the
  // CTA's chunk of G * rpw consecutive rows. Consecutive rows share most of their column nodes, so walking a
  // chunk inside one CTA turns the x gathers of later rows into L1 hits (rpw = 1: one row per group).
  template <int R, int C, int TPR, typename VT, int UNROLL = 2, int MINB = 1>
  __global__ void __launch_bounds__(256, MINB)
  bcsr_spmv_kernel(int n_brows, const int64_t *__restrict__ rowptr, const int *__restrict__ col,
                   const VT *__restrict__ val, const double *__restrict__ x, double *__restrict__ y, int accumulate, int rpw)
  {
    constexpr int G = 256 / TPR;
    const int g = threadIdx.x / TPR, lane = threadIdx.x % TPR;
    const int64_t row0 = (int64_t)blockIdx.x * G * rpw;
    for (int k = 0; k < rpw; ++k)
      {
        const int64_t row = row0 + (int64_t)k * G + g;
        double acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.0;
        if (row < n_brows)
          {
            const int64_t base = rowptr[row];
            const int nb = (int)(rowptr[row + 1] - base);
            const VT *v = val + base * (R * C);
            const int *ci = col + base;
#pragma unroll UNROLL
            for (int j = lane; j < nb; j += TPR)
              {
                const int c0 = ld_stream(ci + j);
                double xv[C];
#pragma unroll
                for (int c = 0; c < C; ++c) xv[c] = __ldg(x + (int64_t)c0 * C + c);
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                  for (int c = 0; c < C; ++c) acc[r] = fma((double)ld_stream(v + (int64_t)(r * C + c) * nb + j), xv[c], acc[r]);
              }
          }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int o = TPR / 2; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
        if (row < n_brows && lane == 0)
          {
#pragma unroll
            for (int r = 0; r < R; ++r)
              {
                double *yp = y + row * R + r;
                *yp = accumulate ? (*yp + acc[r]) : acc[r];
              }
          }
      }
  }

  template <int R, int C, typename VT>
  static void spmv_launch(Context &ctx, const Bcsr &A, const VT *val, const double *x, double *y, bool accumulate)
  {
    const int n_rows = A.n_brows_spmv >= 0 ? A.n_brows_spmv : A.n_brows;
    if (n_rows == 0) return;
    const int threads = 256;
    auto launch = [&](auto tpr_tag) {
      constexpr int TPR = decltype(tpr_tag)::value;
      const int64_t total = (int64_t)n_rows * TPR;
      const int64_t blocks = (total + threads - 1) / threads;
      if (R == 3 && C == 3 && TPR == 32)
        {
          // velocity block: tuned variants, key = 100 * lanes per row + 10 * unroll + min CTAs per SM
          // (IFEM_SPMV_VARIANT overrides; default chosen from the round-1 sweep, profiles/r01_spmv_sweep.txt)
          const int key = ctx.spmv_variant ? ctx.spmv_variant : (sizeof(VT) == 8 ? 3214 : 1614);
#define IFEM_SPMV_V(T, U, M)                                                                                                    \
  case T * 100 + U * 10 + M:                                                                                                    \
    {                                                                                                                           \
      const int rpw = std::max(1, ctx.spmv_rpw);                                                                                \
      const int64_t per_cta = (int64_t)(threads / T) * rpw;                                                                     \
      const int64_t nblk = (n_rows + per_cta - 1) / per_cta;                                                                    \
      if (rpw == 1 && ctx.spmv_l2hint && U == 1 && M == 4)                                                                      \
        bcsr_spmv_row_kernel<R, C, T, VT, 1, 4, true><<<(unsigned)nblk, threads, 0, ctx.stream>>>(n_rows, A.rowptr.p, A.col.p, val, \
                                                                                                  x, y, accumulate ? 1 : 0);   \
      else if (rpw == 1)                                                                                                        \
        bcsr_spmv_row_kernel<R, C, T, VT, U, M><<<(unsigned)nblk, threads, 0, ctx.stream>>>(n_rows, A.rowptr.p, A.col.p, val, x, \
                                                                                            y, accumulate ? 1 : 0);             \
      else                                                                                                                      \
        bcsr_spmv_kernel<R, C, T, VT, U, M><<<(unsigned)nblk, threads, 0, ctx.stream>>>(n_rows, A.rowptr.p, A.col.p, val, x, y, \
                                                                                        accumulate ? 1 : 0, rpw);               \
      return;                                                                                                                   \
    }
          switch (key)
            {
              IFEM_SPMV_V(32, 1, 4) IFEM_SPMV_V(32, 1, 6) IFEM_SPMV_V(32, 1, 8) IFEM_SPMV_V(32, 2, 4) IFEM_SPMV_V(32, 2, 6)
              IFEM_SPMV_V(16, 1, 4) IFEM_SPMV_V(16, 1, 6) IFEM_SPMV_V(16, 1, 8) IFEM_SPMV_V(16, 2, 4) IFEM_SPMV_V(16, 2, 6)
              IFEM_SPMV_V(8, 1, 4) IFEM_SPMV_V(8, 1, 6) IFEM_SPMV_V(8, 1, 8) IFEM_SPMV_V(8, 2, 4) IFEM_SPMV_V(8, 2, 6) IFEM_SPMV_V(8, 4, 4)
              IFEM_SPMV_V(4, 1, 4) IFEM_SPMV_V(4, 1, 6) IFEM_SPMV_V(4, 2, 4) IFEM_SPMV_V(4, 4, 4)
            default: break;
            }
#undef IFEM_SPMV_V
        }
      bcsr_spmv_kernel<R, C, TPR, VT><<<(unsigned)blocks, threads, 0, ctx.stream>>>(n_rows, A.rowptr.p, A.col.p, val, x, y,
                                                                                accumulate ? 1 : 0, 1);
    };
    // short rows (Q1 blocks: 9 / 27 block columns) and the off-diagonal shapes: lanes per row x unroll, 4 CTAs per SM; key =
    // 10 * lanes per row + unroll (ifem_set_spmv_short_variant / IFEM_SPMV_SHORT; 0 = the default kernel above)
    if (!(R == 3 && C == 3 && A.tpr == 32) && (A.tpr < 32 || ctx.spmv_short))
      {
        auto short_launch = [&](auto tpr_tag, auto unroll_tag) {
          constexpr int T = decltype(tpr_tag)::value, U = decltype(unroll_tag)::value;
          const int64_t nblk = ((int64_t)n_rows * T + threads - 1) / threads;
          bcsr_spmv_row_kernel<R, C, T, VT, U, 4><<<(unsigned)nblk, threads, 0, ctx.stream>>>(n_rows, A.rowptr.p, A.col.p, val, x, y,
                                                                                              accumulate ? 1 : 0);
        };
        using I = std::integral_constant<int, 0>;
        (void)sizeof(I);
        bool done = true;
        // default 161: 16 lanes per row, unroll 1 - 4 767 GB/s on the four Q1 blocks of the config-5 system against 3 198 GB/s for the
        // kernel above (profiles/r02_spmv_short_sweep.txt); -1 selects the kernel above
        switch (ctx.spmv_short ? ctx.spmv_short : 161)
          {
          case 41: short_launch(std::integral_constant<int, 4>(), std::integral_constant<int, 1>()); break;
          case 44: short_launch(std::integral_constant<int, 4>(), std::integral_constant<int, 4>()); break;
          case 48: short_launch(std::integral_constant<int, 4>(), std::integral_constant<int, 8>()); break;
          case 81: short_launch(std::integral_constant<int, 8>(), std::integral_constant<int, 1>()); break;
          case 84: short_launch(std::integral_constant<int, 8>(), std::integral_constant<int, 4>()); break;
          case 161: short_launch(std::integral_constant<int, 16>(), std::integral_constant<int, 1>()); break;
          case 162: short_launch(std::integral_constant<int, 16>(), std::integral_constant<int, 2>()); break;
          case 321: short_launch(std::integral_constant<int, 32>(), std::integral_constant<int, 1>()); break;
          default: done = false; break;
          }
        if (done)
          {
            IFEM_KERNEL_CHECK();
            ctx.kernel_launches++;
            return;
          }
      }
    switch (A.tpr)
      {
      case 32: launch(std::integral_constant<int, 32>()); break;
      case 16: launch(std::integral_constant<int, 16>()); break;
      case 8: launch(std::integral_constant<int, 8>()); break;
      default: launch(std::integral_constant<int, 4>()); break;
      }
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }

  void spmv(Context &ctx, const Bcsr &A, const double *x, double *y, bool accumulate)
  {
    const int key = A.R * 10 + A.C;
    switch (key)
      {
      case 11: spmv_launch<1, 1>(ctx, A, A.val.p, x, y, accumulate); break;
      case 22: spmv_launch<2, 2>(ctx, A, A.val.p, x, y, accumulate); break;
      case 21: spmv_launch<2, 1>(ctx, A, A.val.p, x, y, accumulate); break;
      case 12: spmv_launch<1, 2>(ctx, A, A.val.p, x, y, accumulate); break;
      case 33: spmv_launch<3, 3>(ctx, A, A.val.p, x, y, accumulate); break;
      case 31: spmv_launch<3, 1>(ctx, A, A.val.p, x, y, accumulate); break;
      case 13: spmv_launch<1, 3>(ctx, A, A.val.p, x, y, accumulate); break;
      default: throw std::runtime_error("spmv: unsupported block shape");
      }
  }

  // ---------------------------------------------------------------------------
  // fp32-streamed SpMV (preconditioner-only): TPR lanes per block row, each lane owns 4 consecutive blocks
  // per step and reads every plane with one LDG.128; x, products and sums stay fp64.
  // ---------------------------------------------------------------------------
  template <int R, int C, int TPR, int MINB>
  __global__ void __launch_bounds__(256, MINB)
  bcsr_spmv_f32x4_kernel(int n_brows, const int64_t *__restrict__ rowptr32, const int *__restrict__ col32,
                         const float *__restrict__ val32, const double *__restrict__ x, double *__restrict__ y, int rpw)
  {
    constexpr int G = 256 / TPR;
    const int g = threadIdx.x / TPR, lane = threadIdx.x % TPR;
    const int64_t row0 = (int64_t)blockIdx.x * G * rpw;
    for (int kk = 0; kk < rpw; ++kk)
    {
    const int64_t row = row0 + (int64_t)kk * G + g;
    double acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0;
    if (row < n_brows)
      {
        const int64_t base = rowptr32[row];
        const int nbp = (int)(rowptr32[row + 1] - base); // multiple of 4
        const float *v = val32 + base * (R * C);
        const int *ci = col32 + base;
        for (int j = 4 * lane; j < nbp; j += 4 * TPR)
          {
            const int4 c4 = __ldcs(reinterpret_cast<const int4 *>(ci + j));
            const int cc[4] = {c4.x, c4.y, c4.z, c4.w};
            double xv[4][C];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int c = 0; c < C; ++c) xv[k][c] = __ldg(x + (int64_t)cc[k] * C + c);
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
              for (int c = 0; c < C; ++c)
                {
                  const float4 a4 = __ldcs(reinterpret_cast<const float4 *>(v + (int64_t)(r * C + c) * nbp + j));
                  acc[r] = fma((double)a4.x, xv[0][c], acc[r]);
                  acc[r] = fma((double)a4.y, xv[1][c], acc[r]);
                  acc[r] = fma((double)a4.z, xv[2][c], acc[r]);
                  acc[r] = fma((double)a4.w, xv[3][c], acc[r]);
                }
          }
      }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int o = TPR / 2; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
    if (row < n_brows && lane == 0)
      {
#pragma unroll
        for (int r = 0; r < R; ++r) y[row * R + r] = acc[r];
      }
    }
  }

  void spmv_fp32(Context &ctx, const Bcsr &A, const double *x, double *y)
  {
    if (!A.val32.p) throw std::runtime_error("spmv_fp32: no fp32 copy (call make_fp32_copy)");
    const int n_rows = A.n_brows_spmv >= 0 ? A.n_brows_spmv : A.n_brows;
    if (!n_rows) return;
    const int key = A.R * 10 + A.C;
    // lanes per row / min CTAs per SM: IFEM_SPMV32_VARIANT = 10 * lanes + minb (default 8 lanes, 4 CTAs)
    // IFEM_SPMV32_VARIANT: 0 (default) = one block per lane, 16 lanes per row (fastest in the round-1 sweeps:
    // 10.3 ms at config 3); 10 * lanes + min CTAs = the 4-blocks-per-lane LDG.128 kernel
    static const int variant = [] {
      const char *v = std::getenv("IFEM_SPMV32_VARIANT");
      return v ? std::atoi(v) : 0;
    }();
    if (variant == 0 && (key == 33 || key == 22))
      {
        const int64_t nblk = ((int64_t)n_rows * 16 + 255) / 256;
        if (key == 33 && ctx.spmv_l2hint)
          bcsr_spmv_row_kernel<3, 3, 16, float, 1, 4, true><<<(unsigned)nblk, 256, 0, ctx.stream>>>(n_rows, A.rowptr32.p, A.col32.p, A.val32.p, x, y, 0);
        else if (key == 33)
          bcsr_spmv_row_kernel<3, 3, 16, float, 1, 4><<<(unsigned)nblk, 256, 0, ctx.stream>>>(n_rows, A.rowptr32.p, A.col32.p, A.val32.p, x, y, 0);
        else
          bcsr_spmv_row_kernel<2, 2, 16, float, 1, 4><<<(unsigned)nblk, 256, 0, ctx.stream>>>(n_rows, A.rowptr32.p, A.col32.p, A.val32.p, x, y, 0);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
        return;
      }
    if (key == 31 || key == 13 || key == 11)
      {
        // the off-diagonal and pressure blocks of an equal-order system (short rows): the 16-lane row kernel that is the fp64 default
        // for these shapes (profiles/r02_spmv_short_sweep.txt), reading the padded fp32 copy
        const int64_t nblk = ((int64_t)n_rows * 16 + 255) / 256;
        if (key == 31)
          bcsr_spmv_row_kernel<3, 1, 16, float, 1, 4><<<(unsigned)nblk, 256, 0, ctx.stream>>>(n_rows, A.rowptr32.p, A.col32.p, A.val32.p, x, y, 0);
        else if (key == 13)
          bcsr_spmv_row_kernel<1, 3, 16, float, 1, 4><<<(unsigned)nblk, 256, 0, ctx.stream>>>(n_rows, A.rowptr32.p, A.col32.p, A.val32.p, x, y, 0);
        else
          bcsr_spmv_row_kernel<1, 1, 16, float, 1, 4><<<(unsigned)nblk, 256, 0, ctx.stream>>>(n_rows, A.rowptr32.p, A.col32.p, A.val32.p, x, y, 0);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
        return;
      }
    if (key == 21 || key == 12)
      {
        const int64_t nblk = ((int64_t)n_rows * 16 + 255) / 256;
        if (key == 21)
          bcsr_spmv_row_kernel<2, 1, 16, float, 1, 4><<<(unsigned)nblk, 256, 0, ctx.stream>>>(n_rows, A.rowptr32.p, A.col32.p, A.val32.p, x, y, 0);
        else
          bcsr_spmv_row_kernel<1, 2, 16, float, 1, 4><<<(unsigned)nblk, 256, 0, ctx.stream>>>(n_rows, A.rowptr32.p, A.col32.p, A.val32.p, x, y, 0);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
        return;
      }
    auto launch = [&](auto r_tag, auto tpr_tag, auto m_tag) {
      constexpr int RR = decltype(r_tag)::value, T = decltype(tpr_tag)::value, M = decltype(m_tag)::value;
      const int rpw = std::max(1, ctx.spmv_rpw);
      const int64_t per_cta = (int64_t)(256 / T) * rpw;
      const int64_t nblk = (n_rows + per_cta - 1) / per_cta;
      bcsr_spmv_f32x4_kernel<RR, RR, T, M><<<(unsigned)nblk, 256, 0, ctx.stream>>>(n_rows, A.rowptr32.p, A.col32.p, A.val32.p, x, y, rpw);
    };
    using I2 = std::integral_constant<int, 2>;
    using I3 = std::integral_constant<int, 3>;
    using I4 = std::integral_constant<int, 4>;
    using I8 = std::integral_constant<int, 8>;
    using I16 = std::integral_constant<int, 16>;
    if (key == 33)
      switch (variant)
        {
        case 42: launch(I3(), I4(), I2()); break;
        case 44: launch(I3(), I4(), I4()); break;
        case 82: launch(I3(), I8(), I2()); break;
        case 83: launch(I3(), I8(), I3()); break;
        case 162: launch(I3(), I16(), I2()); break;
        case 164: launch(I3(), I16(), I4()); break;
        default: launch(I3(), I8(), I4()); break;
        }
    else if (key == 22)
      launch(I2(), I8(), I4());
    else
      throw std::runtime_error("spmv_fp32: unsupported block shape");
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }

  namespace
  {
    // one thread per (block row, plane, padded slot)
    __global__ void to_fp32_padded_kernel(int n_brows, int rc, const int64_t *__restrict__ rowptr, const int64_t *__restrict__ rowptr32,
                                          const double *__restrict__ val, float *__restrict__ val32)
    {
      const int row = blockIdx.x;
      if (row >= n_brows) return;
      const int64_t b = rowptr[row], b32 = rowptr32[row];
      const int nb = (int)(rowptr[row + 1] - b), nbp = (int)(rowptr32[row + 1] - b32);
      for (int t = threadIdx.x; t < rc * nbp; t += blockDim.x)
        {
          const int plane = t / nbp, j = t % nbp;
          val32[b32 * rc + (int64_t)plane * nbp + j] = j < nb ? (float)val[b * rc + (int64_t)plane * nb + j] : 0.0f;
        }
    }
    __global__ void pad_cols_kernel(int n_brows, const int64_t *__restrict__ rowptr, const int64_t *__restrict__ rowptr32,
                                    const int *__restrict__ col, int *__restrict__ col32)
    {
      const int row = blockIdx.x;
      if (row >= n_brows) return;
      const int64_t b = rowptr[row], b32 = rowptr32[row];
      const int nb = (int)(rowptr[row + 1] - b), nbp = (int)(rowptr32[row + 1] - b32);
      for (int j = threadIdx.x; j < nbp; j += blockDim.x) col32[b32 + j] = j < nb ? col[b + j] : 0;
    }
  } // namespace

  void make_fp32_copy(Context &ctx, Bcsr &A)
  {
    if (!A.n_brows) return;
    if (!A.rowptr32.p)
      {
        // padded pattern, once (the sparsity pattern is fixed)
        const std::vector<int64_t> rp = A.rowptr.to_host(ctx.stream);
        std::vector<int64_t> rp32(rp.size(), 0);
        for (int i = 0; i < A.n_brows; ++i) rp32[i + 1] = rp32[i] + ((rp[i + 1] - rp[i] + 3) / 4) * 4;
        A.rowptr32.upload(rp32, ctx.stream);
        A.col32.alloc(rp32[A.n_brows]);
        A.val32.alloc((size_t)rp32[A.n_brows] * A.R * A.C);
        pad_cols_kernel<<<A.n_brows, 32, 0, ctx.stream>>>(A.n_brows, A.rowptr.p, A.rowptr32.p, A.col.p, A.col32.p);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
    to_fp32_padded_kernel<<<A.n_brows, 128, 0, ctx.stream>>>(A.n_brows, A.R * A.C, A.rowptr.p, A.rowptr32.p, A.val.p, A.val32.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }

  // ---------------------------------------------------------------------------
  // BLAS-1
  // ---------------------------------------------------------------------------
  namespace
  {
    constexpr int kThreads = 256;

    struct Seg
    {
      int64_t len0, shift, total; // i >= len0 -> i + shift
      __host__ __device__ int64_t operator()(int64_t i) const { return i < len0 ? i : i + shift; }
    };
    inline Seg seg_of(const VecSpace &v) { return Seg{v.len0, v.off1 - v.len0, v.len0 + v.len1}; }

    inline int grid_for(const Context &ctx, int64_t n)
    {
      const int64_t want = (n + kThreads * 4 - 1) / (kThreads * 4);
      return (int)std::max<int64_t>(1, std::min<int64_t>(want, std::min(kMaxPartials, ctx.sm_count * 8)));
    }

    __device__ __forceinline__ double block_sum(double v)
    {
      __shared__ double sh[kThreads / 32];
      v = warp_sum(v);
      const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
      __syncthreads();
      if (l == 0) sh[w] = v;
      __syncthreads();
      double r = 0.0;
      if (w == 0)
        {
          r = l < kThreads / 32 ? sh[l] : 0.0;
          r = warp_sum(r);
        }
      return r;
    }

    __global__ void __launch_bounds__(kThreads) dot_partial_kernel(Seg sg, const double *__restrict__ x,
                                                                   const double *__restrict__ y, double *__restrict__ partial)
    {
      double s = 0.0;
      for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < sg.total; k += (int64_t)gridDim.x * blockDim.x)
        {
          const int64_t i = sg(k);
          s = fma(x[i], y[i], s);
        }
      s = block_sum(s);
      if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }

    __global__ void __launch_bounds__(kThreads)
    add_and_dot_partial_kernel(Seg sg, double *__restrict__ aux, double a, const double *__restrict__ V,
                               const double *__restrict__ W, double *__restrict__ partial)
    {
      double s = 0.0;
      for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < sg.total; k += (int64_t)gridDim.x * blockDim.x)
        {
          const int64_t i = sg(k);
          const double t = fma(a, V[i], aux[i]);
          aux[i] = t;
          // W may alias aux (norm of the updated vector)
          const double w = (W == aux) ? t : W[i];
          s = fma(t, w, s);
        }
      s = block_sum(s);
      if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }

    __global__ void __launch_bounds__(kThreads) reduce_final_kernel(int n, const double *__restrict__ partial, double *__restrict__ out)
    {
      double s = 0.0;
      for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
      s = block_sum(s);
      if (threadIdx.x == 0) out[0] = s;
    }

    double finish_reduction(Context &ctx, int g)
    {
      reduce_final_kernel<<<1, kThreads, 0, ctx.stream>>>(g, ctx.partials.p, ctx.results.p);
      IFEM_KERNEL_CHECK();
      ctx.kernel_launches++;
      if (ctx.comm && ctx.comm->size > 1) comm_allreduce_sum(*ctx.comm, ctx.results.p, 1, ctx.stream);
      IFEM_CUDA(cudaMemcpyAsync(ctx.h_results, ctx.results.p, sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
      IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
      return ctx.h_results[0];
    }

    template <typename F>
    __global__ void __launch_bounds__(kThreads) map_kernel(Seg sg, F f)
    {
      for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < sg.total; k += (int64_t)gridDim.x * blockDim.x) f(sg(k));
    }

    template <typename F>
    void map(Context &ctx, const VecSpace &vs, F f)
    {
      const int64_t n = vs.n_owned();
      if (n <= 0) return;
      const int64_t want = (n + kThreads - 1) / kThreads;
      const int g = (int)std::min<int64_t>(want, (int64_t)ctx.sm_count * 16);
      map_kernel<<<g, kThreads, 0, ctx.stream>>>(seg_of(vs), f);
      IFEM_KERNEL_CHECK();
      ctx.kernel_launches++;
    }
  } // namespace

  // ---------------------------------------------------------------------------
  // Fused classical Gram-Schmidt with re-orthogonalisation (CGS2) for the Arnoldi step of GMRES: all inner products of a pass
  // come out of ONE reduction (one kernel pair, one all-reduce over the ranks) instead of one per basis vector as in the
  // modified Gram-Schmidt loop, and the host reads the coefficients of both passes and the norm with ONE synchronisation.
  // ---------------------------------------------------------------------------
  namespace
  {
    struct BasisPtrs
    {
      const double *v[kOrthoMaxBasis];
    };

    // partial[t * gridDim.x + block] = sum over the block's elements of V_t[i] * aux[i], t < k
    __global__ void __launch_bounds__(kThreads) multi_dot_partial_kernel(Seg sg, int k, BasisPtrs V, const double *__restrict__ aux,
                                                                         double *__restrict__ partial)
    {
      for (int t0 = 0; t0 < k; t0 += 8)
        {
          double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          const int nt = min(8, k - t0);
          for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < sg.total; e += (int64_t)gridDim.x * blockDim.x)
            {
              const int64_t i = sg(e);
              const double a = aux[i];
#pragma unroll
              for (int t = 0; t < 8; ++t)
                if (t < nt) s[t] = fma(V.v[t0 + t][i], a, s[t]);
            }
#pragma unroll
          for (int t = 0; t < 8; ++t)
            {
              if (t >= nt) break; // uniform over the block
              const double r = block_sum(s[t]);
              if (threadIdx.x == 0) partial[(int64_t)(t0 + t) * gridDim.x + blockIdx.x] = r;
            }
        }
    }

    // out[t] = sum of the g partials of vector t; one block per t
    __global__ void __launch_bounds__(kThreads) multi_reduce_final_kernel(int g, const double *__restrict__ partial, double *__restrict__ out)
    {
      double s = 0.0;
      for (int i = threadIdx.x; i < g; i += blockDim.x) s += partial[(int64_t)blockIdx.x * g + i];
      s = block_sum(s);
      if (threadIdx.x == 0) out[blockIdx.x] = s;
    }

    // aux -= sum_t h[t] V_t  (h on the device: no host round trip between the reduction and the update)
    __global__ void __launch_bounds__(kThreads) multi_axpy_kernel(Seg sg, int k, BasisPtrs V, const double *__restrict__ h, double *__restrict__ aux)
    {
      __shared__ double sh[kOrthoMaxBasis];
      for (int t = threadIdx.x; t < k; t += blockDim.x) sh[t] = h[t];
      __syncthreads();
      for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < sg.total; e += (int64_t)gridDim.x * blockDim.x)
        {
          const int64_t i = sg(e);
          double a = aux[i];
          for (int t = 0; t < k; ++t) a = fma(-sh[t], V.v[t][i], a);
          aux[i] = a;
        }
    }
  } // namespace

  double orthogonalise_cgs2(Context &ctx, const VecSpace &n, int k, const double *const *basis, double *aux, double *h)
  {
    if (k > kOrthoMaxBasis) throw std::runtime_error("orthogonalise_cgs2: basis larger than 64 vectors");
    BasisPtrs V;
    for (int t = 0; t < k; ++t) V.v[t] = basis[t];
    const Seg sg = seg_of(n);
    const int g = std::min(grid_for(ctx, n.n_owned()), kOrthoMaxGrid);
    const bool many = ctx.comm && ctx.comm->size > 1;
    double *res = ctx.results.p; // [0, k) pass 1, [k, 2k) pass 2, [2k] |aux|^2
    for (int pass = 0; pass < 2; ++pass)
      {
        double *hp = res + (size_t)pass * k;
        multi_dot_partial_kernel<<<g, kThreads, 0, ctx.stream>>>(sg, k, V, aux, ctx.partials.p);
        multi_reduce_final_kernel<<<k, kThreads, 0, ctx.stream>>>(g, ctx.partials.p, hp);
        IFEM_KERNEL_CHECK();
        if (many) comm_allreduce_sum(*ctx.comm, hp, k, ctx.stream);
        multi_axpy_kernel<<<g, kThreads, 0, ctx.stream>>>(sg, k, V, hp, aux);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches += 3;
      }
    dot_partial_kernel<<<g, kThreads, 0, ctx.stream>>>(sg, aux, aux, ctx.partials.p);
    reduce_final_kernel<<<1, kThreads, 0, ctx.stream>>>(g, ctx.partials.p, res + 2 * k);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches += 2;
    if (many) comm_allreduce_sum(*ctx.comm, res + 2 * k, 1, ctx.stream);
    IFEM_CUDA(cudaMemcpyAsync(ctx.h_results, res, (size_t)(2 * k + 1) * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
    for (int t = 0; t < k; ++t) h[t] = ctx.h_results[t] + ctx.h_results[k + t];
    return std::sqrt(std::max(0.0, ctx.h_results[2 * k]));
  }

  double dot(Context &ctx, const VecSpace &n, const double *x, const double *y)
  {
    const int g = grid_for(ctx, n.n_owned());
    dot_partial_kernel<<<g, kThreads, 0, ctx.stream>>>(seg_of(n), x, y, ctx.partials.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
    return finish_reduction(ctx, g);
  }

  double nrm2(Context &ctx, const VecSpace &n, const double *x) { return std::sqrt(dot(ctx, n, x, x)); }

  double add_and_dot(Context &ctx, const VecSpace &n, double *aux, double a, const double *V, const double *W)
  {
    const int g = grid_for(ctx, n.n_owned());
    add_and_dot_partial_kernel<<<g, kThreads, 0, ctx.stream>>>(seg_of(n), aux, a, V, W, ctx.partials.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
    return finish_reduction(ctx, g);
  }

  void axpy(Context &ctx, const VecSpace &n, double a, const double *x, double *y)
  {
    map(ctx, n, [=] __device__(int64_t i) { y[i] = fma(a, x[i], y[i]); });
  }
  void axpby(Context &ctx, const VecSpace &n, double a, const double *x, double b, double *y)
  {
    map(ctx, n, [=] __device__(int64_t i) { y[i] = a * x[i] + b * y[i]; });
  }
  void scale(Context &ctx, const VecSpace &n, double a, double *x)
  {
    map(ctx, n, [=] __device__(int64_t i) { x[i] *= a; });
  }
  void equ(Context &ctx, const VecSpace &n, double a, const double *x, double *y)
  {
    map(ctx, n, [=] __device__(int64_t i) { y[i] = a * x[i]; });
  }
  void copy(Context &ctx, const VecSpace &n, const double *x, double *y)
  {
    // whole allocation, ghosts included (they are scratch between halo updates)
    if (n.n_alloc > 0) IFEM_CUDA(cudaMemcpyAsync(y, x, n.n_alloc * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
  }
  void fill(Context &ctx, const VecSpace &n, double v, double *x)
  {
    if (v == 0.0)
      {
        if (n.n_alloc > 0) IFEM_CUDA(cudaMemsetAsync(x, 0, n.n_alloc * sizeof(double), ctx.stream));
        return;
      }
    map(ctx, n, [=] __device__(int64_t i) { x[i] = v; });
  }
  void lin3(Context &ctx, const VecSpace &n, double *z, const double *x, double a, const double *y, double b, const double *w)
  {
    map(ctx, n, [=] __device__(int64_t i) { z[i] = x[i] + a * y[i] + b * w[i]; });
  }
  void set_indexed(Context &ctx, int n_idx, const int *idx, const double *vals, double *x)
  {
    map(ctx, n_idx, [=] __device__(int64_t k) { x[idx[k]] = vals ? vals[k] : 0.0; });
  }
  void set_flagged(Context &ctx, int64_t n, const unsigned char *flag, const double *vals, double *x)
  {
    map(ctx, n, [=] __device__(int64_t g) {
      if (flag[g]) x[g] = vals ? vals[g] : 0.0;
    });
  }
  void hadamard(Context &ctx, const VecSpace &n, const double *d, const double *x, double *y)
  {
    map(ctx, n, [=] __device__(int64_t i) { y[i] = d[i] * x[i]; });
  }
  void divide(Context &ctx, const VecSpace &n, const double *d, double *y)
  {
    map(ctx, n, [=] __device__(int64_t i) { y[i] /= d[i]; });
  }
  void reciprocal(Context &ctx, const VecSpace &n, const double *x, double *y)
  {
    map(ctx, n, [=] __device__(int64_t i) { y[i] = 1.0 / x[i]; });
  }

  void block_diag_apply(Context &ctx, int n_nodes, int bs, const double *binv, const double *x, double *y)
  {
    if (bs == 2)
      map(ctx, n_nodes, [=] __device__(int64_t i) {
        const double *b = binv + i * 4;
        const double x0 = x[2 * i], x1 = x[2 * i + 1];
        y[2 * i] = b[0] * x0 + b[1] * x1;
        y[2 * i + 1] = b[2] * x0 + b[3] * x1;
      });
    else if (bs == 3)
      map(ctx, n_nodes, [=] __device__(int64_t i) {
        const double *b = binv + i * 9;
        const double x0 = x[3 * i], x1 = x[3 * i + 1], x2 = x[3 * i + 2];
        y[3 * i] = b[0] * x0 + b[1] * x1 + b[2] * x2;
        y[3 * i + 1] = b[3] * x0 + b[4] * x1 + b[5] * x2;
        y[3 * i + 2] = b[6] * x0 + b[7] * x1 + b[8] * x2;
      });
    else if (bs == 1)
      map(ctx, n_nodes, [=] __device__(int64_t i) { y[i] = binv[i] * x[i]; });
    else
      throw std::runtime_error("block_diag_apply: unsupported block size");
  }

  void block_diag_inverse(Context &ctx, const Bcsr &A, double *binv)
  {
    if (A.R != A.C) throw std::runtime_error("block_diag_inverse: square blocks required");
    const int bs = A.R;
    const int64_t *rowptr = A.rowptr.p;
    const int *col = A.col.p;
    const double *val = A.val.p;
    map(ctx, A.n_brows, [=] __device__(int64_t i) {
      const int64_t base = rowptr[i];
      const int nb = (int)(rowptr[i + 1] - base);
      int lo = 0, hi = nb - 1, j = -1;
      while (lo <= hi)
        {
          const int mid = (lo + hi) >> 1;
          const int c = col[base + mid];
          if (c == (int)i) { j = mid; break; }
          if (c < (int)i) lo = mid + 1; else hi = mid - 1;
        }
      double m[9];
      for (int k = 0; k < bs * bs; ++k) m[k] = j >= 0 ? val[base * bs * bs + (int64_t)k * nb + j] : 0.0;
      double *o = binv + i * bs * bs;
      if (bs == 1)
        o[0] = 1.0 / m[0];
      else if (bs == 2)
        {
          const double d = 1.0 / (m[0] * m[3] - m[1] * m[2]);
          o[0] = m[3] * d; o[1] = -m[1] * d; o[2] = -m[2] * d; o[3] = m[0] * d;
        }
      else
        {
          const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
          const double d = 1.0 / (m[0] * c00 + m[1] * c01 + m[2] * c02);
          o[0] = c00 * d; o[1] = (m[2] * m[7] - m[1] * m[8]) * d; o[2] = (m[1] * m[5] - m[2] * m[4]) * d;
          o[3] = c01 * d; o[4] = (m[0] * m[8] - m[2] * m[6]) * d; o[5] = (m[2] * m[3] - m[0] * m[5]) * d;
          o[6] = c02 * d; o[7] = (m[1] * m[6] - m[0] * m[7]) * d; o[8] = (m[0] * m[4] - m[1] * m[3]) * d;
        }
    });
  }
} // namespace ifem
