#include "ilu0.h"

#include <algorithm>

namespace ifem
{
  namespace
  {
    constexpr int kIluThreads = 1024;

    __device__ __forceinline__ int find_col(const int *__restrict__ col, int lo, int hi, int c)
    {
      --hi;
      while (lo <= hi)
        {
          const int mid = (lo + hi) >> 1;
          const int v = col[mid];
          if (v == c) return mid;
          if (v < c) lo = mid + 1; else hi = mid - 1;
        }
      return -1;
    }

    // In-place ILU(0), one warp per row, levels in sequence. Row i: for every k < i of its pattern (ascending)
    // l_ik = a_ik / u_kk, then a_ij -= l_ik u_kj for the j > k that row i holds.
    __global__ void __launch_bounds__(kIluThreads) ilu0_factor_kernel(int n_levels, const int *__restrict__ order, const int *__restrict__ level,
                                                                      const int *__restrict__ rowptr, const int *__restrict__ col,
                                                                      const int *__restrict__ diag, double *__restrict__ val)
    {
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
      for (int lev = 0; lev < n_levels; ++lev)
        {
          for (int r = level[lev] + warp; r < level[lev + 1]; r += n_warps)
            {
              const int i = order[r];
              const int end = rowptr[i + 1];
              for (int p = rowptr[i]; p < diag[i]; ++p)
                {
                  const int k = col[p];
                  const double ukk = val[diag[k]];
                  const double lik = ukk != 0.0 ? val[p] / ukk : 0.0;
                  __syncwarp();
                  if (lane == 0) val[p] = lik;
                  for (int q = diag[k] + 1 + lane; q < rowptr[k + 1]; q += 32)
                    {
                      const int pos = find_col(col, p + 1, end, col[q]);
                      if (pos >= 0) val[pos] -= lik * val[q];
                    }
                  __syncwarp();
                }
            }
          __syncthreads();
        }
    }

    // values of the factors -> the level-ordered padded copy the sweeps stream
    __global__ void ilu0_pack_kernel(int64_t n_l, const int *__restrict__ lsrc, double *__restrict__ lval, int64_t n_u, const int *__restrict__ usrc,
                                     double *__restrict__ uval, int n, const int *__restrict__ order_u, const int *__restrict__ diag,
                                     const double *__restrict__ val, double *__restrict__ udinv)
    {
      const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (t < n_l) lval[t] = lsrc[t] >= 0 ? val[lsrc[t]] : 0.0;
      if (t < n_u) uval[t] = usrc[t] >= 0 ? val[usrc[t]] : 0.0;
      if (t < n)
        {
          const double d = val[diag[order_u[t]]];
          udinv[t] = d != 0.0 ? 1.0 / d : 1.0;
        }
    }

#ifdef IFEM_EMULATED_DEVICE
    __device__ __forceinline__ void prefetch_l1(const void *) {}
#else
    __device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#endif

    // One sweep over the levels: a warp per row, lanes over the row's padded slots, the recurrence on the work vector w (shared
    // memory when it fits). UPPER: w_i = (w_i - sum) / u_ii, else w_i -= sum (unit diagonal).
    template <bool UPPER>
    __device__ __forceinline__ void ilu0_sweep(int n_levels, const int *__restrict__ level, const int *__restrict__ order, int width,
                                               const int *__restrict__ col, const double *__restrict__ val, const double *__restrict__ dinv,
                                               double *w, const int *s_level)
    {
      constexpr int kAhead = 4; // levels between a prefetch and its use
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
      for (int lev = 0; lev < n_levels; ++lev)
        {
          if (lev + kAhead < n_levels)
            {
              const int r = s_level[lev + kAhead] + warp;
              if (r < s_level[lev + kAhead + 1])
                {
                  prefetch_l1(val + (int64_t)r * width + lane);
                  if (lane < 16) prefetch_l1(col + (int64_t)r * width + 2 * lane);
                  if (lane == 0) prefetch_l1(order + r);
                }
            }
          const int end = s_level[lev + 1];
          for (int r = s_level[lev] + warp; r < end; r += n_warps)
            {
              double sum = 0.0;
              for (int k = lane; k < width; k += 32)
                {
                  const double v = val[(int64_t)r * width + k];
                  if (v != 0.0) sum += v * w[col[(int64_t)r * width + k]];
                }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
              if (lane == 0)
                {
                  const int i = order[r];
                  w[i] = UPPER ? (w[i] - sum) * dinv[r] : w[i] - sum;
                }
            }
          __syncthreads();
        }
    }

    // x = U^-1 L^-1 b: forward and backward sweep in one launch
    template <bool SHARED>
    __global__ void __launch_bounds__(kIluThreads) ilu0_solve_kernel(int n, int n_lower, const int *__restrict__ level_l, const int *__restrict__ order_l,
                                                                     int w_lower, const int *__restrict__ lcol, const double *__restrict__ lval,
                                                                     int n_upper, const int *__restrict__ level_u, const int *__restrict__ order_u,
                                                                     int w_upper, const int *__restrict__ ucol, const double *__restrict__ uval,
                                                                     const double *__restrict__ udinv, const double *__restrict__ b,
                                                                     double *__restrict__ work, double *__restrict__ x)
    {
      extern __shared__ double smem[];
      // level offsets of both sweeps first (ints), then the work vector
      int *s_level_l = reinterpret_cast<int *>(smem);
      int *s_level_u = s_level_l + (n_lower + 1);
      const int n_ints = n_lower + n_upper + 2;
      double *w = SHARED ? smem + (n_ints + 1) / 2 : work;
      for (int k = threadIdx.x; k <= n_lower; k += blockDim.x) s_level_l[k] = level_l[k];
      for (int k = threadIdx.x; k <= n_upper; k += blockDim.x) s_level_u[k] = level_u[k];
      for (int i = threadIdx.x; i < n; i += blockDim.x) w[i] = b[i];
      __syncthreads();
      ilu0_sweep<false>(n_lower, level_l, order_l, w_lower, lcol, lval, nullptr, w, s_level_l);
      ilu0_sweep<true>(n_upper, level_u, order_u, w_upper, ucol, uval, udinv, w, s_level_u);
      for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = w[i];
    }

    __global__ void bcsr_to_scalar_kernel(int n_brows, int bs, const int64_t *__restrict__ rp, const double *__restrict__ val,
                                          const int *__restrict__ rowptr_s, double *__restrict__ val_s)
    {
      const int row = blockIdx.x * blockDim.x + threadIdx.x;
      if (row >= n_brows) return;
      const int64_t base = rp[row];
      const int nb = (int)(rp[row + 1] - base);
      for (int r = 0; r < bs; ++r)
        {
          double *out = val_s + rowptr_s[bs * row + r];
          for (int j = 0; j < nb; ++j)
            for (int c = 0; c < bs; ++c) out[j * bs + c] = val[base * bs * bs + (int64_t)(r * bs + c) * nb + j];
        }
    }
  } // namespace

  void Ilu0::setup(Context &ctx, const std::vector<int64_t> &rp, const std::vector<int> &ci)
  {
    n = (int)rp.size() - 1;
    nnz = rp[n];
    std::vector<int> rp32(rp.begin(), rp.end()), dg(n, -1), lev_l(n, 0), lev_u(n, 0);
    for (int i = 0; i < n; ++i)
      {
        for (int64_t p = rp[i]; p < rp[i + 1]; ++p)
          if (ci[p] == i) dg[i] = (int)p;
        if (dg[i] < 0) throw std::runtime_error("Ilu0: a row has no diagonal entry");
      }
    // dependency levels: row i of L waits for the rows k < i of its pattern, row i of U for the rows j > i
    int nl = 0, nu = 0;
    for (int i = 0; i < n; ++i)
      {
        int l = 0;
        for (int64_t p = rp[i]; p < dg[i]; ++p) l = std::max(l, lev_l[ci[p]] + 1);
        lev_l[i] = l;
        nl = std::max(nl, l + 1);
      }
    for (int i = n - 1; i >= 0; --i)
      {
        int l = 0;
        for (int64_t p = dg[i] + 1; p < rp[i + 1]; ++p) l = std::max(l, lev_u[ci[p]] + 1);
        lev_u[i] = l;
        nu = std::max(nu, l + 1);
      }
    auto group = [&](const std::vector<int> &lev, int n_lev, std::vector<int> &order, std::vector<int> &off) {
      off.assign(n_lev + 1, 0);
      for (int i = 0; i < n; ++i) off[lev[i] + 1]++;
      for (int l = 0; l < n_lev; ++l) off[l + 1] += off[l];
      order.resize(n);
      std::vector<int> cur(off.begin(), off.end() - 1);
      for (int i = 0; i < n; ++i) order[cur[lev[i]]++] = i;
    };
    std::vector<int> ol, pl, ou, pu;
    group(lev_l, nl, ol, pl);
    group(lev_u, nu, ou, pu);
    n_levels_lower = nl;
    n_levels_upper = nu;
    cudaStream_t s = ctx.stream;
    rowptr.upload(rp32, s);
    col.upload(ci, s);
    diag.upload(dg, s);
    order_lower.upload(ol, s);
    level_lower.upload(pl, s);
    order_upper.upload(ou, s);
    level_upper.upload(pu, s);
    val.alloc((size_t)nnz);
    tmp.alloc((size_t)n);
    // level-ordered padded copies for the sweeps
    int ml = 1, mu = 1;
    for (int i = 0; i < n; ++i)
      {
        ml = std::max(ml, dg[i] - (int)rp[i]);
        mu = std::max(mu, (int)rp[i + 1] - dg[i] - 1);
      }
    w_lower = (ml + 31) / 32 * 32;
    w_upper = (mu + 31) / 32 * 32;
    std::vector<int> lc((size_t)n * w_lower, 0), ls((size_t)n * w_lower, -1), uc((size_t)n * w_upper, 0), us((size_t)n * w_upper, -1);
    for (int r = 0; r < n; ++r)
      {
        const int il = ol[r], iu = ou[r];
        int k = 0;
        for (int64_t p = rp[il]; p < dg[il]; ++p, ++k)
          {
            lc[(size_t)r * w_lower + k] = ci[p];
            ls[(size_t)r * w_lower + k] = (int)p;
          }
        k = 0;
        for (int64_t p = dg[iu] + 1; p < rp[iu + 1]; ++p, ++k)
          {
            uc[(size_t)r * w_upper + k] = ci[p];
            us[(size_t)r * w_upper + k] = (int)p;
          }
      }
    lcol.upload(lc, s);
    lsrc.upload(ls, s);
    ucol.upload(uc, s);
    usrc.upload(us, s);
    lval.alloc(lc.size());
    uval.alloc(uc.size());
    udinv.alloc((size_t)n);
    IFEM_CUDA(cudaStreamSynchronize(s));
  }

  void Ilu0::factor(Context &ctx)
  {
    ilu0_factor_kernel<<<1, kIluThreads, 0, ctx.stream>>>(n_levels_lower, order_lower.p, level_lower.p, rowptr.p, col.p, diag.p, val.p);
    IFEM_KERNEL_CHECK();
    const int64_t nl = (int64_t)n * w_lower, nu = (int64_t)n * w_upper, total = std::max(nl, nu);
    ilu0_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx.stream>>>(nl, lsrc.p, lval.p, nu, usrc.p, uval.p, n, order_upper.p, diag.p, val.p,
                                                                             udinv.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches += 2;
  }

  void Ilu0::solve(Context &ctx, const double *b, double *x)
  {
    const size_t ints = (size_t)n_levels_lower + n_levels_upper + 2;
    const size_t level_bytes = (ints + 1) / 2 * sizeof(double), with_work = level_bytes + (size_t)n * sizeof(double);
    constexpr size_t kMaxShared = 200 * 1024;
    if (level_bytes > kMaxShared) throw std::runtime_error("Ilu0::solve: too many dependency levels for the one-CTA sweep");
    if (with_work <= kMaxShared)
      {
        static bool opted = false;
        if (!opted)
          {
            IFEM_CUDA(cudaFuncSetAttribute(ilu0_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxShared));
            opted = true;
          }
        ilu0_solve_kernel<true><<<1, kIluThreads, with_work, ctx.stream>>>(n, n_levels_lower, level_lower.p, order_lower.p, w_lower, lcol.p, lval.p,
                                                                         n_levels_upper, level_upper.p, order_upper.p, w_upper, ucol.p, uval.p, udinv.p,
                                                                         b, tmp.p, x);
      }
    else
      ilu0_solve_kernel<false><<<1, kIluThreads, level_bytes, ctx.stream>>>(n, n_levels_lower, level_lower.p, order_lower.p, w_lower, lcol.p, lval.p,
                                                                           n_levels_upper, level_upper.p, order_upper.p, w_upper, ucol.p, uval.p, udinv.p,
                                                                           b, tmp.p, x);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }

  void scalar_pattern(const Pattern &P, int bs, std::vector<int64_t> &rowptr, std::vector<int> &col)
  {
    const int n = P.n_rows;
    rowptr.assign((size_t)n * bs + 1, 0);
    for (int i = 0; i < n; ++i)
      for (int r = 0; r < bs; ++r) rowptr[(size_t)bs * i + r + 1] = (P.rowptr[i + 1] - P.rowptr[i]) * bs;
    for (size_t k = 0; k < (size_t)n * bs; ++k) rowptr[k + 1] += rowptr[k];
    col.resize(rowptr.back());
    for (int i = 0; i < n; ++i)
      for (int r = 0; r < bs; ++r)
        {
          int64_t o = rowptr[(size_t)bs * i + r];
          for (int64_t p = P.rowptr[i]; p < P.rowptr[i + 1]; ++p)
            for (int c = 0; c < bs; ++c) col[o++] = bs * P.col[p] + c;
        }
  }

  void bcsr_to_scalar(Context &ctx, const Bcsr &A, const int *rowptr_s, double *val_s)
  {
    const int n = A.n_brows;
    bcsr_to_scalar_kernel<<<(n + 127) / 128, 128, 0, ctx.stream>>>(n, A.R, A.rowptr.p, A.val.p, rowptr_s, val_s);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }
} // namespace ifem
