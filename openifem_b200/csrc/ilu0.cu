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

    // forward sweep y = L^-1 b (unit diagonal), then backward sweep x = U^-1 y; one thread per row inside a level
    __global__ void __launch_bounds__(kIluThreads) ilu0_solve_kernel(int n_lower, const int *__restrict__ order_l, const int *__restrict__ level_l,
                                                                     int n_upper, const int *__restrict__ order_u, const int *__restrict__ level_u,
                                                                     const int *__restrict__ rowptr, const int *__restrict__ col,
                                                                     const int *__restrict__ diag, const double *__restrict__ val,
                                                                     const double *__restrict__ b, double *__restrict__ y, double *__restrict__ x)
    {
      for (int lev = 0; lev < n_lower; ++lev)
        {
          for (int r = level_l[lev] + threadIdx.x; r < level_l[lev + 1]; r += blockDim.x)
            {
              const int i = order_l[r];
              double s = b[i];
              for (int p = rowptr[i]; p < diag[i]; ++p) s -= val[p] * y[col[p]];
              y[i] = s;
            }
          __syncthreads();
        }
      for (int lev = 0; lev < n_upper; ++lev)
        {
          for (int r = level_u[lev] + threadIdx.x; r < level_u[lev + 1]; r += blockDim.x)
            {
              const int i = order_u[r];
              double s = y[i];
              for (int p = diag[i] + 1; p < rowptr[i + 1]; ++p) s -= val[p] * x[col[p]];
              const double d = val[diag[i]];
              x[i] = d != 0.0 ? s / d : s;
            }
          __syncthreads();
        }
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
    IFEM_CUDA(cudaStreamSynchronize(s));
  }

  void Ilu0::factor(Context &ctx)
  {
    ilu0_factor_kernel<<<1, kIluThreads, 0, ctx.stream>>>(n_levels_lower, order_lower.p, level_lower.p, rowptr.p, col.p, diag.p, val.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }

  void Ilu0::solve(Context &ctx, const double *b, double *x)
  {
    ilu0_solve_kernel<<<1, kIluThreads, 0, ctx.stream>>>(n_levels_lower, order_lower.p, level_lower.p, n_levels_upper, order_upper.p, level_upper.p,
                                                        rowptr.p, col.p, diag.p, val.p, b, tmp.p, x);
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
