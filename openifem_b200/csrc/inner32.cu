#include "inner32.h"

#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "comm.h"
#include "peer_dev.cuh"

namespace ifem
{
  namespace
  {
    constexpr int kT = 256;

    // Loads of the product kernels as explicit PTX: matrix values and column indices are read once
    // (ld.global.cs = evict first), x gathers go through the read-only path and stay in L1 / L2.
    __device__ __forceinline__ float ld_cs_ordered(const float *p)
    {
      float v;
      asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(p));
      return v;
    }
    __device__ __forceinline__ int ld_cs_ordered(const int *p)
    {
      int v;
      asm volatile("ld.global.cs.s32 %0, [%1];" : "=r"(v) : "l"(p));
      return v;
    }
    __device__ __forceinline__ int2 ld_cs_ordered(const int2 *p)
    {
      int2 v;
      asm volatile("ld.global.cs.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
      return v;
    }
    __device__ __forceinline__ float2 ld_cs_half2(const __half2 *p)
    {
      unsigned int v;
      asm volatile("ld.global.cs.b32 %0, [%1];" : "=r"(v) : "l"(p));
      return __half22float2(*reinterpret_cast<const __half2 *>(&v));
    }

    // gather source: float4 per node for bs = 2, 3 (one 16-byte load), float per node for bs = 1
    template <int BS>
    struct XGather
    {
      using T = float4;
      static __device__ __forceinline__ T load(const float *x, int c)
      {
        float4 v;
        asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(reinterpret_cast<const float4 *>(x) + c));
        return v;
      }
      static __device__ __forceinline__ float get(const T &v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
    };
    template <>
    struct XGather<1>
    {
      using T = float;
      static __device__ __forceinline__ T load(const float *x, int c)
      {
        float v;
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(x + c));
        return v;
      }
      static __device__ __forceinline__ float get(const T &v, int) { return v; }
    };

    // ---------------------------------------------------------------------------
    // y = A32 x. One warp per slice, one lane per block row; per block slot a warp issues one coalesced load of
    // 32 column indices, one gather per lane and bs*bs coalesced 128-byte loads of matrix values.
    // Software-pipelined: NS block slots per step. All NS * bs*bs value loads of a step are issued first, then the
    // column indices of the NEXT step (so the x gathers never wait for their index), then the NS gathers, then the
    // FMAs: NS * (bs*bs + 2) 128-byte lines in flight per warp.
    // (col has slack behind the last slice for the unconditional prefetch.)
    // ---------------------------------------------------------------------------
    template <int BS, int NS, int MINB>
    __global__ void __launch_bounds__(kT, MINB)
    sell_spmv_pipe_kernel(int n_slices, const int *__restrict__ slice_off, const int *__restrict__ col, const float *__restrict__ val,
                          const float *__restrict__ x, float *__restrict__ y, const int *__restrict__ skip)
    {
      constexpr int RC = BS * BS;
      using G = XGather<BS>;
      if (skip != nullptr && *skip) return; // the solver that enqueued this product has finished (device-resident state)
      const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
      const int lane = threadIdx.x & 31;
      if (warp >= n_slices) return;
      const int s0 = slice_off[warp];
      const int L = slice_off[warp + 1] - s0;
      const int *cp = col + (int64_t)s0 * 32 + lane;
      const float *vp = val + (int64_t)s0 * (RC * 32) + lane;
      float acc[BS];
#pragma unroll
      for (int r = 0; r < BS; ++r) acc[r] = 0.0f;
      int cn[NS];
#pragma unroll
      for (int u = 0; u < NS; ++u) cn[u] = ld_cs_ordered(cp + u * 32);
      int j = 0;
      for (; j + NS <= L; j += NS)
        {
          float a[NS][RC];
#pragma unroll
          for (int u = 0; u < NS; ++u)
#pragma unroll
            for (int k = 0; k < RC; ++k) a[u][k] = ld_cs_ordered(vp + (size_t)(u * RC + k) * 32);
          vp += (size_t)NS * RC * 32;
          int c[NS];
#pragma unroll
          for (int u = 0; u < NS; ++u) c[u] = cn[u];
          cp += NS * 32;
#pragma unroll
          for (int u = 0; u < NS; ++u) cn[u] = ld_cs_ordered(cp + u * 32);
          typename G::T xv[NS];
#pragma unroll
          for (int u = 0; u < NS; ++u) xv[u] = G::load(x, c[u]);
#pragma unroll
          for (int u = 0; u < NS; ++u)
#pragma unroll
            for (int r = 0; r < BS; ++r)
#pragma unroll
              for (int cc = 0; cc < BS; ++cc) acc[r] = fmaf(a[u][r * BS + cc], G::get(xv[u], cc), acc[r]);
        }
      // tail: fewer than NS slots left, their column indices are already in cn[]
#pragma unroll
      for (int u = 0; u < NS - 1; ++u)
        if (j + u < L)
          {
            float a[RC];
#pragma unroll
            for (int k = 0; k < RC; ++k) a[k] = ld_cs_ordered(vp + (size_t)(u * RC + k) * 32);
            const typename G::T xq = G::load(x, cn[u]);
#pragma unroll
            for (int r = 0; r < BS; ++r)
#pragma unroll
              for (int cc = 0; cc < BS; ++cc) acc[r] = fmaf(a[r * BS + cc], G::get(xq, cc), acc[r]);
          }
      float *yp = y + ((int64_t)warp * 32 + lane) * BS;
#pragma unroll
      for (int r = 0; r < BS; ++r) yp[r] = acc[r];
    }

    // fp16 storage: the values of a row are scaled by 1 / max|row| (per scalar row) and stored as half2 = two
    // consecutive block slots of the same row, so a warp still reads full 128-byte lines; the column indices of
    // the two slots travel as one int2. 22 instead of 40 bytes per 3x3 block; products and sums in fp32.
    template <int BS, int NS, int MINB>
    __global__ void __launch_bounds__(kT, MINB)
    sell_spmv_h_kernel(int n_slices, const int *__restrict__ hoff, const int2 *__restrict__ col2, const __half2 *__restrict__ valh,
                       const float *__restrict__ row_scale, const float *__restrict__ x, float *__restrict__ y, const int *__restrict__ skip)
    {
      constexpr int RC = BS * BS;
      using G = XGather<BS>;
      if (skip != nullptr && *skip) return;
      const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
      const int lane = threadIdx.x & 31;
      if (warp >= n_slices) return;
      const int h0 = hoff[warp];
      const int L = hoff[warp + 1] - h0; // double slots
      const int2 *cp = col2 + (int64_t)h0 * 32 + lane;
      const __half2 *vp = valh + (int64_t)h0 * (RC * 32) + lane;
      float acc[BS];
#pragma unroll
      for (int r = 0; r < BS; ++r) acc[r] = 0.0f;
      int2 cn[NS];
#pragma unroll
      for (int u = 0; u < NS; ++u) cn[u] = ld_cs_ordered(cp + u * 32);
      int j = 0;
      for (; j + NS <= L; j += NS)
        {
          float2 a[NS][RC];
#pragma unroll
          for (int u = 0; u < NS; ++u)
#pragma unroll
            for (int k = 0; k < RC; ++k) a[u][k] = ld_cs_half2(vp + (size_t)(u * RC + k) * 32);
          vp += (size_t)NS * RC * 32;
          int2 c[NS];
#pragma unroll
          for (int u = 0; u < NS; ++u) c[u] = cn[u];
          cp += NS * 32;
#pragma unroll
          for (int u = 0; u < NS; ++u) cn[u] = ld_cs_ordered(cp + u * 32);
          typename G::T xa[NS], xb[NS];
#pragma unroll
          for (int u = 0; u < NS; ++u)
            {
              xa[u] = G::load(x, c[u].x);
              xb[u] = G::load(x, c[u].y);
            }
#pragma unroll
          for (int u = 0; u < NS; ++u)
#pragma unroll
            for (int r = 0; r < BS; ++r)
#pragma unroll
              for (int cc = 0; cc < BS; ++cc)
                {
                  acc[r] = fmaf(a[u][r * BS + cc].x, G::get(xa[u], cc), acc[r]);
                  acc[r] = fmaf(a[u][r * BS + cc].y, G::get(xb[u], cc), acc[r]);
                }
        }
#pragma unroll
      for (int u = 0; u < NS - 1; ++u)
        if (j + u < L)
          {
            float2 a[RC];
#pragma unroll
            for (int k = 0; k < RC; ++k) a[k] = ld_cs_half2(vp + (size_t)(u * RC + k) * 32);
            const typename G::T xq = G::load(x, cn[u].x), xw = G::load(x, cn[u].y);
#pragma unroll
            for (int r = 0; r < BS; ++r)
#pragma unroll
              for (int cc = 0; cc < BS; ++cc)
                {
                  acc[r] = fmaf(a[r * BS + cc].x, G::get(xq, cc), acc[r]);
                  acc[r] = fmaf(a[r * BS + cc].y, G::get(xw, cc), acc[r]);
                }
          }
      const int64_t self = (int64_t)warp * 32 + lane;
#pragma unroll
      for (int r = 0; r < BS; ++r) y[self * BS + r] = acc[r] * row_scale[self * BS + r];
    }

    // ---------------------------------------------------------------------------
    // building the copy (one warp per slice, one lane per row)
    // ---------------------------------------------------------------------------
    template <int BS>
    __global__ void __launch_bounds__(kT)
    sell_fill_val_kernel(int n_slices, const int *__restrict__ slice_off, const int *__restrict__ perm_row,
                         const int64_t *__restrict__ rowptr, const double *__restrict__ aval, float *__restrict__ val)
    {
      constexpr int RC = BS * BS;
      const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
      const int lane = threadIdx.x & 31;
      if (warp >= n_slices) return;
      const int s0 = slice_off[warp];
      const int L = slice_off[warp + 1] - s0;
      const int row = perm_row[(int64_t)warp * 32 + lane];
      int64_t base = 0;
      int nb = 0;
      if (row >= 0)
        {
          base = rowptr[row];
          nb = (int)(rowptr[row + 1] - base);
        }
      const double *av = aval + base * RC;
      float *vp = val + (int64_t)s0 * (RC * 32) + lane;
      for (int k = 0; k < RC; ++k)
        for (int j = 0; j < L; ++j) vp[(size_t)(j * RC + k) * 32] = j < nb ? (float)av[(int64_t)k * nb + j] : 0.0f;
    }

    __global__ void __launch_bounds__(kT)
    sell_fill_col_kernel(int n_slices, const int *__restrict__ slice_off, const int *__restrict__ perm_row,
                         const int64_t *__restrict__ rowptr, const int *__restrict__ acol, const int *__restrict__ pos,
                         int *__restrict__ col)
    {
      const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
      const int lane = threadIdx.x & 31;
      if (warp >= n_slices) return;
      const int s0 = slice_off[warp];
      const int L = slice_off[warp + 1] - s0;
      const int self = warp * 32 + lane;
      const int row = perm_row[self];
      int64_t base = 0;
      int nb = 0;
      if (row >= 0)
        {
          base = rowptr[row];
          nb = (int)(rowptr[row + 1] - base);
        }
      int *cp = col + (int64_t)s0 * 32 + lane;
      for (int j = 0; j < L; ++j) cp[(size_t)j * 32] = j < nb ? pos[acol[base + j]] : self; // padding: value 0 times own x
    }

    template <int BS>
    __global__ void __launch_bounds__(kT)
    sell_fill_valh_kernel(int n_slices, const int *__restrict__ hoff, const int *__restrict__ perm_row,
                          const int64_t *__restrict__ rowptr, const double *__restrict__ aval, __half2 *__restrict__ valh,
                          float *__restrict__ row_scale)
    {
      constexpr int RC = BS * BS;
      const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
      const int lane = threadIdx.x & 31;
      if (warp >= n_slices) return;
      const int h0 = hoff[warp];
      const int L = hoff[warp + 1] - h0;
      const int64_t self = (int64_t)warp * 32 + lane;
      const int row = perm_row[self];
      int64_t base = 0;
      int nb = 0;
      if (row >= 0)
        {
          base = rowptr[row];
          nb = (int)(rowptr[row + 1] - base);
        }
      const double *av = aval + base * RC;
      double inv[BS];
#pragma unroll
      for (int r = 0; r < BS; ++r)
        {
          double m = 0.0;
          for (int cc = 0; cc < BS; ++cc)
            for (int j = 0; j < nb; ++j) m = fmax(m, fabs(av[(int64_t)(r * BS + cc) * nb + j]));
          row_scale[self * BS + r] = (float)m;
          inv[r] = m > 0.0 ? 1.0 / m : 0.0;
        }
      __half2 *vp = valh + (int64_t)h0 * (RC * 32) + lane;
      for (int k = 0; k < RC; ++k)
        {
          const double sc = inv[k / BS];
          for (int jj = 0; jj < L; ++jj)
            {
              const int j0 = 2 * jj, j1 = 2 * jj + 1;
              const float lo = j0 < nb ? (float)(av[(int64_t)k * nb + j0] * sc) : 0.0f;
              const float hi = j1 < nb ? (float)(av[(int64_t)k * nb + j1] * sc) : 0.0f;
              vp[(size_t)(jj * RC + k) * 32] = __floats2half2_rn(lo, hi);
            }
        }
    }

    __global__ void __launch_bounds__(kT)
    sell_fill_col2_kernel(int n_slices, const int *__restrict__ hoff, const int *__restrict__ perm_row,
                          const int64_t *__restrict__ rowptr, const int *__restrict__ acol, const int *__restrict__ pos,
                          int2 *__restrict__ col2)
    {
      const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
      const int lane = threadIdx.x & 31;
      if (warp >= n_slices) return;
      const int h0 = hoff[warp];
      const int L = hoff[warp + 1] - h0;
      const int self = warp * 32 + lane;
      const int row = perm_row[self];
      int64_t base = 0;
      int nb = 0;
      if (row >= 0)
        {
          base = rowptr[row];
          nb = (int)(rowptr[row + 1] - base);
        }
      int2 *cp = col2 + (int64_t)h0 * 32 + lane;
      for (int jj = 0; jj < L; ++jj)
        {
          const int j0 = 2 * jj, j1 = 2 * jj + 1;
          cp[(size_t)jj * 32] = make_int2(j0 < nb ? pos[acol[base + j0]] : self, j1 < nb ? pos[acol[base + j1]] : self);
        }
    }

    // binv32[k][i] = binv[perm_row[i]][k]
    __global__ void __launch_bounds__(kT)
    binv_to_sell_kernel(int n_pad, int rc, const int *__restrict__ perm_row, const double *__restrict__ binv, float *__restrict__ out)
    {
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          const int row = perm_row[i];
          for (int k = 0; k < rc; ++k) out[(size_t)k * n_pad + i] = row >= 0 ? (float)binv[(size_t)row * rc + k] : 0.0f;
        }
    }

    // ---------------------------------------------------------------------------
    // Vector kernels of the fp32 solvers: one thread per node of the SELL numbering, grid-stride. Reductions are summed
    // in fp64 and finished inside the producing kernel (finish_reduce, peer_dev.cuh); the scalars of the recurrences live
    // in a device-resident state that the last CTA advances (`adv` = 1) or, when an NCCL all-reduce has to follow the
    // kernel (several ranks without a peer link), a one-thread kernel advances after it.
    // ---------------------------------------------------------------------------
    struct BicgState
    {
      double rho, rho_new, alpha, omega, res2, tol2;
      int its, max_it;
      int done;      // the solve is over: every later kernel of the stream returns at once
      int skip2;     // done, or the second half of the current iteration is not needed (converged at s / breakdown)
      int early;     // x += alpha ph only, then done
      int converged, breakdown, pad;
    };
    enum BicgStage { kBInit, kBDot1, kBS, kBDot2, kBXR };

    __device__ __forceinline__ void bicg_advance(int stage, BicgState *st, const double *red)
    {
      if (stage == kBInit)
        {
          st->rho_new = red[0]; // r0 . r = |r|^2
          st->res2 = red[0];
          if (!(red[0] > st->tol2))
            {
              st->done = 1;
              st->skip2 = 1;
              st->converged = red[0] <= st->tol2 ? 1 : 0;
            }
          return;
        }
      if (st->done) return;
      if (stage == kBDot1)
        {
          if (red[0] == 0.0 || !isfinite(red[0]))
            {
              st->done = 1;
              st->skip2 = 1;
              st->breakdown = 1;
            }
          else
            st->alpha = st->rho_new / red[0];
        }
      else if (stage == kBS)
        {
          st->its += 1;
          st->res2 = red[0];
          if (red[0] <= st->tol2)
            {
              st->early = 1;
              st->skip2 = 1;
            }
        }
      else if (stage == kBDot2)
        {
          if (st->skip2) return;
          if (red[1] == 0.0 || !isfinite(red[1]))
            {
              st->early = 1;
              st->skip2 = 1;
              st->breakdown = 1;
            }
          else
            st->omega = red[0] / red[1];
        }
      else // kBXR
        {
          if (st->early)
            {
              st->done = 1;
              st->converged = st->breakdown ? 0 : 1;
              return;
            }
          st->res2 = red[0];
          st->rho = st->rho_new;
          st->rho_new = red[1];
          if (red[0] <= st->tol2)
            {
              st->done = 1;
              st->converged = 1;
            }
          else if (red[1] == 0.0 || st->omega == 0.0 || !isfinite(red[1]) || !isfinite(red[0]))
            {
              st->done = 1;
              st->breakdown = 1;
            }
          else if (st->its >= st->max_it)
            st->done = 1;
          if (st->done) st->skip2 = 1;
        }
    }

    __global__ void bicg_advance_kernel(int stage, BicgState *st, const double *red) { bicg_advance(stage, st, red); }

    __global__ void bicg_begin_kernel(BicgState *st, double tol2, int max_it)
    {
      st->rho = st->alpha = st->omega = 1.0;
      st->rho_new = st->res2 = 0.0;
      st->tol2 = tol2;
      st->its = 0;
      st->max_it = max_it;
      st->done = st->skip2 = st->early = st->converged = st->breakdown = 0;
    }

    // r = r0 = src / |src| in SELL order; p = v = x = 0
    template <int BS>
    __global__ void __launch_bounds__(kT)
    init_kernel(int n_pad, const int *__restrict__ perm_row, const double *__restrict__ src, double scale, float *__restrict__ r,
                float *__restrict__ r0, float *__restrict__ p, float *__restrict__ v, float *__restrict__ x, BicgState *st,
                double *__restrict__ partials, unsigned int *__restrict__ counter, double *__restrict__ red, PeerDev pd, int adv)
    {
      double acc[1] = {0.0};
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          const int row = perm_row[i];
#pragma unroll
          for (int c = 0; c < BS; ++c)
            {
              const float a = row >= 0 ? (float)(src[(size_t)row * BS + c] * scale) : 0.0f;
              const size_t k = (size_t)i * BS + c;
              r[k] = a;
              r0[k] = a;
              p[k] = 0.0f;
              v[k] = 0.0f;
              x[k] = 0.0f;
              acc[0] += (double)a * (double)a;
            }
        }
      if (finish_reduce<1>(acc, partials, counter, red, pd) && adv && threadIdx.x == 0) bicg_advance(kBInit, st, acc);
    }

    template <int BS>
    __device__ __forceinline__ float4 apply_binv(const float *__restrict__ binv, int n_pad, int i, const float (&a)[BS])
    {
      float o[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int rr = 0; rr < BS; ++rr)
#pragma unroll
        for (int c = 0; c < BS; ++c) o[rr] = fmaf(binv[(size_t)(rr * BS + c) * n_pad + i], a[c], o[rr]);
      return make_float4(o[0], o[1], o[2], o[3]);
    }

    // p = r + beta (p - omega v), beta = (rho_new / rho) (alpha / omega);  ph = D^-1 p
    template <int BS>
    __global__ void __launch_bounds__(kT)
    update_p_kernel(int n_pad, const BicgState *__restrict__ st, const float *__restrict__ r, float *__restrict__ p, const float *__restrict__ v,
                    const float *__restrict__ binv, float4 *__restrict__ ph)
    {
      if (st->done) return;
      const float omega = (float)st->omega;
      const float beta = (float)((st->rho_new / st->rho) * (st->alpha / st->omega));
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          float pc[BS];
#pragma unroll
          for (int c = 0; c < BS; ++c)
            {
              const size_t k = (size_t)i * BS + c;
              pc[c] = fmaf(beta, fmaf(-omega, v[k], p[k]), r[k]);
              p[k] = pc[c];
            }
          ph[i] = apply_binv<BS>(binv, n_pad, i, pc);
        }
    }

    // alpha = rho_new / (r0 . v)
    __global__ void __launch_bounds__(kT)
    dot1_kernel(int64_t n, const float *__restrict__ r0, const float *__restrict__ v, BicgState *st, double *__restrict__ partials,
                unsigned int *__restrict__ counter, double *__restrict__ red, PeerDev pd, int adv)
    {
      if (st->done) return;
      double acc[1] = {0.0};
      for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        acc[0] += (double)r0[k] * (double)v[k];
      if (finish_reduce<1>(acc, partials, counter, red, pd) && adv && threadIdx.x == 0) bicg_advance(kBDot1, st, acc);
    }

    // s = r - alpha v;  sh = D^-1 s;  |s|^2
    template <int BS>
    __global__ void __launch_bounds__(kT)
    update_s_kernel(int n_pad, BicgState *st, const float *__restrict__ r, const float *__restrict__ v, float *__restrict__ s,
                    const float *__restrict__ binv, float4 *__restrict__ sh, double *__restrict__ partials, unsigned int *__restrict__ counter,
                    double *__restrict__ red, PeerDev pd, int adv)
    {
      if (st->done) return;
      const float alpha = (float)st->alpha;
      double acc[1] = {0.0};
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          float sc[BS];
#pragma unroll
          for (int c = 0; c < BS; ++c)
            {
              const size_t k = (size_t)i * BS + c;
              sc[c] = fmaf(-alpha, v[k], r[k]);
              s[k] = sc[c];
              acc[0] += (double)sc[c] * (double)sc[c];
            }
          sh[i] = apply_binv<BS>(binv, n_pad, i, sc);
        }
      if (finish_reduce<1>(acc, partials, counter, red, pd) && adv && threadIdx.x == 0) bicg_advance(kBS, st, acc);
    }

    // omega = (t . s) / (t . t)
    __global__ void __launch_bounds__(kT)
    dot2_kernel(int64_t n, const float *__restrict__ t, const float *__restrict__ s, BicgState *st, double *__restrict__ partials,
                unsigned int *__restrict__ counter, double *__restrict__ red, PeerDev pd, int adv)
    {
      if (st->skip2) return;
      double acc[2] = {0.0, 0.0};
      for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        {
          const double tk = (double)t[k];
          acc[0] += tk * (double)s[k];
          acc[1] += tk * tk;
        }
      if (finish_reduce<2>(acc, partials, counter, red, pd) && adv && threadIdx.x == 0) bicg_advance(kBDot2, st, acc);
    }

    // x += alpha ph + omega sh;  r = s - omega t;  |r|^2 and r0 . r   (omega = 0 when the iteration ends at s)
    template <int BS>
    __global__ void __launch_bounds__(kT)
    update_xr_kernel(int n_pad, BicgState *st, float *__restrict__ x, const float4 *__restrict__ ph, const float4 *__restrict__ sh,
                     float *__restrict__ r, const float *__restrict__ s, const float *__restrict__ t, const float *__restrict__ r0,
                     double *__restrict__ partials, unsigned int *__restrict__ counter, double *__restrict__ red, PeerDev pd, int adv)
    {
      if (st->done) return;
      const float alpha = (float)st->alpha;
      const float omega = st->early ? 0.0f : (float)st->omega;
      double acc[2] = {0.0, 0.0};
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          const float4 a4 = ph[i], b4 = sh[i];
          const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int c = 0; c < BS; ++c)
            {
              const size_t k = (size_t)i * BS + c;
              x[k] = fmaf(alpha, a[c], fmaf(omega, b[c], x[k]));
              const float rc = fmaf(-omega, t[k], s[k]);
              r[k] = rc;
              acc[0] += (double)rc * (double)rc;
              acc[1] += (double)r0[k] * (double)rc;
            }
        }
      if (finish_reduce<2>(acc, partials, counter, red, pd) && adv && threadIdx.x == 0) bicg_advance(kBXR, st, acc);
    }

    // dst (fp64, original numbering) = scale * x (SELL numbering), owned rows only
    template <int BS>
    __global__ void __launch_bounds__(kT)
    final_kernel(int n_pad, const int *__restrict__ perm_row, const float *__restrict__ x, double scale, double *__restrict__ dst)
    {
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          const int row = perm_row[i];
          if (row < 0) continue;
#pragma unroll
          for (int c = 0; c < BS; ++c) dst[(size_t)row * BS + c] = scale * (double)x[(size_t)i * BS + c];
        }
    }

    // dst (fp64, original numbering) = scale * D^-1 r (one block-Jacobi step), owned rows only
    template <int BS>
    __global__ void __launch_bounds__(kT)
    jacobi_out_kernel(int n_pad, const int *__restrict__ perm_row, const float *__restrict__ r, const float *__restrict__ binv, double scale,
                      double *__restrict__ dst)
    {
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          const int row = perm_row[i];
          if (row < 0) continue;
          float rc[BS];
#pragma unroll
          for (int c = 0; c < BS; ++c) rc[c] = r[(size_t)i * BS + c];
          const float4 q = apply_binv<BS>(binv, n_pad, i, rc);
          const float a[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int c = 0; c < BS; ++c) dst[(size_t)row * BS + c] = scale * (double)a[c];
        }
    }

    // x4 (SELL numbering, float4 per node) from a fp64 vector in the original numbering (owned nodes)
    template <int BS>
    __global__ void __launch_bounds__(kT)
    to_sell4_kernel(int n_pad, const int *__restrict__ perm_row, const double *__restrict__ src, float4 *__restrict__ x4)
    {
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          const int row = perm_row[i];
          float o[4] = {0.0f, 0.0f, 0.0f, 0.0f};
          if (row >= 0)
            {
#pragma unroll
              for (int c = 0; c < BS; ++c) o[c] = (float)src[(size_t)row * BS + c];
            }
          x4[i] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }

    // pack the owned entries a neighbour needs: xs floats per node (NCCL path)
    __global__ void __launch_bounds__(kT) halo_pack32_kernel(int n, int xs, const int *__restrict__ idx, const float *__restrict__ v, float *__restrict__ buf)
    {
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t < n * xs) buf[t] = v[(size_t)idx[t / xs] * xs + t % xs];
    }

    // ---------------------------------------------------------------------------
    // fp32 CG on a scalar matrix
    // ---------------------------------------------------------------------------
    struct CgState
    {
      double rr, alpha, beta, tol2, gamma;
      int its, max_it, done, converged;
    };
    enum CgStage { kCInit, kCDot, kCXR, kCGear };

    __device__ __forceinline__ void cg_advance(int stage, CgState *st, const double *red)
    {
      if (stage == kCInit)
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
      if (stage == kCDot)
        {
          if (!(red[0] > 0.0) || !isfinite(red[0]))
            st->done = 1; // p^T A p <= 0: not positive definite in fp32 (or breakdown)
          else
            st->alpha = st->rr / red[0];
        }
      else
        {
          const double rr_new = red[0];
          st->beta = rr_new / st->rr;
          st->rr = rr_new;
          st->its += 1;
          if (!(rr_new > st->tol2) || !isfinite(rr_new) || !isfinite(st->beta))
            {
              st->done = 1;
              st->converged = rr_new <= st->tol2 ? 1 : 0;
            }
          else if (st->its >= st->max_it)
            st->done = 1;
        }
    }

    // Single-reduction CG (Chronopoulos & Gear): with w = A r, gamma = r . r and delta = r . w come out of ONE reduction and
    //   beta = gamma / gamma_old,  alpha = gamma / (delta - beta gamma / alpha_old),
    //   p = r + beta p,  s = w + beta s (= A p),  x += alpha p,  r -= alpha s.
    // One product, one fused reduction and one fused vector kernel per iteration instead of one product, two reductions and
    // three vector kernels: at 8 GPUs an iteration of "CG for Sm" is latency bound (a 44 us product against two all-reduces
    // and five launches), so the reduction count is what it costs. The residual test uses gamma, i.e. it sees the residual of
    // the previous update (one extra product at the very end).
    __device__ __forceinline__ void cg_gear_advance(CgState *st, const double *red)
    {
      if (st->done) return;
      const double gamma = red[0], delta = red[1];
      if (!(gamma > st->tol2) || !isfinite(gamma))
        {
          st->rr = gamma;
          st->done = 1;
          st->converged = gamma <= st->tol2 ? 1 : 0;
          return;
        }
      if (st->its >= st->max_it)
        {
          st->rr = gamma;
          st->done = 1;
          return;
        }
      const double beta = st->its == 0 ? 0.0 : gamma / st->rr;
      const double denom = st->its == 0 ? delta : delta - beta * gamma / st->alpha;
      if (!(denom > 0.0) || !isfinite(denom))
        {
          st->rr = gamma;
          st->done = 1; // not positive definite in fp32 (or breakdown): the caller falls back
          return;
        }
      st->beta = beta;
      st->alpha = gamma / denom;
      st->rr = gamma;
      st->its += 1;
    }

    __global__ void cg_advance_kernel(int stage, CgState *st, const double *red)
    {
      if (stage == kCGear) cg_gear_advance(st, red);
      else cg_advance(stage, st, red);
    }

    // r = src / |src| in SELL order; x = p = s = 0
    __global__ void __launch_bounds__(kT)
    cg_gear_init_kernel(int n_pad, const int *__restrict__ perm_row, const double *__restrict__ src, double scale, float *__restrict__ r,
                        float *__restrict__ p, float *__restrict__ s, float *__restrict__ x)
    {
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          const int row = perm_row[i];
          r[i] = row >= 0 ? (float)(src[row] * scale) : 0.0f;
          p[i] = s[i] = x[i] = 0.0f;
        }
    }

    // gamma = r . r, delta = r . w in one reduction; the last CTA advances the recurrence
    __global__ void __launch_bounds__(kT)
    cg_gear_dot_kernel(int n_pad, const float *__restrict__ r, const float *__restrict__ w, CgState *st, double *__restrict__ partials,
                       unsigned int *__restrict__ counter, double *__restrict__ red, PeerDev pd, int adv)
    {
      if (st->done) return;
      double acc[2] = {0.0, 0.0};
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          const double ri = (double)r[i];
          acc[0] += ri * ri;
          acc[1] += ri * (double)w[i];
        }
      if (finish_reduce<2>(acc, partials, counter, red, pd) && adv && threadIdx.x == 0) cg_gear_advance(st, acc);
    }

    // ---- two-level preconditioned single-reduction CG ----
    // u = M^-1 r, w = A u; gamma = r . u, delta = u . w, rho = r . r in ONE reduction; beta = gamma / gamma_old,
    // alpha = gamma / (delta - beta gamma / alpha_old); p = u + beta p, s = w + beta s, x += alpha p, r -= alpha s
    __device__ __forceinline__ void pcg_advance(CgState *st, const double *red)
    {
      if (st->done) return;
      const double gamma = red[0], delta = red[1], rho = red[2];
      if (!(rho > st->tol2) || !isfinite(rho))
        {
          st->rr = rho;
          st->done = 1;
          st->converged = rho <= st->tol2 ? 1 : 0;
          return;
        }
      if (st->its >= st->max_it)
        {
          st->rr = rho;
          st->done = 1;
          return;
        }
      const double beta = st->its == 0 ? 0.0 : gamma / st->gamma;
      const double denom = st->its == 0 ? delta : delta - beta * gamma / st->alpha;
      if (!(denom > 0.0) || !(gamma > 0.0) || !isfinite(denom))
        {
          st->rr = rho;
          st->done = 1; // breakdown: the caller falls back
          return;
        }
      st->beta = beta;
      st->alpha = gamma / denom;
      st->gamma = gamma;
      st->rr = rho;
      st->its += 1;
    }
    __global__ void pcg_advance_kernel(CgState *st, const double *red) { pcg_advance(st, red); }

    // c[a] = sum of r over the rows of aggregate a, in a fixed order (one CTA per aggregate)
    __global__ void __launch_bounds__(kT)
    pcg_restrict_kernel(const CgState *__restrict__ st, const int *__restrict__ agg_ptr, const int *__restrict__ agg_rows, const float *__restrict__ r,
                        double *__restrict__ c)
    {
      if (st->done) return;
      const int a = blockIdx.x;
      double acc[1] = {0.0};
      for (int k = agg_ptr[a] + threadIdx.x; k < agg_ptr[a + 1]; k += blockDim.x) acc[0] += (double)r[agg_rows[k]];
      cta_sum<1>(acc);
      if (threadIdx.x == 0) c[a] = acc[0];
    }

    // y = E^+ c: one warp per row
    __global__ void __launch_bounds__(kT)
    pcg_coarse_kernel(const CgState *__restrict__ st, int n_c, const double *__restrict__ Einv, const double *__restrict__ c, double *__restrict__ y)
    {
      if (st->done) return;
      const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
      if (row >= n_c) return;
      double s = 0.0;
      for (int j = lane; j < n_c; j += 32) s += Einv[(size_t)row * n_c + j] * c[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) y[row] = s;
    }

    // u = diag^-1 r + Z y (into the gather source of the product)
    __global__ void __launch_bounds__(kT)
    pcg_build_u_kernel(int n_pad, const CgState *__restrict__ st, const float *__restrict__ r, const float *__restrict__ dinv,
                       const int *__restrict__ agg, const double *__restrict__ y, float *__restrict__ u)
    {
      if (st->done) return;
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          const int a = agg[i];
          u[i] = a >= 0 ? fmaf(dinv[i], r[i], (float)y[a]) : 0.0f;
        }
    }

    __global__ void __launch_bounds__(kT)
    pcg_dot_kernel(int n_pad, const float *__restrict__ r, const float *__restrict__ u, const float *__restrict__ w, CgState *st,
                   double *__restrict__ partials, unsigned int *__restrict__ counter, double *__restrict__ red, PeerDev pd, int adv)
    {
      if (st->done) return;
      double acc[3] = {0.0, 0.0, 0.0};
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          const double ri = (double)r[i], ui = (double)u[i];
          acc[0] += ri * ui;
          acc[1] += ui * (double)w[i];
          acc[2] += ri * ri;
        }
      if (finish_reduce<3>(acc, partials, counter, red, pd) && adv && threadIdx.x == 0) pcg_advance(st, acc);
    }

    // p = u + beta p; s = w + beta s; x += alpha p; r -= alpha s
    __global__ void __launch_bounds__(kT)
    pcg_update_kernel(int n_pad, const CgState *__restrict__ st, float *__restrict__ r, const float *__restrict__ u, const float *__restrict__ w,
                      float *__restrict__ p, float *__restrict__ s, float *__restrict__ x)
    {
      if (st->done) return;
      const float alpha = (float)st->alpha, beta = (float)st->beta;
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          const float pi = fmaf(beta, p[i], u[i]), si = fmaf(beta, s[i], w[i]);
          p[i] = pi;
          s[i] = si;
          x[i] = fmaf(alpha, pi, x[i]);
          r[i] = fmaf(-alpha, si, r[i]);
        }
    }

    // p = r + beta p; s = w + beta s; x += alpha p; r -= alpha s
    __global__ void __launch_bounds__(kT)
    cg_gear_update_kernel(int n_pad, const CgState *__restrict__ st, float *__restrict__ r, const float *__restrict__ w, float *__restrict__ p,
                          float *__restrict__ s, float *__restrict__ x)
    {
      if (st->done) return;
      const float alpha = (float)st->alpha, beta = (float)st->beta;
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          const float pi = fmaf(beta, p[i], r[i]), si = fmaf(beta, s[i], w[i]);
          p[i] = pi;
          s[i] = si;
          x[i] = fmaf(alpha, pi, x[i]);
          r[i] = fmaf(-alpha, si, r[i]);
        }
    }

    __global__ void cg_begin_kernel(CgState *st, double tol2, int max_it)
    {
      st->rr = 0.0;
      st->alpha = st->beta = st->gamma = 0.0;
      st->tol2 = tol2;
      st->its = 0;
      st->max_it = max_it;
      st->done = st->converged = 0;
    }

    // r = p = src / |src| in SELL order, x = 0; |r|^2
    __global__ void __launch_bounds__(kT)
    cg_init_kernel(int n_pad, const int *__restrict__ perm_row, const double *__restrict__ src, double scale, float *__restrict__ r,
                   float *__restrict__ p, float *__restrict__ x, CgState *st, double *__restrict__ partials, unsigned int *__restrict__ counter,
                   double *__restrict__ red, PeerDev pd, int adv)
    {
      double acc[1] = {0.0};
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          const int row = perm_row[i];
          const float a = row >= 0 ? (float)(src[row] * scale) : 0.0f;
          r[i] = a;
          p[i] = a;
          x[i] = 0.0f;
          acc[0] += (double)a * (double)a;
        }
      if (finish_reduce<1>(acc, partials, counter, red, pd) && adv && threadIdx.x == 0) cg_advance(kCInit, st, acc);
    }

    // alpha = rr / (p . Ap)
    __global__ void __launch_bounds__(kT)
    cg_dot_kernel(int n_pad, const float *__restrict__ p, const float *__restrict__ ap, CgState *st, double *__restrict__ partials,
                  unsigned int *__restrict__ counter, double *__restrict__ red, PeerDev pd, int adv)
    {
      if (st->done) return;
      double acc[1] = {0.0};
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) acc[0] += (double)p[i] * (double)ap[i];
      if (finish_reduce<1>(acc, partials, counter, red, pd) && adv && threadIdx.x == 0) cg_advance(kCDot, st, acc);
    }

    // x += alpha p; r -= alpha Ap; |r|^2
    __global__ void __launch_bounds__(kT)
    cg_xr_kernel(int n_pad, CgState *st, float *__restrict__ x, const float *__restrict__ p, float *__restrict__ r, const float *__restrict__ ap,
                 double *__restrict__ partials, unsigned int *__restrict__ counter, double *__restrict__ red, PeerDev pd, int adv)
    {
      if (st->done) return;
      const float alpha = (float)st->alpha;
      double acc[1] = {0.0};
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        {
          x[i] = fmaf(alpha, p[i], x[i]);
          const float rc = fmaf(-alpha, ap[i], r[i]);
          r[i] = rc;
          acc[0] += (double)rc * (double)rc;
        }
      if (finish_reduce<1>(acc, partials, counter, red, pd) && adv && threadIdx.x == 0) cg_advance(kCXR, st, acc);
    }

    // p = r + beta p
    __global__ void __launch_bounds__(kT) cg_p_kernel(int n_pad, const CgState *__restrict__ st, const float *__restrict__ r, float *__restrict__ p)
    {
      if (st->done) return;
      const float beta = (float)st->beta;
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) p[i] = fmaf(beta, p[i], r[i]);
    }
  } // namespace

  // ---------------------------------------------------------------------------
  // Sell32
  // ---------------------------------------------------------------------------
  void Sell32::build(Context &ctx, const Bcsr &A, const NodeTable &nodes, const Halo *halo_, int precision_)
  {
    if (A.R != A.C || A.R < 1 || A.R > 3) throw std::runtime_error("Sell32: square blocks of size 1..3 required");
    if (precision_ != 32 && precision_ != 16) throw std::runtime_error("Sell32: precision must be 32 or 16");
    precision = precision_;
    const int dim = nodes.dim;
    if (const char *e = std::getenv("IFEM_SELL_VARIANT")) variant = std::atoi(e);
    bs = A.R;
    n_rows = A.n_brows_spmv >= 0 ? A.n_brows_spmv : A.n_brows;
    n_cols = A.n_bcols;
    if (n_rows > nodes.n_nodes || n_cols < n_rows) throw std::runtime_error("Sell32: node table does not match the matrix");
    const int n = n_rows;
    const std::vector<int64_t> rp = A.rowptr.to_host(ctx.stream);

    // ---- row order: columns of T x T nodes along the sweep axis, plane by plane, rows of one length together ----
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d) lo[d] = hi[d] = n ? nodes.coords[d] : 0.0;
    for (int i = 0; i < n; ++i)
      for (int d = 0; d < dim; ++d)
        {
          const double c = nodes.coords[(size_t)i * dim + d];
          lo[d] = std::min(lo[d], c);
          hi[d] = std::max(hi[d], c);
        }
    double vol = 1.0;
    int n_ext = 0;
    for (int d = 0; d < dim; ++d)
      if (hi[d] > lo[d])
        {
          vol *= hi[d] - lo[d];
          ++n_ext;
        }
    const double h = n_ext && n > 1 ? std::pow(vol / n, 1.0 / n_ext) : 1.0; // node spacing estimate
    const int T = dim == 3 ? 32 : 128;
    const int sweep = dim - 1;
    int nt[2] = {1, 1};
    for (int d = 0; d < sweep; ++d) nt[d] = std::max(1, std::min(1023, (int)std::ceil((hi[d] - lo[d]) / (T * h))));
    std::vector<std::pair<uint64_t, int>> keys((size_t)n);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i)
      {
        uint64_t tile = 0;
        for (int d = sweep - 1; d >= 0; --d)
          {
            const int b = std::max(0, std::min(nt[d] - 1, (int)std::floor((nodes.coords[(size_t)i * dim + d] - lo[d]) / (T * h))));
            tile = tile * (uint64_t)nt[d] + (uint64_t)b;
          }
        const int64_t zb = std::max<int64_t>(0, std::min<int64_t>((1 << 20) - 1, (int64_t)std::floor((nodes.coords[(size_t)i * dim + sweep] - lo[sweep]) / h + 0.5)));
        const int64_t len = std::min<int64_t>(rp[i + 1] - rp[i], 4095);
        keys[i] = {(tile << 32) | ((uint64_t)zb << 12) | (uint64_t)len, i};
      }
    std::sort(keys.begin(), keys.end());

    n_slices = (n + 31) / 32;
    n_pad = n_slices * 32;
    std::vector<int> perm((size_t)n_pad, -1), off((size_t)n_slices + 1, 0), ho((size_t)n_slices + 1, 0);
    h_pos.assign((size_t)n_cols, 0);
    n_blocks = 0;
    for (int i = 0; i < n; ++i)
      {
        perm[i] = keys[i].second;
        h_pos[keys[i].second] = i;
        n_blocks += rp[keys[i].second + 1] - rp[keys[i].second];
      }
    for (int g = n; g < n_cols; ++g) h_pos[g] = n_pad + (g - n);
    int64_t slots = 0;
    for (int sl = 0; sl < n_slices; ++sl)
      {
        int64_t L = 0;
        for (int l = 0; l < 32; ++l)
          {
            const int row = perm[(size_t)sl * 32 + l];
            if (row >= 0) L = std::max<int64_t>(L, rp[row + 1] - rp[row]);
          }
        slots += L;
        if (slots > INT32_MAX) throw std::runtime_error("Sell32: too many block slots for 32-bit slice offsets");
        off[sl + 1] = (int)slots;
        ho[sl + 1] = ho[sl] + (int)((L + 1) / 2);
      }
    n_slots = slots;
    n_hslots = ho[n_slices];
    keys.clear();
    keys.shrink_to_fit();

    slice_off.upload(off, ctx.stream);
    perm_row.upload(perm, ctx.stream);
    pos.upload(h_pos, ctx.stream);
    const int wgrid = (n_slices + kT / 32 - 1) / (kT / 32);
    col.release();
    val.release();
    col2.release();
    valh.release();
    row_scale.release();
    if (precision == 32)
      {
        col.alloc((size_t)n_slots * 32 + 32 * 8); // slack: the pipelined kernel prefetches up to 8 slots past a slice
        col.zero(ctx.stream);
        val.alloc((size_t)n_slots * 32 * bs * bs);
        if (n_slices)
          {
            sell_fill_col_kernel<<<wgrid, kT, 0, ctx.stream>>>(n_slices, slice_off.p, perm_row.p, A.rowptr.p, A.col.p, pos.p, col.p);
            IFEM_KERNEL_CHECK();
            ctx.kernel_launches++;
          }
      }
    else
      {
        hoff.upload(ho, ctx.stream);
        col2.alloc(((size_t)n_hslots * 32 + 32 * 8) * 2);
        col2.zero(ctx.stream);
        valh.alloc((size_t)n_hslots * 32 * bs * bs);
        row_scale.alloc((size_t)n_pad * bs);
        if (n_slices)
          {
            sell_fill_col2_kernel<<<wgrid, kT, 0, ctx.stream>>>(n_slices, hoff.p, perm_row.p, A.rowptr.p, A.col.p, pos.p,
                                                              reinterpret_cast<int2 *>(col2.p));
            IFEM_KERNEL_CHECK();
            ctx.kernel_launches++;
          }
      }

    // ---- halo plan in SELL numbering ----
    halo_plan = (halo_ && halo_->active()) ? halo_ : nullptr;
    if (halo_plan)
      {
        if (halo_plan->n_owned != n_rows) throw std::runtime_error("Sell32: halo plan does not match the matrix rows");
        if (halo_plan->n_send_total)
          {
            const std::vector<int> idx = halo_plan->d_send_idx.to_host(ctx.stream);
            std::vector<int> sp(idx.size());
            for (size_t k = 0; k < idx.size(); ++k) sp[k] = h_pos[idx[k]];
            send_pos.upload(sp, ctx.stream);
            send_buf.alloc((size_t)halo_plan->n_send_total * xs());
          }
      }
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
    setup_peer_halo(ctx);
  }

  // Collective over the ranks: where does my message land in each neighbour's gather source, and the arrival flags.
  void Sell32::setup_peer_halo(Context &ctx)
  {
    const bool was = peer_halo;
    peer_halo = false;
    if (!ctx.comm || ctx.comm->size < 2) return;
    PeerLink &link = peer_link(ctx);
    if (!link.active || !(link.mask & 2)) return;
    const int size = link.size;
    // row of this rank: for every rank s the float offsets of the segments that receive s's messages, in message order
    // (the halo plan has one entry per ghost layer and neighbour; as with ncclSend / ncclRecv an empty side of an entry is
    // skipped, so the j-th non-empty send of s to this rank meets its j-th non-empty receive from s), then an ok flag
    std::vector<int64_t> row((size_t)size * kPeerMaxMsgs + 1, -1);
    bool ok = halo_plan != nullptr;
    if (ok)
      {
        std::vector<int> n_recv((size_t)size, 0), n_send((size_t)size, 0), distinct;
        int n_msgs = 0;
        for (size_t k = 0; k < halo_plan->neighbours.size(); ++k)
          {
            const int s2 = halo_plan->neighbours[k];
            if (std::find(distinct.begin(), distinct.end(), s2) == distinct.end()) distinct.push_back(s2);
            if (halo_plan->recv_cnt[k] > 0)
              {
                if (n_recv[s2] >= kPeerMaxMsgs) { ok = false; break; }
                row[(size_t)s2 * kPeerMaxMsgs + n_recv[s2]++] = ((int64_t)n_pad + (int64_t)(halo_plan->recv_off[k] - n_rows)) * xs();
              }
            if (halo_plan->send_cnt[k] > 0) ++n_msgs;
          }
        ok = ok && (int)distinct.size() <= kPeerMaxNeighbours && n_msgs <= kPeerMaxMsgs;
      }
    row[(size_t)size * kPeerMaxMsgs] = ok ? 1 : 0;
    const std::vector<int64_t> all = comm_allgather_i64(ctx, row);
    const size_t stride = (size_t)size * kPeerMaxMsgs + 1;
    ghost_off.assign((size_t)size * size * kPeerMaxMsgs, -1);
    bool all_ok = true;
    for (int r = 0; r < size; ++r)
      {
        all_ok = all_ok && all[(size_t)r * stride + (size_t)size * kPeerMaxMsgs] == 1;
        for (size_t e = 0; e < (size_t)size * kPeerMaxMsgs; ++e) ghost_off[(size_t)r * size * kPeerMaxMsgs + e] = all[(size_t)r * stride + e];
      }
    if (!all_ok) return;
    if (!was || flag_peers.empty())
      {
        flag_peers = link.alloc_shared(ctx, sizeof(unsigned int) * kPeerMaxRanks);
        if (flag_peers.empty()) return;
        halo_state.alloc(2);
        halo_state.zero(ctx.stream);
        IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
      }
    peer_halo = true;
  }

  void Sell32::refresh(Context &ctx, const Bcsr &A)
  {
    if (!built()) throw std::runtime_error("Sell32::refresh before build");
    const int wgrid = (n_slices + kT / 32 - 1) / (kT / 32);
    if (precision == 16)
      {
        __half2 *vh = reinterpret_cast<__half2 *>(valh.p);
#define IFEM_FILL_H(B) sell_fill_valh_kernel<B><<<wgrid, kT, 0, ctx.stream>>>(n_slices, hoff.p, perm_row.p, A.rowptr.p, A.val.p, vh, row_scale.p)
        if (bs == 3) IFEM_FILL_H(3); else if (bs == 2) IFEM_FILL_H(2); else IFEM_FILL_H(1);
#undef IFEM_FILL_H
      }
    else
      {
#define IFEM_FILL(B) sell_fill_val_kernel<B><<<wgrid, kT, 0, ctx.stream>>>(n_slices, slice_off.p, perm_row.p, A.rowptr.p, A.val.p, val.p)
        if (bs == 3) IFEM_FILL(3); else if (bs == 2) IFEM_FILL(2); else IFEM_FILL(1);
#undef IFEM_FILL
      }
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }

  void Sell32::apply(Context &ctx, const float *x, float *y, const int *skip) const
  {
    const int wgrid = (n_slices + kT / 32 - 1) / (kT / 32);
    // variant = 10 * NS + MINB: NS slots per step, MINB resident CTAs per SM
    if (precision == 16)
      {
        const int2 *c2 = reinterpret_cast<const int2 *>(col2.p);
        const __half2 *vh = reinterpret_cast<const __half2 *>(valh.p);
#define IFEM_SELL_H(B, NS, M) sell_spmv_h_kernel<B, NS, M><<<wgrid, kT, 0, ctx.stream>>>(n_slices, hoff.p, c2, vh, row_scale.p, x, y, skip)
        if (bs == 3)
          switch (variant)
            {
            case 13: IFEM_SELL_H(3, 1, 3); break;
            case 16: IFEM_SELL_H(3, 1, 6); break;
            case 23: IFEM_SELL_H(3, 2, 3); break;
            case 26: IFEM_SELL_H(3, 2, 6); break;
            case 43: IFEM_SELL_H(3, 4, 3); break;
            default: IFEM_SELL_H(3, 2, 4); break; // 24
            }
        else if (bs == 2)
          IFEM_SELL_H(2, 2, 4);
        else
          switch (variant)
            {
            case 26: IFEM_SELL_H(1, 2, 6); break;
            case 44: IFEM_SELL_H(1, 4, 4); break;
            case 84: IFEM_SELL_H(1, 8, 4); break;
            default: IFEM_SELL_H(1, 4, 6); break;
            }
#undef IFEM_SELL_H
      }
    else
      {
#define IFEM_SELL_PIPE(B, NS, M) sell_spmv_pipe_kernel<B, NS, M><<<wgrid, kT, 0, ctx.stream>>>(n_slices, slice_off.p, col.p, val.p, x, y, skip)
        if (bs == 3)
          switch (variant)
            {
            case 13: IFEM_SELL_PIPE(3, 1, 3); break;
            case 16: IFEM_SELL_PIPE(3, 1, 6); break;
            case 23: IFEM_SELL_PIPE(3, 2, 3); break;
            case 26: IFEM_SELL_PIPE(3, 2, 6); break;
            case 43: IFEM_SELL_PIPE(3, 4, 3); break;
            default: IFEM_SELL_PIPE(3, 2, 4); break; // 24
            }
        else if (bs == 2)
          IFEM_SELL_PIPE(2, 2, 4);
        else
          switch (variant)
            {
            case 26: IFEM_SELL_PIPE(1, 2, 6); break;
            case 44: IFEM_SELL_PIPE(1, 4, 4); break;
            case 84: IFEM_SELL_PIPE(1, 8, 4); break;
            default: IFEM_SELL_PIPE(1, 4, 6); break;
            }
#undef IFEM_SELL_PIPE
      }
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }

  float *Sell32::gather_source(Context &ctx, int slot)
  {
    if (slot < 0 || slot >= 4) throw std::runtime_error("Sell32::gather_source: slot out of range");
    const size_t len = x_len();
    if (peer_halo)
      {
        // collective: every rank asks for its sources in the same order
        if ((int)sources.size() <= slot) sources.resize(slot + 1);
        Source &src = sources[slot];
        if (!src.local)
          {
            PeerLink &link = peer_link(ctx);
            // one length for all ranks (the mapping is by base address; the ghost offsets are per rank)
            int64_t max_len = 0;
            for (int64_t v : comm_allgather_i64(ctx, {(int64_t)len})) max_len = std::max(max_len, v);
            src.peers = link.alloc_shared(ctx, (size_t)max_len * sizeof(float));
            if (src.peers.empty())
              peer_halo = false; // every rank reached the same verdict inside alloc_shared
            else
              src.local = static_cast<float *>(src.peers[link.rank]);
          }
        if (src.local)
          {
            IFEM_CUDA(cudaMemsetAsync(src.local, 0, len * sizeof(float), ctx.stream));
            return src.local;
          }
      }
    plain_sources[slot].alloc(len);
    plain_sources[slot].zero(ctx.stream);
    return plain_sources[slot].p;
  }

  void Sell32::halo(Context &ctx, float *x, const int *skip)
  {
    if (!halo_plan) return;
    if (!ctx.comm) throw std::runtime_error("Sell32::halo: no communicator");
    const Halo &H = *halo_plan;
    const int w = xs();
    if (peer_halo)
      for (const Source &src : sources)
        if (src.local == x)
          {
            PeerLink &link = peer_link(ctx);
            PeerHaloDev<float> h;
            h.width = w;
            std::vector<int> sent((size_t)link.size, 0);
            for (size_t k = 0; k < H.neighbours.size(); ++k)
              {
                const int nb = H.neighbours[k];
                bool seen = false;
                for (int j = 0; j < h.n_nb; ++j) seen = seen || h.nb_rank[j] == nb;
                if (!seen)
                  {
                    h.flag[h.n_nb] = static_cast<unsigned int *>(flag_peers[nb]) + link.rank;
                    h.nb_rank[h.n_nb++] = nb;
                  }
                if (H.send_cnt[k] <= 0) continue;
                // my j-th non-empty message to nb lands in nb's j-th non-empty receive segment from me
                const int64_t off = ghost_off[((size_t)nb * link.size + link.rank) * kPeerMaxMsgs + sent[nb]++];
                if (off < 0) throw std::runtime_error("Sell32::halo: the neighbour has no receive segment for this message");
                h.send_off[h.n_msg] = H.send_off[k];
                h.dst[h.n_msg++] = static_cast<float *>(src.peers[nb]) + off;
              }
            h.send_off[h.n_msg] = H.n_send_total;
            h.my_flags = static_cast<const unsigned int *>(flag_peers[link.rank]);
            h.epoch = halo_state.p;
            h.counter = halo_state.p + 1;
            const int total = H.n_send_total * w;
            const int grid = std::max(1, std::min((total + kT * 8 - 1) / (kT * 8), ctx.sm_count));
            peer_halo_push_kernel<float><<<grid, kT, 0, ctx.stream>>>(h, send_pos.p, x, skip);
            IFEM_KERNEL_CHECK();
            peer_halo_wait_kernel<float><<<1, 32, 0, ctx.stream>>>(h, skip);
            IFEM_KERNEL_CHECK();
            ctx.kernel_launches += 2;
            return;
          }
    if (H.n_send_total)
      {
        const int total = H.n_send_total * w;
        halo_pack32_kernel<<<(total + kT - 1) / kT, kT, 0, ctx.stream>>>(H.n_send_total, w, send_pos.p, x, send_buf.p);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
    comm_group_start(*ctx.comm);
    for (size_t k = 0; k < H.neighbours.size(); ++k)
      comm_sendrecv_f32(*ctx.comm, H.neighbours[k], send_buf.p + (size_t)H.send_off[k] * w, (int64_t)H.send_cnt[k] * w,
                        x + ((size_t)n_pad + (size_t)(H.recv_off[k] - n_rows)) * w, (int64_t)H.recv_cnt[k] * w, ctx.stream);
    comm_group_end(*ctx.comm);
  }

  // ---------------------------------------------------------------------------
  // InnerSolver32: BiCGStab + node-block Jacobi
  // ---------------------------------------------------------------------------
  InnerSolver32::~InnerSolver32()
  {
    if (h_state) cudaFreeHost(h_state);
  }

  void InnerSolver32::setup(Context &ctx, const Bcsr &A, const NodeTable &nodes, const Halo *halo_, int precision)
  {
    if (A.R != 2 && A.R != 3) throw std::runtime_error("InnerSolver32: 2x2 / 3x3 blocks required");
    S.build(ctx, A, nodes, halo_, precision);
    const size_t nv = (size_t)S.n_pad * S.bs;
    for (DevBuf<float> *b : {&r, &r0, &p, &v, &s, &t, &x})
      {
        b->alloc(nv);
        b->zero(ctx.stream);
      }
    ph = S.gather_source(ctx, 0);
    sh = S.gather_source(ctx, 1);
    binv.alloc((size_t)S.bs * S.bs * S.n_pad);
    grid = std::max(1, std::min((S.n_pad + kT - 1) / kT, ctx.sm_count * 8));
    partials.alloc((size_t)grid * 2);
    red.alloc(kPeerMaxVals);
    red.zero(ctx.stream);
    counter.alloc(1);
    counter.zero(ctx.stream);
    state.alloc((sizeof(BicgState) + sizeof(int) - 1) / sizeof(int));
    state.zero(ctx.stream);
    if (!h_state) IFEM_CUDA(cudaMallocHost(&h_state, sizeof(BicgState)));
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void InnerSolver32::refresh(Context &ctx, const Bcsr &A, const double *binv64)
  {
    S.refresh(ctx, A);
    if (binv64)
      {
        binv_to_sell_kernel<<<grid, kT, 0, ctx.stream>>>(S.n_pad, S.bs * S.bs, S.perm_row.p, binv64, binv.p);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
  }

  template <int BS>
  SolveResult InnerSolver32::solve_impl(Context &ctx, const double *src, double src_norm, double *dst, double rel_tol, int max_it)
  {
    SolveResult out;
    const int n_pad = S.n_pad;
    const int64_t nflat = (int64_t)n_pad * BS;
    float4 *ph4 = reinterpret_cast<float4 *>(ph), *sh4 = reinterpret_cast<float4 *>(sh);
    BicgState *st = reinterpret_cast<BicgState *>(state.p);
    const BicgState *h = static_cast<const BicgState *>(h_state);
    auto launched = [&](int k = 1) {
      IFEM_KERNEL_CHECK();
      ctx.kernel_launches += k;
    };
    if (!(src_norm > 0.0))
      {
        // A^-1 0 = 0
        final_kernel<BS><<<grid, kT, 0, ctx.stream>>>(n_pad, S.perm_row.p, x.p, 0.0, dst);
        launched();
        out.converged = true;
        return out;
      }
    const ReduceMode m = reduce_mode(ctx);
    // several ranks without a peer link: the sums of the kernel are local; NCCL adds them up, then the state advances
    auto after = [&](int stage, int nr) {
      if (!m.nccl) return;
      comm_allreduce_sum(*ctx.comm, red.p, nr, ctx.stream);
      bicg_advance_kernel<<<1, 1, 0, ctx.stream>>>(stage, st, red.p);
      launched();
    };
    bicg_begin_kernel<<<1, 1, 0, ctx.stream>>>(st, rel_tol * rel_tol, max_it);
    launched();
    init_kernel<BS><<<grid, kT, 0, ctx.stream>>>(n_pad, S.perm_row.p, src, 1.0 / src_norm, r.p, r0.p, p.p, v.p, x.p, st, partials.p, counter.p,
                                                 red.p, m.pd, m.adv);
    launched();
    after(kBInit, 1);
    int enqueued = 0;
    while (true)
      {
        const int chunk = std::max(1, std::min(check_every, max_it - enqueued));
        for (int k = 0; k < chunk; ++k)
          {
            update_p_kernel<BS><<<grid, kT, 0, ctx.stream>>>(n_pad, st, r.p, p.p, v.p, binv.p, ph4);
            launched();
            S.halo(ctx, ph, &st->done);
            S.apply(ctx, ph, v.p, &st->done);
            dot1_kernel<<<grid, kT, 0, ctx.stream>>>(nflat, r0.p, v.p, st, partials.p, counter.p, red.p, m.pd, m.adv);
            launched();
            after(kBDot1, 1);
            update_s_kernel<BS><<<grid, kT, 0, ctx.stream>>>(n_pad, st, r.p, v.p, s.p, binv.p, sh4, partials.p, counter.p, red.p, m.pd, m.adv);
            launched();
            after(kBS, 1);
            S.halo(ctx, sh, &st->skip2);
            S.apply(ctx, sh, t.p, &st->skip2);
            dot2_kernel<<<grid, kT, 0, ctx.stream>>>(nflat, t.p, s.p, st, partials.p, counter.p, red.p, m.pd, m.adv);
            launched();
            after(kBDot2, 2);
            update_xr_kernel<BS><<<grid, kT, 0, ctx.stream>>>(n_pad, st, x.p, ph4, sh4, r.p, s.p, t.p, r0.p, partials.p, counter.p, red.p, m.pd,
                                                              m.adv);
            launched();
            after(kBXR, 2);
          }
        enqueued += chunk;
        IFEM_CUDA(cudaMemcpyAsync(h_state, state.p, sizeof(BicgState), cudaMemcpyDeviceToHost, ctx.stream));
        IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
        if (h->done || enqueued >= max_it) break;
      }
    out.iterations = h->its;
    out.residual = std::sqrt(std::max(0.0, h->res2));
    out.converged = h->converged != 0;
    static const bool debug = std::getenv("IFEM_INNER_DEBUG") != nullptr;
    if (debug)
      std::fprintf(stderr, "[inner32] bicgstab its %d rel.res %.3e converged %d breakdown %d\n", out.iterations, out.residual, (int)out.converged,
                   h->breakdown);
    if (!std::isfinite(out.residual) || out.residual >= 1.0)
      {
        // no progress over x = 0 (breakdown in fp32): hand back one block-Jacobi step D^-1 src, which is always a
        // valid (if weak) preconditioner application for the flexible outer iteration. r0 still holds src / |src|.
        jacobi_out_kernel<BS><<<grid, kT, 0, ctx.stream>>>(n_pad, S.perm_row.p, r0.p, binv.p, src_norm, dst);
        launched();
        out.residual = src_norm;
        out.converged = false;
        n_fallbacks++;
        return out;
      }
    final_kernel<BS><<<grid, kT, 0, ctx.stream>>>(n_pad, S.perm_row.p, x.p, src_norm, dst);
    launched();
    out.residual *= src_norm;
    return out;
  }

  SolveResult InnerSolver32::solve(Context &ctx, const double *src, double src_norm, double *dst, double rel_tol, int max_it)
  {
    if (!S.built()) throw std::runtime_error("InnerSolver32::solve before setup");
    if (S.bs == 3) return solve_impl<3>(ctx, src, src_norm, dst, rel_tol, max_it);
    return solve_impl<2>(ctx, src, src_norm, dst, rel_tol, max_it);
  }

  void InnerSolver32::probe_load(Context &ctx, const double *xin)
  {
    if (!S.built()) throw std::runtime_error("InnerSolver32::probe_load before setup");
    float4 *ph4 = reinterpret_cast<float4 *>(ph);
    if (S.bs == 3)
      to_sell4_kernel<3><<<grid, kT, 0, ctx.stream>>>(S.n_pad, S.perm_row.p, xin, ph4);
    else
      to_sell4_kernel<2><<<grid, kT, 0, ctx.stream>>>(S.n_pad, S.perm_row.p, xin, ph4);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
    S.halo(ctx, ph);
  }

  void InnerSolver32::probe_apply(Context &ctx) { S.apply(ctx, ph, v.p); }

  void InnerSolver32::probe_store(Context &ctx, double *yout)
  {
    if (S.bs == 3)
      final_kernel<3><<<grid, kT, 0, ctx.stream>>>(S.n_pad, S.perm_row.p, v.p, 1.0, yout);
    else
      final_kernel<2><<<grid, kT, 0, ctx.stream>>>(S.n_pad, S.perm_row.p, v.p, 1.0, yout);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }

  // ---------------------------------------------------------------------------
  // InnerCG32
  // ---------------------------------------------------------------------------
  InnerCG32::~InnerCG32()
  {
    if (h_state) cudaFreeHost(h_state);
  }

  void InnerCG32::setup(Context &ctx, const Bcsr &A, const NodeTable &nodes, const Halo *halo_, int precision)
  {
    if (A.R != 1) throw std::runtime_error("InnerCG32: scalar matrix required");
    S.build(ctx, A, nodes, halo_, precision);
    for (DevBuf<float> *b : {&r, &ap, &x, &pg, &sg})
      {
        b->alloc((size_t)S.n_pad);
        b->zero(ctx.stream);
      }
    p = S.gather_source(ctx, 0);
    rg = S.gather_source(ctx, 1);
    if (const char *e = std::getenv("IFEM_CG_SM_GEAR")) single_reduction = std::atoi(e) != 0;
    if (const char *e = std::getenv("IFEM_CG_SM_COARSE")) coarse_space = std::atoi(e) != 0;
    dim = nodes.dim;
    n_coarse = 0;
    grid = std::max(1, std::min((S.n_pad + kT - 1) / kT, ctx.sm_count * 4));
    partials.alloc((size_t)grid * 3);
    red.alloc(kPeerMaxVals);
    red.zero(ctx.stream);
    counter.alloc(1);
    counter.zero(ctx.stream);
    state.alloc((sizeof(CgState) + sizeof(int) - 1) / sizeof(int));
    state.zero(ctx.stream);
    if (!h_state) IFEM_CUDA(cudaMallocHost(&h_state, sizeof(CgState)));
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  SolveResult InnerCG32::solve(Context &ctx, const double *src, double src_norm, double *dst, double tol_abs, int max_it)
  {
    if (!S.built()) throw std::runtime_error("InnerCG32::solve before setup");
    SolveResult out;
    const int n_pad = S.n_pad;
    CgState *st = reinterpret_cast<CgState *>(state.p);
    const CgState *h = static_cast<const CgState *>(h_state);
    auto launched = [&](int k = 1) {
      IFEM_KERNEL_CHECK();
      ctx.kernel_launches += k;
    };
    if (!(src_norm > 0.0))
      {
        final_kernel<1><<<grid, kT, 0, ctx.stream>>>(n_pad, S.perm_row.p, x.p, 0.0, dst);
        launched();
        out.converged = true;
        return out;
      }
    const ReduceMode m = reduce_mode(ctx);
    auto after = [&](int stage) {
      if (!m.nccl) return;
      comm_allreduce_sum(*ctx.comm, red.p, 1, ctx.stream);
      cg_advance_kernel<<<1, 1, 0, ctx.stream>>>(stage, st, red.p);
      launched();
    };
    const double tol_rel = tol_abs / src_norm;
    cg_begin_kernel<<<1, 1, 0, ctx.stream>>>(st, tol_rel * tol_rel, max_it);
    launched();
    if (coarse_space && n_coarse > 0)
      {
        // r in `r`, u = M^-1 r in the gather source `rg` (what the product reads), w = A u in `ap`
        cg_gear_init_kernel<<<grid, kT, 0, ctx.stream>>>(n_pad, S.perm_row.p, src, 1.0 / src_norm, r.p, pg.p, sg.p, x.p);
        launched();
        const bool many = ctx.comm && ctx.comm->size > 1;
        int enqueued = 0;
        while (true)
          {
            const int chunk = std::max(1, std::min(check_every, max_it + 1 - enqueued));
            for (int k = 0; k < chunk; ++k)
              {
                pcg_restrict_kernel<<<n_coarse, kT, 0, ctx.stream>>>(st, agg_ptr.p, agg_rows.p, r.p, cvec.p);
                launched();
                if (many) comm_allreduce_sum(*ctx.comm, cvec.p, n_coarse, ctx.stream);
                pcg_coarse_kernel<<<(n_coarse * 32 + kT - 1) / kT, kT, 0, ctx.stream>>>(st, n_coarse, Einv.p, cvec.p, yvec.p);
                pcg_build_u_kernel<<<grid, kT, 0, ctx.stream>>>(n_pad, st, r.p, dinv.p, agg.p, yvec.p, rg);
                launched(2);
                S.halo(ctx, rg, &st->done);
                S.apply(ctx, rg, ap.p, &st->done);
                pcg_dot_kernel<<<grid, kT, 0, ctx.stream>>>(n_pad, r.p, rg, ap.p, st, partials.p, counter.p, red.p, m.pd, m.adv);
                launched();
                if (m.nccl)
                  {
                    comm_allreduce_sum(*ctx.comm, red.p, 3, ctx.stream);
                    pcg_advance_kernel<<<1, 1, 0, ctx.stream>>>(st, red.p);
                    launched();
                  }
                pcg_update_kernel<<<grid, kT, 0, ctx.stream>>>(n_pad, st, r.p, rg, ap.p, pg.p, sg.p, x.p);
                launched();
              }
            enqueued += chunk;
            IFEM_CUDA(cudaMemcpyAsync(h_state, state.p, sizeof(CgState), cudaMemcpyDeviceToHost, ctx.stream));
            IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
            if (h->done || enqueued > max_it) break;
          }
        out.iterations = h->its;
        out.residual = std::sqrt(std::max(0.0, h->rr)) * src_norm;
        out.converged = h->done && h->converged;
        final_kernel<1><<<grid, kT, 0, ctx.stream>>>(n_pad, S.perm_row.p, x.p, src_norm, dst);
        launched();
        return out;
      }
    if (single_reduction)
      {
        // w lives in `ap`, r in the gather source `rg` (it is what the product reads)
        cg_gear_init_kernel<<<grid, kT, 0, ctx.stream>>>(n_pad, S.perm_row.p, src, 1.0 / src_norm, rg, pg.p, sg.p, x.p);
        launched();
        int enqueued = 0;
        while (true)
          {
            const int chunk = std::max(1, std::min(check_every, max_it + 1 - enqueued));
            for (int k = 0; k < chunk; ++k)
              {
                S.halo(ctx, rg, &st->done);
                S.apply(ctx, rg, ap.p, &st->done);
                cg_gear_dot_kernel<<<grid, kT, 0, ctx.stream>>>(n_pad, rg, ap.p, st, partials.p, counter.p, red.p, m.pd, m.adv);
                launched();
                if (m.nccl)
                  {
                    comm_allreduce_sum(*ctx.comm, red.p, 2, ctx.stream);
                    cg_advance_kernel<<<1, 1, 0, ctx.stream>>>(kCGear, st, red.p);
                    launched();
                  }
                cg_gear_update_kernel<<<grid, kT, 0, ctx.stream>>>(n_pad, st, rg, ap.p, pg.p, sg.p, x.p);
                launched();
              }
            enqueued += chunk;
            IFEM_CUDA(cudaMemcpyAsync(h_state, state.p, sizeof(CgState), cudaMemcpyDeviceToHost, ctx.stream));
            IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
            if (h->done || enqueued > max_it) break;
          }
        out.iterations = h->its;
        out.residual = std::sqrt(std::max(0.0, h->rr)) * src_norm;
        out.converged = h->done && h->converged;
        final_kernel<1><<<grid, kT, 0, ctx.stream>>>(n_pad, S.perm_row.p, x.p, src_norm, dst);
        launched();
        return out;
      }
    cg_init_kernel<<<grid, kT, 0, ctx.stream>>>(n_pad, S.perm_row.p, src, 1.0 / src_norm, r.p, p, x.p, st, partials.p, counter.p, red.p, m.pd,
                                                m.adv);
    launched();
    after(kCInit);
    int enqueued = 0;
    while (true)
      {
        const int chunk = std::max(1, std::min(check_every, max_it - enqueued));
        for (int k = 0; k < chunk; ++k)
          {
            S.halo(ctx, p, &st->done);
            S.apply(ctx, p, ap.p, &st->done);
            cg_dot_kernel<<<grid, kT, 0, ctx.stream>>>(n_pad, p, ap.p, st, partials.p, counter.p, red.p, m.pd, m.adv);
            launched();
            after(kCDot);
            cg_xr_kernel<<<grid, kT, 0, ctx.stream>>>(n_pad, st, x.p, p, r.p, ap.p, partials.p, counter.p, red.p, m.pd, m.adv);
            launched();
            after(kCXR);
            cg_p_kernel<<<grid, kT, 0, ctx.stream>>>(n_pad, st, r.p, p);
            launched();
          }
        enqueued += chunk;
        IFEM_CUDA(cudaMemcpyAsync(h_state, state.p, sizeof(CgState), cudaMemcpyDeviceToHost, ctx.stream));
        IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
        if (h->done || enqueued >= max_it) break;
      }
    static const bool debug = std::getenv("IFEM_INNER_DEBUG") != nullptr;
    if (debug) std::fprintf(stderr, "[inner32] cg its %d rel.res %.3e done %d\n", h->its, std::sqrt(std::max(0.0, h->rr)), h->done);
    out.iterations = h->its;
    out.residual = std::sqrt(std::max(0.0, h->rr)) * src_norm;
    out.converged = h->rr <= h->tol2;
    final_kernel<1><<<grid, kT, 0, ctx.stream>>>(n_pad, S.perm_row.p, x.p, src_norm, dst);
    launched();
    return out;
  }
  void InnerCG32::build_coarse(Context &ctx, const Bcsr &A, const NodeTable &nodes)
  {
    n_coarse = 0;
    if (!coarse_space) return;
    const int n = S.n_rows, n_cols = S.n_cols, n_pad = S.n_pad;
    // boxes per direction: at most 729 aggregates (9^3 / 27^2), at least ~8 rows per aggregate on small problems; every rank must
    // come to the same G
    int64_t n_global = n;
    if (ctx.comm && ctx.comm->size > 1)
      {
        n_global = 0;
        for (int64_t k : comm_allgather_i64(ctx, std::vector<int64_t>{(int64_t)n})) n_global += k;
      }
    const int g_max = dim == 3 ? 9 : 27;
    const int G = std::max(2, std::min(g_max, (int)std::floor(std::pow((double)n_global / 8.0, 1.0 / dim))));
    const int n_c = dim == 3 ? G * G * G : G * G;
    // aggregate of every local node (owned and ghost) from its position in the global box
    std::vector<int> agg_node((size_t)n_cols);
    for (int i = 0; i < n_cols; ++i)
      {
        int a = 0, stride = 1;
        for (int d = 0; d < dim; ++d)
          {
            const double ext = box[2 * d + 1] - box[2 * d];
            int k = ext > 0 ? (int)std::floor((nodes.coords[(size_t)i * dim + d] - box[2 * d]) / ext * G) : 0;
            k = std::min(std::max(k, 0), G - 1);
            a += k * stride;
            stride *= G;
          }
        agg_node[i] = a;
      }
    // E = Z^T A Z over the owned rows, summed over the ranks
    std::vector<int64_t> rp;
    std::vector<int> ci;
    std::vector<double> v;
    A.to_host_csr(ctx.stream, rp, ci, v);
    std::vector<double> E((size_t)n_c * n_c, 0.0), diag((size_t)n, 1.0), count((size_t)n_c, 0.0);
    for (int i = 0; i < n; ++i)
      {
        double *Er = &E[(size_t)agg_node[i] * n_c];
        count[agg_node[i]] += 1.0;
        for (int64_t k = rp[i]; k < rp[i + 1]; ++k)
          {
            Er[agg_node[ci[k]]] += v[k];
            if (ci[k] == i) diag[i] = v[k];
          }
      }
    const bool many = ctx.comm && ctx.comm->size > 1;
    if (many)
      {
        DevBuf<double> dE(E.size() + count.size());
        dE.upload(E.data(), E.size(), ctx.stream);
        IFEM_CUDA(cudaMemcpyAsync(dE.p + E.size(), count.data(), count.size() * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
        comm_allreduce_sum(*ctx.comm, dE.p, (int)(E.size() + count.size()), ctx.stream);
        std::vector<double> back = dE.to_host(ctx.stream);
        std::copy(back.begin(), back.begin() + E.size(), E.begin());
        std::copy(back.begin() + E.size(), back.end(), count.begin());
      }
    // empty aggregates decouple; the constant null space of a singular S_m is shifted away
    double mean_diag = 0.0;
    int n_used = 0;
    for (int a = 0; a < n_c; ++a)
      if (count[a] > 0)
        {
          mean_diag += E[(size_t)a * n_c + a];
          ++n_used;
        }
    if (!n_used || !(mean_diag > 0.0)) return;
    mean_diag /= n_used;
    for (int a = 0; a < n_c; ++a)
      for (int b = 0; b < n_c; ++b)
        {
          double &e = E[(size_t)a * n_c + b];
          if (count[a] > 0 && count[b] > 0) e += mean_diag / n_used;
          else e = a == b ? 1.0 : 0.0;
        }
    // inverse by Gauss-Jordan with partial pivoting (n_c <= 729, once per matrix)
    std::vector<double> inv((size_t)n_c * n_c, 0.0);
    for (int a = 0; a < n_c; ++a) inv[(size_t)a * n_c + a] = 1.0;
    for (int c = 0; c < n_c; ++c)
      {
        int piv = c;
        for (int r2 = c + 1; r2 < n_c; ++r2)
          if (std::fabs(E[(size_t)r2 * n_c + c]) > std::fabs(E[(size_t)piv * n_c + c])) piv = r2;
        if (!(std::fabs(E[(size_t)piv * n_c + c]) > 1e-300)) return; // singular beyond the shift: no coarse space
        if (piv != c)
          for (int k = 0; k < n_c; ++k)
            {
              std::swap(E[(size_t)c * n_c + k], E[(size_t)piv * n_c + k]);
              std::swap(inv[(size_t)c * n_c + k], inv[(size_t)piv * n_c + k]);
            }
        const double d = 1.0 / E[(size_t)c * n_c + c];
        for (int k = 0; k < n_c; ++k)
          {
            E[(size_t)c * n_c + k] *= d;
            inv[(size_t)c * n_c + k] *= d;
          }
#pragma omp parallel for schedule(static)
        for (int r2 = 0; r2 < n_c; ++r2)
          {
            if (r2 == c) continue;
            const double f = E[(size_t)r2 * n_c + c];
            if (f == 0.0) continue;
            for (int k = 0; k < n_c; ++k)
              {
                E[(size_t)r2 * n_c + k] -= f * E[(size_t)c * n_c + k];
                inv[(size_t)r2 * n_c + k] -= f * inv[(size_t)c * n_c + k];
              }
          }
      }
    // SELL-ordered tables
    const std::vector<int> perm = S.perm_row.to_host(ctx.stream);
    std::vector<int> agg_h((size_t)n_pad, -1), ptr((size_t)n_c + 1, 0), rows;
    std::vector<float> dinv_h((size_t)n_pad, 0.0f);
    for (int i = 0; i < n_pad; ++i)
      if (perm[i] >= 0)
        {
          agg_h[i] = agg_node[perm[i]];
          dinv_h[i] = (float)(1.0 / diag[perm[i]]);
          ptr[agg_h[i] + 1]++;
        }
    for (int a = 0; a < n_c; ++a) ptr[a + 1] += ptr[a];
    rows.resize(ptr[n_c]);
    std::vector<int> cur(ptr.begin(), ptr.end() - 1);
    for (int i = 0; i < n_pad; ++i)
      if (agg_h[i] >= 0) rows[cur[agg_h[i]]++] = i;
    agg.upload(agg_h, ctx.stream);
    agg_ptr.upload(ptr, ctx.stream);
    if (rows.empty()) rows.push_back(0);
    agg_rows.upload(rows, ctx.stream);
    dinv.upload(dinv_h, ctx.stream);
    Einv.upload(inv, ctx.stream);
    cvec.alloc((size_t)n_c);
    yvec.alloc((size_t)n_c);
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
    n_coarse = n_c;
  }
} // namespace ifem
