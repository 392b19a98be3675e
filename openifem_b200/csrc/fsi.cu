#include "fsi.h"

#include "contact.h"

#include "comm.h"
#include "scnsim.h"

#include <algorithm>
#include <chrono>
#include <cmath>

namespace ifem
{
  namespace
  {
    constexpr double kUnitTol = 1e-10;  // find_active_cell_around_point tolerance on the unit cell
    constexpr double kBoxSlack = 1e-12; // slack of the per-cell bounding-box reject

    struct SolidView
    {
      int n_cells;
      const int *cells;   // [n_cells][2^dim] node ids
      const double *x;    // deformed vertices [n_nodes][dim]
      const double *box;  // [2*dim]
      int nbin[3];
      const int *bin_start, *bin_items;
      int n_bseg;
      const int *bseg;    // 2-D boundary segments (node pairs)
    };

    template <int DIM>
    __device__ __forceinline__ void q1_shape(const double *xi, double *N, double *dN)
    {
      constexpr int NV = 1 << DIM;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        {
          double n = 1.0, d[DIM];
#pragma unroll
          for (int e = 0; e < DIM; ++e) d[e] = 1.0;
#pragma unroll
          for (int k = 0; k < DIM; ++k)
            {
              const int bit = (v >> k) & 1;
              const double f = bit ? xi[k] : 1.0 - xi[k], df = bit ? 1.0 : -1.0;
              n *= f;
#pragma unroll
              for (int e = 0; e < DIM; ++e) d[e] *= (e == k) ? df : f;
            }
          N[v] = n;
#pragma unroll
          for (int e = 0; e < DIM; ++e) dN[v * DIM + e] = d[e];
        }
    }

    template <int DIM>
    __device__ __forceinline__ bool solve_small(const double *J, const double *r, double *dx)
    {
      if (DIM == 2)
        {
          const double det = J[0] * J[3] - J[1] * J[2];
          if (det == 0.0) return false;
          dx[0] = (J[3] * r[0] - J[1] * r[1]) / det;
          dx[1] = (-J[2] * r[0] + J[0] * r[1]) / det;
        }
      else
        {
          const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
          const double det = J[0] * c00 + J[1] * c01 + J[2] * c02;
          if (det == 0.0) return false;
          const double d = 1.0 / det;
          const double i0 = c00 * d, i1 = (J[2] * J[7] - J[1] * J[8]) * d, i2 = (J[1] * J[5] - J[2] * J[4]) * d;
          const double i3 = c01 * d, i4 = (J[0] * J[8] - J[2] * J[6]) * d, i5 = (J[2] * J[3] - J[0] * J[5]) * d;
          const double i6 = c02 * d, i7 = (J[1] * J[6] - J[0] * J[7]) * d, i8 = (J[0] * J[4] - J[1] * J[3]) * d;
          dx[0] = i0 * r[0] + i1 * r[1] + i2 * r[2];
          dx[1] = i3 * r[0] + i4 * r[1] + i5 * r[2];
          dx[2] = i6 * r[0] + i7 * r[1] + i8 * r[2];
        }
      return true;
    }

    // MappingQ1::transform_real_to_unit_cell + unit-cell test (CellAccessor::point_inside)
    template <int DIM>
    __device__ bool point_in_cell(const SolidView &S, int cell, const double *p, double *xi_out)
    {
      constexpr int NV = 1 << DIM;
      double V[NV * DIM], lo[DIM], hi[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        {
          lo[d] = 1e300;
          hi[d] = -1e300;
        }
#pragma unroll
      for (int v = 0; v < NV; ++v)
        {
          const double *X = S.x + (int64_t)S.cells[(int64_t)cell * NV + v] * DIM;
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              V[v * DIM + d] = X[d];
              lo[d] = fmin(lo[d], X[d]);
              hi[d] = fmax(hi[d], X[d]);
            }
        }
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        if (p[d] < lo[d] - kBoxSlack || p[d] > hi[d] + kBoxSlack) return false;
      double xi[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) xi[d] = 0.5;
      bool converged = false;
      for (int it = 0; it < 30 && !converged; ++it)
        {
          double N[NV], dN[NV * DIM], r[DIM], J[DIM * DIM], dx[DIM];
          q1_shape<DIM>(xi, N, dN);
#pragma unroll
          for (int i = 0; i < DIM; ++i)
            {
              double s = -p[i];
#pragma unroll
              for (int v = 0; v < NV; ++v) s += N[v] * V[v * DIM + i];
              r[i] = s;
#pragma unroll
              for (int j = 0; j < DIM; ++j)
                {
                  double t = 0.0;
#pragma unroll
                  for (int v = 0; v < NV; ++v) t += V[v * DIM + i] * dN[v * DIM + j];
                  J[i * DIM + j] = t;
                }
            }
          if (!solve_small<DIM>(J, r, dx)) return false;
          double n2 = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              xi[d] -= dx[d];
              n2 += dx[d] * dx[d];
            }
          converged = sqrt(n2) < 1e-13;
        }
      if (!converged) return false;
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        if (xi[d] < -kUnitTol || xi[d] > 1.0 + kUnitTol) return false;
#pragma unroll
      for (int d = 0; d < DIM; ++d) xi_out[d] = fmin(fmax(xi[d], 0.0), 1.0); // GeometryInfo::project_to_unit_cell
      return true;
    }

    template <int DIM>
    __device__ __forceinline__ bool in_box(const SolidView &S, const double *p)
    {
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        if (p[d] < S.box[2 * d] || p[d] > S.box[2 * d + 1]) return false;
      return true;
    }

    template <int DIM>
    __device__ __forceinline__ int bin_of(const SolidView &S, const double *p)
    {
      int b = 0, stride = 1;
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        {
          const double ext = S.box[2 * d + 1] - S.box[2 * d];
          int k = ext > 0 ? (int)floor((p[d] - S.box[2 * d]) / ext * S.nbin[d]) : 0;
          k = min(max(k, 0), S.nbin[d] - 1);
          b += k * stride;
          stride *= S.nbin[d];
        }
      return b;
    }

    // lowest-index solid cell containing p (CellLocator / find_active_cell_around_point), -1 if none
    template <int DIM>
    __device__ int locate(const SolidView &S, const double *p, double *xi)
    {
      if (!in_box<DIM>(S, p)) return -1;
      const int b = bin_of<DIM>(S, p);
      int best = -1;
      for (int k = S.bin_start[b]; k < S.bin_start[b + 1]; ++k)
        {
          const int c = S.bin_items[k];
          if (best >= 0 && c > best) continue;
          double t[DIM];
          if (point_in_cell<DIM>(S, c, p, t))
            {
              best = c;
#pragma unroll
              for (int d = 0; d < DIM; ++d) xi[d] = t[d];
            }
        }
      return best;
    }

    // FSI::point_in_solid (mpi_fsi.cpp:143-224). 2-D: crossing number, statement for statement, with
    // round-to-nearest intrinsics so that no FMA contraction changes the exact comparisons.
    template <int DIM>
    __device__ bool point_in_solid(const SolidView &S, const double *p)
    {
      if (!in_box<DIM>(S, p)) return false;
      if (DIM == 3)
        {
          double xi[DIM];
          return locate<DIM>(S, p, xi) >= 0;
        }
      unsigned int cross_number = 0, half_cross_number = 0;
      for (int f = 0; f < S.n_bseg; ++f)
        {
          const double *p1 = S.x + (int64_t)S.bseg[2 * f] * DIM, *p2 = S.x + (int64_t)S.bseg[2 * f + 1] * DIM;
          const double y_diff1 = __dsub_rn(p1[1], p[1]), y_diff2 = __dsub_rn(p2[1], p[1]);
          const double x_diff1 = __dsub_rn(p1[0], p[0]), x_diff2 = __dsub_rn(p2[0], p[0]);
          const double r1x = __dsub_rn(p1[0], p2[0]), r1y = __dsub_rn(p1[1], p2[1]);
          double r2x = 0.0;
          if (r1y != 0.0) r2x = __ddiv_rn(__dmul_rn(r1x, __dsub_rn(p[1], p2[1])), r1y);
          const double yy = __dmul_rn(y_diff1, y_diff2);
          const double xs = __dadd_rn(r2x, p2[0]);
          if (yy < 0)
            {
              if (xs > p[0])
                ++cross_number;
              else if (xs == p[0])
                return true;
            }
          else if (yy == 0)
            {
              if (y_diff1 == 0 && y_diff2 == 0)
                {
                  if (__dmul_rn(x_diff1, x_diff2) < 0) return true;
                  continue;
                }
              else if (xs > p[0])
                {
                  if (p[1] != S.box[2] && p[1] != S.box[3]) ++half_cross_number;
                }
              else if ((p[0] == p1[0] && p[1] == p1[1]) || (p[0] == p2[0] && p[1] == p2[1]))
                return true;
            }
        }
      cross_number += half_cross_number / 2;
      return cross_number % 2 == 1;
    }

    template <int DIM>
    __device__ __forceinline__ void interpolate(const SolidView &S, int cell, const double *xi, const double *field, double *out)
    {
      constexpr int NV = 1 << DIM;
      double N[NV], dN[NV * DIM];
      q1_shape<DIM>(xi, N, dN);
#pragma unroll
      for (int c = 0; c < DIM; ++c) out[c] = 0.0;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        {
          const int node = S.cells[(int64_t)cell * NV + v];
#pragma unroll
          for (int c = 0; c < DIM; ++c) out[c] = fma(N[v], field[(int64_t)node * DIM + c], out[c]);
        }
    }

    // ---- kernels ---------------------------------------------------------------------------------
    __global__ void deform_kernel(int64_t n, const double *__restrict__ X, const double *__restrict__ u, double *__restrict__ x)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i < n) x[i] = X[i] + u[i];
    }

    // update_solid_box: min / max of every coordinate over the deformed vertices (one block; the solid is small)
    __global__ void __launch_bounds__(256) box_kernel(int n_nodes, int dim, const double *__restrict__ x, double *__restrict__ box)
    {
      __shared__ double smin[3][256], smax[3][256];
      for (int d = 0; d < dim; ++d)
        {
          double lo = 1e300, hi = -1e300;
          for (int i = threadIdx.x; i < n_nodes; i += blockDim.x)
            {
              const double v = x[(int64_t)i * dim + d];
              lo = fmin(lo, v);
              hi = fmax(hi, v);
            }
          smin[d][threadIdx.x] = lo;
          smax[d][threadIdx.x] = hi;
        }
      __syncthreads();
      for (int s = 128; s > 0; s >>= 1)
        {
          if (threadIdx.x < s)
            for (int d = 0; d < dim; ++d)
              {
                smin[d][threadIdx.x] = fmin(smin[d][threadIdx.x], smin[d][threadIdx.x + s]);
                smax[d][threadIdx.x] = fmax(smax[d][threadIdx.x], smax[d][threadIdx.x + s]);
              }
          __syncthreads();
        }
      if (threadIdx.x == 0)
        for (int d = 0; d < dim; ++d)
          {
            box[2 * d] = smin[d][0];
            box[2 * d + 1] = smax[d][0];
          }
    }

    template <int DIM>
    __device__ __forceinline__ void cell_bin_range(const SolidView &S, int cell, int *b0, int *b1)
    {
      constexpr int NV = 1 << DIM;
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        {
          double lo = 1e300, hi = -1e300;
          for (int v = 0; v < NV; ++v)
            {
              const double c = S.x[(int64_t)S.cells[(int64_t)cell * NV + v] * DIM + d];
              lo = fmin(lo, c);
              hi = fmax(hi, c);
            }
          const double ext = S.box[2 * d + 1] - S.box[2 * d];
          // pad by the query slack so that a point within tolerance of a cell is found in its bin
          const double pad = 1e-9 * fmax(ext, 1.0);
          int k0 = ext > 0 ? (int)floor((lo - pad - S.box[2 * d]) / ext * S.nbin[d]) : 0;
          int k1 = ext > 0 ? (int)floor((hi + pad - S.box[2 * d]) / ext * S.nbin[d]) : 0;
          b0[d] = min(max(k0, 0), S.nbin[d] - 1);
          b1[d] = min(max(k1, 0), S.nbin[d] - 1);
        }
    }

    // pass 0: count cells per bin; pass 1: fill (cursor starts at bin_start)
    template <int DIM>
    __global__ void bin_kernel(SolidView S, int pass, int *__restrict__ count_or_cursor, int *__restrict__ items)
    {
      const int cell = blockIdx.x * blockDim.x + threadIdx.x;
      if (cell >= S.n_cells) return;
      int b0[3] = {0, 0, 0}, b1[3] = {0, 0, 0};
      cell_bin_range<DIM>(S, cell, b0, b1);
      for (int k = b0[2]; k <= b1[2]; ++k)
        for (int j = b0[1]; j <= b1[1]; ++j)
          for (int i = b0[0]; i <= b1[0]; ++i)
            {
              const int b = i + S.nbin[0] * (j + S.nbin[1] * k);
              const int pos = atomicAdd(&count_or_cursor[b], 1);
              if (pass == 1) items[pos] = cell;
            }
    }

    // update_indicator (mpi_fsi.cpp:296-317): one thread per local fluid cell
    template <int DIM>
    __global__ void indicator_kernel(SolidView S, int n_cells, const double *__restrict__ cell_x, int *__restrict__ indicator)
    {
      constexpr int NV = 1 << DIM;
      const int c = blockIdx.x * blockDim.x + threadIdx.x;
      if (c >= n_cells) return;
      int inside = 0;
      for (int v = 0; v < NV; ++v)
        {
          if (!point_in_solid<DIM>(S, cell_x + ((int64_t)c * NV + v) * DIM)) break;
          ++inside;
        }
      indicator[c] = inside == NV ? 1 : 0;
    }

    struct FluidBcArgs
    {
      int n_owned_nodes, nu, n_u_dofs_offset;
      const int *n2c_ptr, *n2c_cell, *n2c_loc, *cell_un, *indicator;
      const unsigned char *node_interior, *cell_owned;
      const double *un_coords, *cell_x, *sp_tables, *present, *solid_vel, *solid_acc;
      double inv_dt;
      int use_dirichlet;
      double *fsi_acc;
      unsigned char *inner_con;
      double *inner_inhom;
      int *error_flag;
    };

    // find_fluid_bc (mpi_fsi.cpp:478-638): one thread per owned fluid velocity node. The reference's
    // "first touching cell wins" (dof_touched) becomes: the lowest-numbered adjacent indicator-1 cell.
    __global__ void merge_constraints_kernel(int64_t n, const unsigned char *__restrict__ inner_con, const double *__restrict__ inner_inhom,
                                             const unsigned char *__restrict__ hanging, unsigned char *__restrict__ con, double *__restrict__ val)
    {
      const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (g >= n) return;
      if (inner_con[g] && !con[g] && !(hanging && hanging[g]))
        {
          con[g] = 1;
          val[g] = inner_inhom[g];
        }
    }

    template <int DIM>
    __global__ void fluid_bc_kernel(SolidView S, FluidBcArgs A)
    {
      constexpr int NV = 1 << DIM;
      const int node = blockIdx.x * blockDim.x + threadIdx.x;
      if (node >= A.n_owned_nodes) return;
      const double *p = A.un_coords + (int64_t)node * DIM;
      if (A.use_dirichlet)
        {
          if (A.node_interior[node]) return; // skip the in-cell support point (:590-602)
          if (!point_in_solid<DIM>(S, p)) return;
          double xi[DIM], vs[DIM];
          const int sc = locate<DIM>(S, p, xi);
          if (sc < 0)
            {
              atomicExch(A.error_flag, 1); // "Cannot find point in solid"
              return;
            }
          interpolate<DIM>(S, sc, xi, A.solid_vel, vs);
#pragma unroll
          for (int c = 0; c < DIM; ++c)
            {
              const int64_t g = (int64_t)DIM * node + c;
              A.inner_con[g] = 1;
              A.inner_inhom[g] = vs[c] - A.present[g];
            }
          return;
        }
      // first adjacent cell with indicator == 1
      int cell = -1, loc = 0;
      for (int k = A.n2c_ptr[node]; k < A.n2c_ptr[node + 1]; ++k)
        if (A.indicator[A.n2c_cell[k]] == 1 && A.cell_owned[A.n2c_cell[k]])
          {
            cell = A.n2c_cell[k];
            loc = A.n2c_loc[k];
            break;
          }
      if (cell < 0) return;
      if (!point_in_solid<DIM>(S, p)) return;
      double xi[DIM], vs[DIM], as[DIM];
      const int sc = locate<DIM>(S, p, xi);
      if (sc < 0)
        {
          atomicExch(A.error_flag, 1);
          return;
        }
      interpolate<DIM>(S, sc, xi, A.solid_vel, vs);
      interpolate<DIM>(S, sc, xi, A.solid_acc, as);
      // geometry of the fluid cell at this support point
      const int nu = A.nu;
      const double *dNu = A.sp_tables + (int64_t)loc * nu * DIM;                       // [nu][DIM] at support point loc
      const double *dNg = A.sp_tables + (int64_t)nu * nu * DIM + (int64_t)loc * NV * DIM; // [NV][DIM]
      const double *X = A.cell_x + (int64_t)cell * NV * DIM;
      double J[DIM * DIM], Ji[DIM * DIM];
#pragma unroll
      for (int i = 0; i < DIM * DIM; ++i) J[i] = 0.0;
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int i = 0; i < DIM; ++i)
#pragma unroll
          for (int j = 0; j < DIM; ++j) J[i * DIM + j] = fma(X[v * DIM + i], dNg[v * DIM + j], J[i * DIM + j]);
      if (DIM == 2)
        {
          const double d = 1.0 / (J[0] * J[3] - J[1] * J[2]);
          Ji[0] = J[3] * d; Ji[1] = -J[1] * d; Ji[2] = -J[2] * d; Ji[3] = J[0] * d;
        }
      else
        {
          const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
          const double d = 1.0 / (J[0] * c00 + J[1] * c01 + J[2] * c02);
          Ji[0] = c00 * d; Ji[1] = (J[2] * J[7] - J[1] * J[8]) * d; Ji[2] = (J[1] * J[5] - J[2] * J[4]) * d;
          Ji[3] = c01 * d; Ji[4] = (J[0] * J[8] - J[2] * J[6]) * d; Ji[5] = (J[2] * J[3] - J[0] * J[5]) * d;
          Ji[6] = c02 * d; Ji[7] = (J[1] * J[6] - J[0] * J[7]) * d; Ji[8] = (J[0] * J[4] - J[1] * J[3]) * d;
        }
      // grad_v[i][k] = sum_b U_b[i] dN_b/dx_k at the support point; v = nodal value
      double gv[DIM * DIM], v[DIM];
#pragma unroll
      for (int i = 0; i < DIM * DIM; ++i) gv[i] = 0.0;
      for (int b = 0; b < nu; ++b)
        {
          const int nb = A.cell_un[(int64_t)cell * nu + b];
          double g[DIM];
#pragma unroll
          for (int k = 0; k < DIM; ++k)
            {
              double s = 0.0;
#pragma unroll
              for (int j = 0; j < DIM; ++j) s = fma(dNu[b * DIM + j], Ji[j * DIM + k], s);
              g[k] = s;
            }
#pragma unroll
          for (int i = 0; i < DIM; ++i)
            {
              const double ub = A.present[(int64_t)DIM * nb + i];
#pragma unroll
              for (int k = 0; k < DIM; ++k) gv[i * DIM + k] = fma(ub, g[k], gv[i * DIM + k]);
            }
        }
#pragma unroll
      for (int i = 0; i < DIM; ++i) v[i] = A.present[(int64_t)DIM * node + i];
#pragma unroll
      for (int c = 0; c < DIM; ++c)
        {
          double conv = 0.0;
#pragma unroll
          for (int k = 0; k < DIM; ++k) conv = fma(gv[c * DIM + k], v[k], conv);
          // (v_s - v_f)/dt + (grad v_f) v_f - a_s   (:559-565)
          A.fsi_acc[(int64_t)DIM * node + c] = (vs[c] - v[c]) * A.inv_dt + conv - as[c];
        }
    }

    // first part of find_fluid_bc (mpi_fsi.cpp:411-476): fsi_stress[k](dof) = sigma_fluid_k(dof) - sigma_solid_k(x_dof)
    // on the scalar FE_Q(pu) support points of owned indicator-1 cells that lie in the solid, k over (i, j <= i)
    template <int DIM>
    __global__ void fluid_stress_bc_kernel(SolidView S, int n_owned_nodes, const int *__restrict__ n2c_ptr, const int *__restrict__ n2c_cell,
                                           const int *__restrict__ indicator, const unsigned char *__restrict__ cell_owned,
                                           const double *__restrict__ un_coords, int n_unodes, const double *__restrict__ fluid_stress,
                                           int n_snodes, const double *__restrict__ solid_stress, double *__restrict__ fsi_stress,
                                           int *__restrict__ error_flag)
    {
      constexpr int NV = 1 << DIM;
      const int node = blockIdx.x * blockDim.x + threadIdx.x;
      if (node >= n_owned_nodes) return;
      bool hit = false;
      for (int k = n2c_ptr[node]; k < n2c_ptr[node + 1] && !hit; ++k)
        hit = indicator[n2c_cell[k]] != 0 && cell_owned[n2c_cell[k]];
      if (!hit) return;
      const double *p = un_coords + (int64_t)node * DIM;
      if (!point_in_solid<DIM>(S, p)) return;
      double xi[DIM];
      const int sc = locate<DIM>(S, p, xi);
      if (sc < 0)
        {
          // GridInterpolator::point_value returns 0 when the point is not found (utilities.cpp:228-233)
          int si = 0;
          for (int i = 0; i < DIM; ++i)
            for (int j = 0; j <= i; ++j, ++si) fsi_stress[(int64_t)si * n_unodes + node] = fluid_stress[(int64_t)(i * DIM + j) * n_unodes + node];
          (void)error_flag;
          return;
        }
      double N[NV], dN[NV * DIM];
      q1_shape<DIM>(xi, N, dN);
      int si = 0;
      for (int i = 0; i < DIM; ++i)
        for (int j = 0; j <= i; ++j, ++si)
          {
            double ss = 0.0;
#pragma unroll
            for (int v = 0; v < NV; ++v) ss = fma(N[v], solid_stress[(int64_t)(i * DIM + j) * n_snodes + S.cells[(int64_t)sc * NV + v]], ss);
            fsi_stress[(int64_t)si * n_unodes + node] = fluid_stress[(int64_t)(i * DIM + j) * n_unodes + node] - ss;
          }
    }

    // ---- locating a point in the (static) fluid mesh: GridInterpolator on the fluid DoFHandler -------------------
    struct FluidView
    {
      const double *cell_x; // [n_cells][2^dim][dim]
      const unsigned char *cell_owned;
      const double *box;    // [2*dim]
      int nbin[3];
      const int *bin_start, *bin_items;
    };

    template <int DIM>
    __device__ bool point_in_fluid_cell(const double *V, const double *p, double *xi_out)
    {
      constexpr int NV = 1 << DIM;
      double lo[DIM], hi[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        {
          lo[d] = 1e300;
          hi[d] = -1e300;
        }
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int d = 0; d < DIM; ++d)
          {
            lo[d] = fmin(lo[d], V[v * DIM + d]);
            hi[d] = fmax(hi[d], V[v * DIM + d]);
          }
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        if (p[d] < lo[d] - kBoxSlack || p[d] > hi[d] + kBoxSlack) return false;
      double xi[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) xi[d] = 0.5;
      bool converged = false;
      for (int it = 0; it < 30 && !converged; ++it)
        {
          double N[NV], dN[NV * DIM], r[DIM], J[DIM * DIM], dx[DIM];
          q1_shape<DIM>(xi, N, dN);
#pragma unroll
          for (int i = 0; i < DIM; ++i)
            {
              double sacc = -p[i];
#pragma unroll
              for (int v = 0; v < NV; ++v) sacc += N[v] * V[v * DIM + i];
              r[i] = sacc;
#pragma unroll
              for (int j = 0; j < DIM; ++j)
                {
                  double t = 0.0;
#pragma unroll
                  for (int v = 0; v < NV; ++v) t += V[v * DIM + i] * dN[v * DIM + j];
                  J[i * DIM + j] = t;
                }
            }
          if (!solve_small<DIM>(J, r, dx)) return false;
          double n2 = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              xi[d] -= dx[d];
              n2 += dx[d] * dx[d];
            }
          converged = sqrt(n2) < 1e-13;
        }
      if (!converged) return false;
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        if (xi[d] < -kUnitTol || xi[d] > 1.0 + kUnitTol) return false;
#pragma unroll
      for (int d = 0; d < DIM; ++d) xi_out[d] = fmin(fmax(xi[d], 0.0), 1.0);
      return true;
    }

    // lowest-numbered local fluid cell containing p, -1 if none
    template <int DIM>
    __device__ int locate_fluid(const FluidView &F, const double *p, double *xi)
    {
      constexpr int NV = 1 << DIM;
      int b = 0, stride = 1;
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        {
          if (p[d] < F.box[2 * d] - kBoxSlack || p[d] > F.box[2 * d + 1] + kBoxSlack) return -1;
          const double ext = F.box[2 * d + 1] - F.box[2 * d];
          int k = ext > 0 ? (int)floor((p[d] - F.box[2 * d]) / ext * F.nbin[d]) : 0;
          k = min(max(k, 0), F.nbin[d] - 1);
          b += k * stride;
          stride *= F.nbin[d];
        }
      int best = -1;
      for (int k = F.bin_start[b]; k < F.bin_start[b + 1]; ++k)
        {
          const int c = F.bin_items[k];
          if (best >= 0 && c > best) continue;
          double t[DIM];
          if (point_in_fluid_cell<DIM>(F.cell_x + (int64_t)c * NV * DIM, p, t))
            {
              best = c;
#pragma unroll
              for (int d = 0; d < DIM; ++d) xi[d] = t[d];
            }
        }
      return best;
    }

    // FE_Q(p) shape values at xi, p = 1 or 2, nodes lexicographic (x fastest)
    template <int DIM>
    __device__ void lagrange_shape(int p, const double *xi, double *N)
    {
      double L[DIM][3];
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        {
          const double x = xi[d];
          if (p == 1)
            {
              L[d][0] = 1.0 - x;
              L[d][1] = x;
              L[d][2] = 0.0;
            }
          else
            {
              L[d][0] = 2.0 * (x - 0.5) * (x - 1.0);
              L[d][1] = -4.0 * x * (x - 1.0);
              L[d][2] = 2.0 * x * (x - 0.5);
            }
        }
      const int n1 = p + 1;
      int n = 1;
      for (int d = 0; d < DIM; ++d) n *= n1;
      for (int a = 0; a < n; ++a)
        {
          int r = a;
          double v = 1.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              v *= L[d][r % n1];
              r /= n1;
            }
          N[a] = v;
        }
    }

    struct SolidBcArgs
    {
      int n_bvert;
      const int *bvert;
      const double *x; // deformed solid vertices
      int pu, pp, nu, np;
      const int *cell_un, *cell_pn;
      int64_t n_u;
      int n_unodes;
      const double *present, *fluid_stress; // fluid_stress may be null
      int64_t n_sdofs;
      double *rows, *fluid_velocity, *fluid_pressure;
    };

    // find_solid_bc (mpi_fsi.cpp:704-806): one thread per vertex of the solid's non-fixed boundary faces
    template <int DIM>
    __global__ void solid_bc_kernel(FluidView F, SolidBcArgs A)
    {
      const int k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= A.n_bvert) return;
      const int node = A.bvert[k];
      const double *p = A.x + (int64_t)node * DIM;
      double xi[DIM];
      const int cell = locate_fluid<DIM>(F, p, xi);
      // not found, or found in a cell another rank owns: contributes zero (utilities.cpp:228-233) and the
      // all-reduce over ranks fills it in
      if (cell < 0 || !F.cell_owned[cell]) return;
      double Nu[27], Np[27];
      lagrange_shape<DIM>(A.pu, xi, Nu);
      lagrange_shape<DIM>(A.pp, xi, Np);
      double v[DIM], pr = 0.0, visc[DIM * DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) v[d] = 0.0;
#pragma unroll
      for (int i = 0; i < DIM * DIM; ++i) visc[i] = 0.0;
      for (int b = 0; b < A.nu; ++b)
        {
          const int un = A.cell_un[(int64_t)cell * A.nu + b];
#pragma unroll
          for (int d = 0; d < DIM; ++d) v[d] = fma(Nu[b], A.present[(int64_t)DIM * un + d], v[d]);
          if (A.fluid_stress)
#pragma unroll
            for (int i = 0; i < DIM; ++i)
#pragma unroll
              for (int j = i; j < DIM; ++j) visc[i * DIM + j] = fma(Nu[b], A.fluid_stress[(int64_t)(i * DIM + j) * A.n_unodes + un], visc[i * DIM + j]);
        }
      for (int j = 0; j < A.np; ++j) pr = fma(Np[j], A.present[A.n_u + A.cell_pn[(int64_t)cell * A.np + j]], pr);
      // sigma = -p I + viscous (symmetric)
#pragma unroll
      for (int d1 = 0; d1 < DIM; ++d1)
        {
#pragma unroll
          for (int d2 = 0; d2 < DIM; ++d2)
            {
              const double sv = d1 <= d2 ? visc[d1 * DIM + d2] : visc[d2 * DIM + d1];
              A.rows[(int64_t)d1 * A.n_sdofs + (int64_t)DIM * node + d2] = sv - (d1 == d2 ? pr : 0.0);
            }
          A.fluid_velocity[(int64_t)DIM * node + d1] = v[d1];
        }
      A.fluid_pressure[node] = pr;
    }

    template <int DIM>
    __global__ void query_kernel(SolidView S, int n, const double *__restrict__ pts, int *__restrict__ inside, const double *field,
                                 double *__restrict__ values, int *__restrict__ found)
    {
      const int i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n) return;
      const double *p = pts + (int64_t)i * DIM;
      if (inside) inside[i] = point_in_solid<DIM>(S, p) ? 1 : 0;
      if (field)
        {
          double xi[DIM], out[DIM];
          const int c = locate<DIM>(S, p, xi);
          found[i] = c;
          if (c >= 0)
            {
              interpolate<DIM>(S, c, xi, field, out);
              for (int d = 0; d < DIM; ++d) values[(int64_t)i * DIM + d] = out[d];
            }
          else
            for (int d = 0; d < DIM; ++d) values[(int64_t)i * DIM + d] = 0.0; // point_value returns 0 when not found
        }
    }

    struct ScopedTimer
    {
      Context &ctx;
      double &acc;
      std::chrono::steady_clock::time_point t0;
      ScopedTimer(Context &c, double &a) : ctx(c), acc(a)
      {
        cudaStreamSynchronize(ctx.stream);
        t0 = std::chrono::steady_clock::now();
      }
      ~ScopedTimer()
      {
        cudaStreamSynchronize(ctx.stream);
        acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      }
    };
  } // namespace

  // ===========================================================================
  FsiCoupling::FsiCoupling(Context &ctx_, InsIM &fluid_, SolidSolver &solid_, const Parameters::AllParameters &params,
                           bool use_dirichlet_bc_)
    : ctx(ctx_), fluid(fluid_), solid(solid_), parameters(params), use_dirichlet_bc(use_dirichlet_bc_),
      time(params.end_time, params.time_step, params.output_interval, params.refinement_interval, params.save_interval)
  {
    if (!fluid.dofs_ready || !solid.dofs_ready) throw std::runtime_error("FSI: set up the fluid and solid solvers first");
    dim = fluid.fs.dim;
    if (solid.ss.dim != dim) throw std::runtime_error("FSI: fluid and solid dimensions differ");
    const FluidSpace &fs = fluid.fs;
    const SolidSpace &ss = solid.ss;
    cudaStream_t s = ctx.stream;
    d_x.alloc((size_t)ss.nt.n_nodes * dim);
    d_box.alloc(2 * dim);
    solid_box.assign(2 * dim, 0.0);
    // 2-D boundary segments of the solid: both vertices of every boundary face (collect_solid_boundaries, :77-92)
    if (dim == 2)
      {
        static const int fv[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};
        std::vector<int> seg;
        const Triangulation &st = solid.triangulation;
        for (int f = 0; f < st.n_boundary_faces(); ++f)
          {
            const int cell = st.boundary_faces[3 * f], face = st.boundary_faces[3 * f + 1];
            seg.push_back(ss.nt.cell_nodes[(size_t)cell * ss.npc + fv[face][0]]);
            seg.push_back(ss.nt.cell_nodes[(size_t)cell * ss.npc + fv[face][1]]);
          }
        n_bseg = (int)seg.size() / 2;
        d_bseg.upload(seg, s);
      }
    // vertices of the solid's boundary faces that are not fully fixed (mpi_fsi.cpp:690-703)
    {
      const Triangulation &st = solid.triangulation;
      const unsigned fixed = (1u << dim) - 1;
      std::vector<int> bv;
      for (int f = 0; f < st.n_boundary_faces(); ++f)
        {
          auto bc = parameters.solid_dirichlet_bcs.find((unsigned)st.boundary_faces[3 * f + 2]);
          if (bc != parameters.solid_dirichlet_bcs.end() && bc->second == fixed) continue;
          const int cell = st.boundary_faces[3 * f], face = st.boundary_faces[3 * f + 1];
          for (int a : face_local_nodes(dim, 1, face)) bv.push_back(ss.nt.cell_nodes[(size_t)cell * ss.npc + a]);
        }
      std::sort(bv.begin(), bv.end());
      bv.erase(std::unique(bv.begin(), bv.end()), bv.end());
      n_solid_bvert = (int)bv.size();
      if (n_solid_bvert) d_solid_bvert.upload(bv, s);
    }
    build_fluid_side();
    IFEM_CUDA(cudaStreamSynchronize(s));
  }

  void FsiCoupling::build_fluid_side()
  {
    const FluidSpace &fs = fluid.fs;
    cudaStream_t s = ctx.stream;
    // fluid velocity node -> (cell, local index), cells ascending
    {
      const int nn = fs.un.n_nodes, nu = fs.nu;
      std::vector<int> ptr(nn + 1, 0);
      for (size_t k = 0; k < fs.un.cell_nodes.size(); ++k) ptr[fs.un.cell_nodes[k] + 1]++;
      for (int i = 0; i < nn; ++i) ptr[i + 1] += ptr[i];
      std::vector<int> cell(ptr[nn]), loc(ptr[nn]), pos(ptr.begin(), ptr.end() - 1);
      std::vector<unsigned char> interior(nn, 0);
      FEQ fe(dim, fs.pu);
      for (int c = 0; c < fs.n_cells; ++c)
        for (int a = 0; a < nu; ++a)
          {
            const int n = fs.un.cell_nodes[(size_t)c * nu + a];
            cell[pos[n]] = c;
            loc[pos[n]++] = a;
            bool in = true;
            for (int d = 0; d < dim; ++d) in = in && fe.lattice[a][d] > 0 && fe.lattice[a][d] < fs.pu;
            if (in) interior[n] = 1;
          }
      // is_locally_owned() of every local fluid cell: the slab rank of the cell is this rank
      std::vector<unsigned char> owned(fs.n_cells, 1);
      if (fs.n_ranks > 1)
        {
          const std::vector<int> cr = slab_cell_ranks(fluid.triangulation, fs.n_ranks);
          for (int c = 0; c < fs.n_cells; ++c) owned[c] = cr[fs.local_cells[c]] == fs.rank;
        }
      d_cell_owned.upload(owned, s);
      d_n2c_ptr.upload(ptr, s);
      d_n2c_cell.upload(cell, s);
      d_n2c_loc.upload(loc, s);
      d_node_interior.upload(interior, s);
      d_un_coords.upload(fs.un.coords, s);
      // shape-gradient tables at the unit support points of the velocity element
      std::vector<double> t((size_t)nu * nu * dim + (size_t)nu * fs.nv * dim), N(nu), g(fs.nv);
      for (int a = 0; a < nu; ++a)
        {
          double xi[3];
          for (int d = 0; d < dim; ++d) xi[d] = double(fe.lattice[a][d]) / fs.pu;
          fs.fe_u.eval(xi, N.data(), &t[(size_t)a * nu * dim]);
          fs.fe_geo.eval(xi, g.data(), &t[(size_t)nu * nu * dim + (size_t)a * fs.nv * dim]);
        }
      d_sp_tables.upload(t, s);
    }
    d_inner_con.alloc(fs.n_dofs);
    d_inner_inhom.alloc(fs.n_dofs);
    build_fluid_bins();
  }

  void FsiCoupling::refine_mesh(unsigned int min_grid_level, unsigned int max_grid_level)
  {
    ScopedTimer t(ctx, timer_ms["Refine mesh"]);
    Triangulation &ft = fluid.triangulation;
    const Triangulation &st = solid.triangulation;
    const SolidSpace &ss = solid.ss;
    // move_solid_mesh(true): deformed vertex positions (:1029)
    refresh_deformed();
    const std::vector<double> x = d_x.to_host(ctx.stream);
    // one point per solid boundary cell: the centre of its first boundary face (:1030-1050)
    std::vector<double> pts;
    {
      std::vector<int> first_face((size_t)st.n_cells(), -1);
      for (int f = 0; f < st.n_boundary_faces(); ++f)
        {
          int &ff = first_face[st.boundary_faces[3 * f]];
          const int face = st.boundary_faces[3 * f + 1];
          if (ff < 0 || face < ff) ff = face;
        }
      for (int c = 0; c < st.n_cells(); ++c)
        {
          if (first_face[c] < 0) continue;
          const std::vector<int> fn = face_local_nodes(dim, 1, first_face[c]);
          double p[3] = {0, 0, 0};
          for (int a : fn)
            for (int d = 0; d < dim; ++d) p[d] += x[(size_t)ss.nt.cell_nodes[(size_t)c * ss.npc + a] * dim + d] / fn.size();
          pts.insert(pts.end(), p, p + dim);
        }
    }
    const int n_pts = (int)(pts.size() / dim), nc = ft.n_cells(), vpc = ft.verts_per_cell();
    std::vector<unsigned char> refine((size_t)nc, 0), coarsen((size_t)nc, 0);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < nc; ++c)
      {
        double ctr[3] = {0, 0, 0}, diam2 = 0.0;
        for (int v = 0; v < vpc; ++v)
          for (int d = 0; d < dim; ++d) ctr[d] += ft.vertices[(size_t)ft.cells[(size_t)c * vpc + v] * dim + d] / vpc;
        for (int v = 0; v < vpc; ++v)
          for (int w = v + 1; w < vpc; ++w)
            {
              double d2 = 0.0;
              for (int d = 0; d < dim; ++d)
                {
                  const double e = ft.vertices[(size_t)ft.cells[(size_t)c * vpc + v] * dim + d] - ft.vertices[(size_t)ft.cells[(size_t)c * vpc + w] * dim + d];
                  d2 += e * e;
                }
              diam2 = std::max(diam2, d2);
            }
        double best = std::numeric_limits<double>::max();
        for (int k = 0; k < n_pts; ++k)
          {
            double d2 = 0.0;
            for (int d = 0; d < dim; ++d) d2 += (ctr[d] - pts[(size_t)k * dim + d]) * (ctr[d] - pts[(size_t)k * dim + d]);
            best = std::min(best, d2);
          }
        if (std::sqrt(best) < std::sqrt(diam2)) refine[c] = 1;
        else coarsen[c] = 1;
      }
    // level limits (:1064-1080)
    const std::vector<int> &level = ft.cell_level;
    auto lvl = [&](int c) { return level.empty() ? 0 : level[c]; };
    if ((unsigned)ft.n_levels() > max_grid_level)
      for (int c = 0; c < nc; ++c)
        if ((unsigned)lvl(c) >= max_grid_level) refine[c] = 0;
    for (int c = 0; c < nc; ++c)
      if ((unsigned)lvl(c) == min_grid_level) coarsen[c] = 0;
    const std::vector<double> old_vertices = ft.vertices;
    Triangulation::TransferPlan plan;
    ft.execute_coarsening_and_refinement(refine, coarsen, &plan);
    fluid.after_mesh_change(plan, old_vertices);
    build_fluid_side();
    deformed_valid = false;
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void FsiCoupling::build_fluid_bins()
  {
    const FluidSpace &fs = fluid.fs;
    const Triangulation &ft = fluid.triangulation;
    const int nv = fs.nv;
    // bounding box and per-cell boxes of the local fluid cells
    for (int d = 0; d < 3; ++d)
      {
        fluid_box[2 * d] = 1e300;
        fluid_box[2 * d + 1] = -1e300;
      }
    std::vector<double> lo((size_t)fs.n_cells * dim), hi((size_t)fs.n_cells * dim);
    for (int c = 0; c < fs.n_cells; ++c)
      for (int d = 0; d < dim; ++d)
        {
          double a = 1e300, b = -1e300;
          for (int v = 0; v < nv; ++v)
            {
              const double x = ft.vertices[(size_t)ft.cells[(size_t)fs.local_cells[c] * nv + v] * dim + d];
              a = std::min(a, x);
              b = std::max(b, x);
            }
          lo[(size_t)c * dim + d] = a;
          hi[(size_t)c * dim + d] = b;
          fluid_box[2 * d] = std::min(fluid_box[2 * d], a);
          fluid_box[2 * d + 1] = std::max(fluid_box[2 * d + 1], b);
        }
    double vol = 1.0, ext[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d)
      {
        ext[d] = std::max(fluid_box[2 * d + 1] - fluid_box[2 * d], 1e-300);
        vol *= ext[d];
      }
    const double h = std::pow(vol / std::max(1, fs.n_cells), 1.0 / dim);
    int64_t total = 1;
    for (int d = 0; d < 3; ++d)
      {
        fbin[d] = d < dim ? std::max(1, std::min(512, (int)std::floor(ext[d] / h))) : 1;
        total *= fbin[d];
      }
    auto range = [&](int c, int d, int &k0, int &k1) {
      const double pad = 1e-9 * std::max(ext[d], 1.0);
      k0 = (int)std::floor((lo[(size_t)c * dim + d] - pad - fluid_box[2 * d]) / ext[d] * fbin[d]);
      k1 = (int)std::floor((hi[(size_t)c * dim + d] + pad - fluid_box[2 * d]) / ext[d] * fbin[d]);
      k0 = std::min(std::max(k0, 0), fbin[d] - 1);
      k1 = std::min(std::max(k1, 0), fbin[d] - 1);
    };
    std::vector<int> start((size_t)total + 1, 0);
    for (int pass = 0; pass < 2; ++pass)
      {
        std::vector<int> cursor;
        std::vector<int> items;
        if (pass == 1)
          {
            for (int64_t b = 0; b < total; ++b) start[b + 1] += start[b];
            cursor.assign(start.begin(), start.end() - 1);
            items.resize(start[total]);
          }
        for (int c = 0; c < fs.n_cells; ++c)
          {
            int k0[3] = {0, 0, 0}, k1[3] = {0, 0, 0};
            for (int d = 0; d < dim; ++d) range(c, d, k0[d], k1[d]);
            for (int k = k0[2]; k <= k1[2]; ++k)
              for (int j = k0[1]; j <= k1[1]; ++j)
                for (int i = k0[0]; i <= k1[0]; ++i)
                  {
                    const int64_t b = i + (int64_t)fbin[0] * (j + (int64_t)fbin[1] * k);
                    if (pass == 0) start[b + 1]++;
                    else items[cursor[b]++] = c;
                  }
          }
        if (pass == 1)
          {
            d_fbin_start.upload(start, ctx.stream);
            d_fbin_items.upload(items, ctx.stream);
          }
      }
    d_fluid_box.upload(std::vector<double>(fluid_box, fluid_box + 6), ctx.stream);
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void FsiCoupling::find_solid_bc()
  {
    ScopedTimer t(ctx, timer_ms["Find solid BC"]);
    refresh_deformed(); // must use the updated solid coordinates (:669-670)
    const FluidSpace &fs = fluid.fs;
    SolidSpace &ss = solid.ss;
    cudaStream_t s = ctx.stream;
    solid.fsi_stress_rows.zero(s);
    solid.fluid_velocity.zero(s);
    solid.fluid_pressure.zero(s);
    if (n_solid_bvert)
      {
        FluidView F{};
        F.cell_x = fs.d_cell_x.p;
        F.cell_owned = d_cell_owned.p;
        F.box = d_fluid_box.p;
        for (int d = 0; d < 3; ++d) F.nbin[d] = fbin[d];
        F.bin_start = d_fbin_start.p;
        F.bin_items = d_fbin_items.p;
        SolidBcArgs A{};
        A.n_bvert = n_solid_bvert;
        A.bvert = d_solid_bvert.p;
        A.x = d_x.p;
        A.pu = fs.pu;
        A.pp = fs.pp;
        A.nu = fs.nu;
        A.np = fs.np;
        A.cell_un = fs.d_cell_un.p;
        A.cell_pn = fs.d_cell_pn.p;
        A.n_u = fs.n_u;
        A.n_unodes = fs.un.n_nodes;
        A.present = fluid.present_solution.p;
        A.fluid_stress = fluid.stress.p; // update_stress() output of the last fluid step
        A.n_sdofs = ss.n_dofs;
        A.rows = solid.fsi_stress_rows.p;
        A.fluid_velocity = solid.fluid_velocity.p;
        A.fluid_pressure = solid.fluid_pressure.p;
        const int blocks = (n_solid_bvert + 127) / 128;
        if (dim == 2) solid_bc_kernel<2><<<blocks, 128, 0, s>>>(F, A);
        else solid_bc_kernel<3><<<blocks, 128, 0, s>>>(F, A);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
    if (fs.n_ranks > 1)
      {
        // Utilities::MPI::sum of the replicated solid vectors (:849-865)
        comm_allreduce_sum(*ctx.comm, solid.fsi_stress_rows.p, (int)solid.fsi_stress_rows.n, s);
        comm_allreduce_sum(*ctx.comm, solid.fluid_velocity.p, (int)solid.fluid_velocity.n, s);
        comm_allreduce_sum(*ctx.comm, solid.fluid_pressure.p, (int)solid.fluid_pressure.n, s);
      }
  }

  // one pass of the while loop of FSI::run (mpi_fsi.cpp:1172-1214)
  void FsiCoupling::run_one_step(bool first_step, bool stop_before_fluid_step)
  {
    find_solid_bc();
    if (restarted) solid.after_restart(); // "solid_solver.assemble_system(true)" in every pass after a restart (:1176-1179)
    {
      ScopedTimer t(ctx, timer_ms["Run solid solver"]);
      if (penetration_criterion)
        apply_contact_model(first_step);
      else
        solid.run_one_step(first_step);
    }
    update_solid_box();
    update_indicator();
    fluid.make_constraints();
    if (!first_step) fluid.fs.d_nonzero_val.zero(ctx.stream); // nonzero_constraints.copy_from(zero_constraints) (:1193-1198): homogeneous increments from now on
    SCnsIM *supg = dynamic_cast<SCnsIM *>(&fluid);
    SpalartAllmaras *turbulence_model = supg ? supg->turbulence_model.get() : nullptr;
    if (turbulence_model) turbulence_model->update_boundary_condition(first_step); // :1199-1203
    find_fluid_bc();
    if (stop_before_fluid_step) return; // tests: the state the fluid solver is about to see
    {
      ScopedTimer t(ctx, timer_ms["Run fluid solver"]);
      if (turbulence_model) turbulence_model->run_one_step(true); // :1207-1210
      fluid.run_one_step(true);
    }
    time.increment();
    if (time.time_to_save() && !solid.output_directory.empty() && !fluid.output_directory.empty()) // :1219-1223
      {
        solid.save_checkpoint((int)time.get_timestep());
        fluid.save_checkpoint((int)time.get_timestep());
      }
  }

  void FsiCoupling::set_penetration_criterion(std::function<double(const double *)> criterion, const double *direction)
  {
    penetration_criterion = std::move(criterion);
    for (int d = 0; d < dim; ++d) penetration_direction[d] = direction[d];
  }

  // FSI::apply_contact_model (mpi_fsi.cpp:869-970). The criterion is a host callback and the solid is small and
  // replicated, so the penetration scan runs on the host over the boundary faces; the solid steps run on the device.
  void FsiCoupling::apply_contact_model(bool first_step)
  {
    if (!penetration_criterion) throw std::runtime_error("No penetration criterion specified!");
    const SolidSpace &ss = solid.ss;
    const Triangulation &st = solid.triangulation;
    cudaStream_t s = ctx.stream;
    const VecSpace n(ss.n_dofs);
    const double force_increment = parameters.contact_force_multiplier;
    // cache the six Newmark vectors
    DevBuf<double> cache[6];
    DevBuf<double> *vecs[6] = {&solid.current_acceleration, &solid.current_velocity, &solid.current_displacement,
                               &solid.previous_acceleration, &solid.previous_velocity, &solid.previous_displacement};
    for (int k = 0; k < 6; ++k)
      {
        cache[k].alloc(ss.n_dofs);
        copy(ctx, n, vecs[k]->p, cache[k].p);
      }
    const ContactScan scan(dim, ss.degree);
    bool still_penetrate = true;
    while (still_penetrate)
      {
        still_penetrate = false;
        solid.run_one_step(first_step);
        ++contact_iterations;
        const std::vector<double> u = solid.current_displacement.to_host(s);
        std::vector<double> rows = solid.fsi_stress_rows.to_host(s);
        still_penetrate = scan.run(st.n_boundary_faces(), st.boundary_faces.data(), ss.nt.cell_nodes.data(), ss.npc, ss.nt.coords.data(),
                                   u.data(), ss.n_dofs, penetration_criterion, penetration_direction, force_increment, rows.data());
        if (still_penetrate)
          {
            solid.fsi_stress_rows.upload(rows, s);
            for (int k = 0; k < 6; ++k) copy(ctx, n, cache[k].p, vecs[k]->p);
            solid.time.decrement();
          }
      }
    deformed_valid = false;
  }

  void FsiCoupling::run()
  {
    // restart (:1127-1151): when both solvers write into output directories and have not stepped yet, their latest
    // checkpoints are restored and the coupling clock catches up
    if (!solid.output_directory.empty() && !fluid.output_directory.empty() && solid.time.get_timestep() == 0 && fluid.time.get_timestep() == 0)
      {
        restarted = solid.load_checkpoint() && fluid.load_checkpoint();
        if (solid.time.current() != fluid.time.current())
          throw std::runtime_error("Solid and fluid restart files have different time steps. Check and remove inconsistent restart files!");
        while (time.get_timestep() < solid.time.get_timestep()) time.increment();
      }
    const unsigned int g0 = parameters.global_refinements.empty() ? 0u : (unsigned)parameters.global_refinements[0];
    const bool adaptive = parameters.refinement_interval < parameters.end_time;
    if (adaptive) // :1164-1168
      {
        refine_mesh(g0, g0 + 3);
        refine_mesh(g0, g0 + 3);
      }
    bool first_step = !restarted;
    while (time.end() - time.current() > 1e-12)
      {
        run_one_step(first_step);
        first_step = false;
        if (adaptive && time.time_to_refine()) refine_mesh(g0, g0 + 3); // :1215-1218
      }
  }

  void FsiCoupling::refresh_deformed()
  {
    const SolidSpace &ss = solid.ss;
    const int64_t n = (int64_t)ss.nt.n_nodes * dim;
    deform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx.stream>>>(n, ss.d_node_x.p, solid.current_displacement.p, d_x.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
    deformed_valid = true;
  }

  std::vector<double> FsiCoupling::update_solid_box()
  {
    ScopedTimer t(ctx, timer_ms["Update solid box"]);
    refresh_deformed();
    box_kernel<<<1, 256, 0, ctx.stream>>>(solid.ss.nt.n_nodes, dim, d_x.p, d_box.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
    d_box.download(solid_box.data(), 2 * dim, ctx.stream);
    build_bins();
    return solid_box;
  }

  void FsiCoupling::build_bins()
  {
    const SolidSpace &ss = solid.ss;
    // about one solid cell per bin, split over the dimensions in proportion to the box extents
    double ext[3] = {1, 1, 1}, vol = 1.0;
    for (int d = 0; d < dim; ++d)
      {
        ext[d] = std::max(solid_box[2 * d + 1] - solid_box[2 * d], 1e-300);
        vol *= ext[d];
      }
    const double h = std::pow(vol / std::max(1, ss.n_cells), 1.0 / dim);
    int64_t total = 1;
    for (int d = 0; d < 3; ++d)
      {
        nbin[d] = d < dim ? std::max(1, std::min(256, (int)std::floor(ext[d] / h))) : 1;
        total *= nbin[d];
      }
    SolidView S{};
    S.n_cells = ss.n_cells;
    S.cells = ss.d_cell_nodes.p;
    S.x = d_x.p;
    S.box = d_box.p;
    for (int d = 0; d < 3; ++d) S.nbin[d] = nbin[d];
    DevBuf<int> count((size_t)total + 1);
    count.zero(ctx.stream);
    const int blocks = (ss.n_cells + 127) / 128;
    if (dim == 2) bin_kernel<2><<<blocks, 128, 0, ctx.stream>>>(S, 0, count.p, nullptr);
    else bin_kernel<3><<<blocks, 128, 0, ctx.stream>>>(S, 0, count.p, nullptr);
    IFEM_KERNEL_CHECK();
    std::vector<int> h_count = count.to_host(ctx.stream), start((size_t)total + 1, 0);
    for (int64_t b = 0; b < total; ++b) start[b + 1] = start[b] + h_count[b];
    d_bin_start.upload(start, ctx.stream);
    d_bin_cursor.upload(start, ctx.stream);
    if ((int64_t)d_bin_items.n < start[total]) d_bin_items.alloc(start[total]);
    if (dim == 2) bin_kernel<2><<<blocks, 128, 0, ctx.stream>>>(S, 1, d_bin_cursor.p, d_bin_items.p);
    else bin_kernel<3><<<blocks, 128, 0, ctx.stream>>>(S, 1, d_bin_cursor.p, d_bin_items.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches += 2;
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  static SolidView make_view(const FsiCoupling &f, const SolidSpace &ss, const double *x, const double *box, const int *nbin,
                             const int *bin_start, const int *bin_items, int n_bseg, const int *bseg)
  {
    SolidView S{};
    S.n_cells = ss.n_cells;
    S.cells = ss.d_cell_nodes.p;
    S.x = x;
    S.box = box;
    for (int d = 0; d < 3; ++d) S.nbin[d] = nbin[d];
    S.bin_start = bin_start;
    S.bin_items = bin_items;
    S.n_bseg = n_bseg;
    S.bseg = bseg;
    (void)f;
    return S;
  }

  void FsiCoupling::update_indicator()
  {
    ScopedTimer t(ctx, timer_ms["Update indicator"]);
    if (!deformed_valid) update_solid_box();
    const SolidView S = make_view(*this, solid.ss, d_x.p, d_box.p, nbin, d_bin_start.p, d_bin_items.p, n_bseg, d_bseg.p);
    FluidSpace &fs = fluid.fs;
    const int blocks = (fs.n_cells + 127) / 128;
    if (dim == 2) indicator_kernel<2><<<blocks, 128, 0, ctx.stream>>>(S, fs.n_cells, fs.d_cell_x.p, fs.d_indicator.p);
    else indicator_kernel<3><<<blocks, 128, 0, ctx.stream>>>(S, fs.n_cells, fs.d_cell_x.p, fs.d_indicator.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }

  void FsiCoupling::find_fluid_bc()
  {
    ScopedTimer t(ctx, timer_ms["Find fluid BC"]);
    if (!deformed_valid) update_solid_box();
    const SolidView S = make_view(*this, solid.ss, d_x.p, d_box.p, nbin, d_bin_start.p, d_bin_items.p, n_bseg, d_bseg.p);
    FluidSpace &fs = fluid.fs;
    cudaStream_t s = ctx.stream;
    fluid.fsi_acceleration.zero(s); // fresh tmp_fsi_acceleration (:347-349)
    d_inner_con.zero(s);
    d_inner_inhom.zero(s);
    DevBuf<int> err(1);
    err.zero(s);
    if (auto *scns = dynamic_cast<SCnsIM *>(&fluid))
      {
        // implementing the stress part for the fsi force (:411-476); the solid's nodal Cauchy stress comes from
        // update_strain_and_stress()
        const int blocks_s = (fs.n_owned_unodes + 127) / 128;
        if (dim == 2)
          fluid_stress_bc_kernel<2><<<blocks_s, 128, 0, s>>>(S, fs.n_owned_unodes, d_n2c_ptr.p, d_n2c_cell.p, fs.d_indicator.p, d_cell_owned.p,
                                                             d_un_coords.p, fs.un.n_nodes, scns->stress.p, solid.ss.nt.n_nodes, solid.stress.p,
                                                             scns->fsi_stress.p, err.p);
        else
          fluid_stress_bc_kernel<3><<<blocks_s, 128, 0, s>>>(S, fs.n_owned_unodes, d_n2c_ptr.p, d_n2c_cell.p, fs.d_indicator.p, d_cell_owned.p,
                                                             d_un_coords.p, fs.un.n_nodes, scns->stress.p, solid.ss.nt.n_nodes, solid.stress.p,
                                                             scns->fsi_stress.p, err.p);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
        if (fs.n_ranks > 1)
          for (int k = 0; k < dim * (dim + 1) / 2; ++k) fs.halo_p.update(ctx, scns->fsi_stress.p + (size_t)k * fs.un.n_nodes);
      }
    FluidBcArgs A{};
    // Dirichlet variant: the test is purely geometric on a replicated solid, so every LOCAL node (owned and ghost) is
    // flagged - the assembly reads the constraint of ghost column nodes too (column elimination, lifting), and a ghost copy
    // left unconstrained would make the assembled system depend on the partition
    A.n_owned_nodes = use_dirichlet_bc ? fs.un.n_nodes : fs.n_owned_unodes;
    A.nu = fs.nu;
    A.n2c_ptr = d_n2c_ptr.p;
    A.n2c_cell = d_n2c_cell.p;
    A.n2c_loc = d_n2c_loc.p;
    A.cell_un = fs.d_cell_un.p;
    A.indicator = fs.d_indicator.p;
    A.node_interior = d_node_interior.p;
    A.cell_owned = d_cell_owned.p;
    A.un_coords = d_un_coords.p;
    A.cell_x = fs.d_cell_x.p;
    A.sp_tables = d_sp_tables.p;
    A.present = fluid.present_solution.p;
    A.solid_vel = solid.current_velocity.p;
    A.solid_acc = solid.current_acceleration.p;
    A.inv_dt = 1.0 / fluid.time.get_delta_t();
    A.use_dirichlet = use_dirichlet_bc ? 1 : 0;
    A.fsi_acc = fluid.fsi_acceleration.p;
    A.inner_con = d_inner_con.p;
    A.inner_inhom = d_inner_inhom.p;
    A.error_flag = err.p;
    const int blocks = (A.n_owned_nodes + 127) / 128;
    if (dim == 2) fluid_bc_kernel<2><<<blocks, 128, 0, s>>>(S, A);
    else fluid_bc_kernel<3><<<blocks, 128, 0, s>>>(S, A);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
    if (err.to_host(s)[0]) throw std::runtime_error("Cannot find point in solid");
    fs.halo_update(ctx, fluid.fsi_acceleration.p); // fsi_acceleration = tmp (ghosted), :639-640
    if (use_dirichlet_bc)
      {
        // nonzero_constraints.merge(inner_nonzero, left_object_wins) and the same for zero_constraints (:641-651), on the device:
        // an existing line - Dirichlet or hanging-node - wins
        const int64_t n = fs.n_dofs;
        merge_constraints_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, d_inner_con.p, d_inner_inhom.p,
                                                                            fs.hanging.active ? fs.hanging.d_is_hanging_dof.p : nullptr,
                                                                            fs.d_con.p, fs.d_nonzero_val.p);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
        fs.flags_merged = true;
        fs.schur_valid = false;
      }
  }

  void FsiCoupling::point_in_solid(int n, const double *pts, int *inside)
  {
    if (!deformed_valid) update_solid_box();
    const SolidView S = make_view(*this, solid.ss, d_x.p, d_box.p, nbin, d_bin_start.p, d_bin_items.p, n_bseg, d_bseg.p);
    DevBuf<double> dp((size_t)n * dim);
    DevBuf<int> di(n);
    dp.upload(pts, (size_t)n * dim, ctx.stream);
    if (dim == 2) query_kernel<2><<<(n + 127) / 128, 128, 0, ctx.stream>>>(S, n, dp.p, di.p, nullptr, nullptr, nullptr);
    else query_kernel<3><<<(n + 127) / 128, 128, 0, ctx.stream>>>(S, n, dp.p, di.p, nullptr, nullptr, nullptr);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
    di.download(inside, n, ctx.stream);
  }

  void FsiCoupling::interpolate(int which, int n, const double *pts, double *values, int *found)
  {
    if (!deformed_valid) update_solid_box();
    const SolidView S = make_view(*this, solid.ss, d_x.p, d_box.p, nbin, d_bin_start.p, d_bin_items.p, n_bseg, d_bseg.p);
    const double *field = which == 0 ? solid.current_velocity.p : which == 1 ? solid.current_acceleration.p : solid.current_displacement.p;
    DevBuf<double> dp((size_t)n * dim), dv((size_t)n * dim);
    DevBuf<int> df(n);
    dp.upload(pts, (size_t)n * dim, ctx.stream);
    if (dim == 2) query_kernel<2><<<(n + 127) / 128, 128, 0, ctx.stream>>>(S, n, dp.p, nullptr, field, dv.p, df.p);
    else query_kernel<3><<<(n + 127) / 128, 128, 0, ctx.stream>>>(S, n, dp.p, nullptr, field, dv.p, df.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
    dv.download(values, (size_t)n * dim, ctx.stream);
    df.download(found, n, ctx.stream);
  }
} // namespace ifem
