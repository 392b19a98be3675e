// ORACLE (test infrastructure, NOT product code): shared helpers for the CPU
// restatement of OpenIFEM's hot path. Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference arm may load the library built from
// these files.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace oracle
{
  // Small dense tensors in the spirit of dealii::Tensor<rank,dim>; everything is
  // fully dense on purpose so the arithmetic (and its cost) follows the
  // reference's Tensor operations term by term.
  template <int dim>
  struct T1
  {
    double v[dim];
    T1() { for (int i = 0; i < dim; ++i) v[i] = 0; }
    double &operator[](int i) { return v[i]; }
    const double &operator[](int i) const { return v[i]; }
  };

  template <int dim>
  struct T2
  {
    double v[dim][dim];
    T2() { for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) v[i][j] = 0; }
    double *operator[](int i) { return v[i]; }
    const double *operator[](int i) const { return v[i]; }
  };

  template <int dim>
  inline double dot(const T1<dim> &a, const T1<dim> &b)
  {
    double s = 0;
    for (int i = 0; i < dim; ++i) s += a[i] * b[i];
    return s;
  }
  // scalar_product(Tensor<2>, Tensor<2>) = double contraction
  template <int dim>
  inline double scalar_product(const T2<dim> &a, const T2<dim> &b)
  {
    double s = 0;
    for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) s += a[i][j] * b[i][j];
    return s;
  }
  // Tensor<2> * Tensor<1>
  template <int dim>
  inline T1<dim> mul(const T2<dim> &a, const T1<dim> &b)
  {
    T1<dim> r;
    for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) r[i] += a[i][j] * b[j];
    return r;
  }
  template <int dim>
  inline T2<dim> mul(const T2<dim> &a, const T2<dim> &b)
  {
    T2<dim> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        for (int k = 0; k < dim; ++k) r[i][j] += a[i][k] * b[k][j];
    return r;
  }
  template <int dim>
  inline double trace(const T2<dim> &a)
  {
    double s = 0;
    for (int i = 0; i < dim; ++i) s += a[i][i];
    return s;
  }
  inline double det(const T2<2> &a) { return a[0][0] * a[1][1] - a[0][1] * a[1][0]; }
  inline double det(const T2<3> &a)
  {
    return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) -
           a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
           a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
  }
  inline T2<2> invert(const T2<2> &a)
  {
    T2<2> r;
    const double d = 1.0 / det(a);
    r[0][0] = a[1][1] * d;  r[0][1] = -a[0][1] * d;
    r[1][0] = -a[1][0] * d; r[1][1] = a[0][0] * d;
    return r;
  }
  inline T2<3> invert(const T2<3> &a)
  {
    T2<3> r;
    const double d = 1.0 / det(a);
    r[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) * d;
    r[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * d;
    r[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * d;
    r[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) * d;
    r[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * d;
    r[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * d;
    r[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) * d;
    r[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * d;
    r[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * d;
    return r;
  }

  // Q1 geometry of one cell at one reference point: Jacobian J_ij = dx_i/dxi_j.
  template <int dim>
  inline T2<dim> jacobian(const double *vertices, const int *cell_v, const double *dNgeo /*[2^d][dim]*/)
  {
    T2<dim> J;
    for (int v = 0; v < (1 << dim); ++v)
      {
        const double *X = vertices + (size_t)cell_v[v] * dim;
        for (int i = 0; i < dim; ++i)
          for (int j = 0; j < dim; ++j) J[i][j] += X[i] * dNgeo[v * dim + j];
      }
    return J;
  }

  // Position of entry (row, c) in a CSR row with sorted column indices
  // (what PETSc MatSetValues does with a binary search).
  inline int64_t csr_find(const int64_t *rowptr, const int *col, int row, int c)
  {
    const int *b = col + rowptr[row], *e = col + rowptr[row + 1];
    const int *it = std::lower_bound(b, e, c);
    return (it != e && *it == c) ? (it - col) : -1;
  }

  // AffineConstraints::distribute_local_to_global for constraints that are pure
  // Dirichlet lines (no hanging nodes): unconstrained row i gets its
  // unconstrained columns, rhs_i -= sum_j(constrained) K_ij * inhomogeneity_j;
  // a constrained row only receives |K_ii| (or the average |diag| of the local
  // matrix if that is zero) on the diagonal and, when
  // use_inhomogeneities_for_rhs, rhs_i += diag * inhomogeneity_i.
  // Restated from deal.II's documented algorithm
  // (affine_constraints.templates.h, set_matrix_diagonals / resolve_vector_entry);
  // call sites: reference mpi_insim.cpp:348-355, mpi_hyper_elasticity.cpp:507-522.
  inline void distribute_local_to_global(int n, const double *K, const double *f, const int *dofs,
                                         const unsigned char *con, const double *inhom,
                                         const int64_t *rowptr, const int *col, double *A, double *rhs,
                                         bool use_inhomogeneities_for_rhs)
  {
    double average_diagonal = 0;
    for (int i = 0; i < n; ++i) average_diagonal += std::fabs(K[i * n + i]);
    average_diagonal /= n;
    for (int i = 0; i < n; ++i)
      {
        const int gi = dofs[i];
        if (con[gi])
          {
            const double d = std::fabs(K[i * n + i]) != 0 ? std::fabs(K[i * n + i]) : average_diagonal;
            const int64_t p = csr_find(rowptr, col, gi, gi);
#pragma omp atomic
            A[p] += d;
            if (rhs && use_inhomogeneities_for_rhs && inhom)
              {
#pragma omp atomic
                rhs[gi] += d * inhom[gi];
              }
            continue;
          }
        double fi = f ? f[i] : 0.0;
        for (int j = 0; j < n; ++j)
          {
            const int gj = dofs[j];
            const double kij = K[i * n + j];
            if (con[gj])
              {
                if (inhom) fi -= kij * inhom[gj];
                continue;
              }
            if (kij == 0.0) continue; // deal.II elides exact zeros
            const int64_t p = csr_find(rowptr, col, gi, gj);
#pragma omp atomic
            A[p] += kij;
          }
        if (rhs)
          {
#pragma omp atomic
            rhs[gi] += fi;
          }
      }
  }
} // namespace oracle
